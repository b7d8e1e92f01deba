"""Distance weighting, spans and pair distances (CUDA, through the C ABI) against the committed
outputs of the unmodified reference and against the oracle port.  Tolerance (north star):
distances and weights within 1e-5 relative in fp32, stated below as rtol 1e-5 + atol 1e-6 on
values in [0, 1]."""
import numpy as np
import pytest
import scipy.sparse as sps

from conftest import csr_from_pairs, hypergraph_from_pairs, load_golden

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


def _embedding(xn, xe):
  from hypergraphembedding_b200 import HypergraphEmbedding
  emb = HypergraphEmbedding()
  emb.dim = xn.shape[1]
  emb.method_name = "AlgebraicDistance"
  for i in range(xn.shape[0]):
    emb.node[i].values.extend(xn[i].tolist())
  for i in range(xe.shape[0]):
    emb.edge[i].values.extend(xe[i].tolist())
  return emb


def _golden_csr(g, prefix):
  return sps.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]),
                        shape=tuple(g[prefix + "_shape"]))


def _assert_sparse_close(actual, expected):
  assert actual.shape == expected.shape
  a, e = actual.tocsr(), expected.tocsr()
  a.sort_indices()
  e.sort_indices()
  # element by element, |a - e| <= 1e-5 |e| + 1e-6 (the north star's relative bar plus the
  # absolute floor a weight near 0 needs); an entry stored on one side only is the dropped exact
  # zero of the other and must itself be within the floor
  diff = (a - e).tocoo()
  ref = np.abs(np.asarray(e[diff.row, diff.col]).ravel()) if diff.nnz else np.zeros(0)
  bad = np.abs(diff.data) > RTOL * ref + ATOL
  assert not bad.any(), "%d of %d entries off, worst |a - e| = %g" % (bad.sum(), e.nnz, np.abs(diff.data).max())
  only_one_side = (a != 0).astype(np.int8) - (e != 0).astype(np.int8)
  assert abs(only_one_side).sum() <= 1


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube"])
def test_weight_by_distance_matches_reference(name):
  from hypergraphembedding_b200 import WeightByDistance
  g = load_golden("weights_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  emb = _embedding(g["xn"], g["xe"])
  for alpha in (0, 0.3):
    n2e, e2n = WeightByDistance(hg, alpha, emb, np.linalg.norm, True)
    assert n2e.dtype == np.float32 and e2n.dtype == np.float32
    _assert_sparse_close(n2e, _golden_csr(g, "wbd_a%s_n2e" % alpha))
    _assert_sparse_close(e2n, _golden_csr(g, "wbd_a%s_e2n" % alpha))
    if alpha == 0:
      # the farthest incidence gets weight exactly 0 and is not stored
      assert n2e.nnz == _golden_csr(g, "wbd_a0_n2e").nnz == len(g["pairs"]) - 1


@pytest.mark.parametrize("name", ["tiny", "rand25"])
def test_weight_by_same_type_distance_matches_reference(name):
  from hypergraphembedding_b200 import WeightBySameTypeDistance
  g = load_golden("weights_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  emb = _embedding(g["xn"], g["xe"])
  for alpha in (0, 0.3):
    n2n, e2e = WeightBySameTypeDistance(hg, alpha, emb, np.linalg.norm, True)
    _assert_sparse_close(n2n, _golden_csr(g, "wbstd_a%s_n2n" % alpha))
    _assert_sparse_close(e2e, _golden_csr(g, "wbstd_a%s_e2e" % alpha))


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube"])
def test_compute_spans_matches_reference(name):
  from hypergraphembedding_b200 import ComputeSpans
  g = load_golden("weights_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  emb = _embedding(g["xn"], g["xe"])
  node2span, edge2span = ComputeSpans(hg, emb, run_in_parallel=False, disable_pbar=True)
  assert sorted(node2span) == list(range(len(g["node_span"])))
  got_n = np.asarray([node2span[i] for i in range(len(g["node_span"]))])
  got_e = np.asarray([edge2span[i] for i in range(len(g["edge_span"]))])
  assert np.allclose(got_n, g["node_span"], rtol=RTOL, atol=ATOL)
  assert np.allclose(got_e, g["edge_span"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", ["rand25", "youtube"])
def test_weight_by_neighborhood_matches_reference(name):
  """Host-side scheme (degrees only), pinned to keep the whole module a drop-in."""
  from hypergraphembedding_b200 import WeightByNeighborhood
  g = load_golden("weights_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  for alpha in (0, 0.3):
    n2e, e2n = WeightByNeighborhood(hg, alpha)
    assert abs(sps.csr_matrix(n2e) - _golden_csr(g, "wbn_a%s_n2e" % alpha)).max() <= 1e-7
    assert abs(sps.csr_matrix(e2n) - _golden_csr(g, "wbn_a%s_e2n" % alpha)).max() <= 1e-7


def test_non_l2_norm_is_refused():
  from hypergraphembedding_b200 import WeightByDistance
  g = load_golden("weights_tiny")
  hg = hypergraph_from_pairs(g["pairs"])
  with pytest.raises(NotImplementedError):
    WeightByDistance(hg, 0, _embedding(g["xn"], g["xe"]), lambda v: np.abs(v).sum(), True)
  with pytest.raises(AssertionError):
    WeightByDistance(hg, 1.5, _embedding(g["xn"], g["xe"]), np.linalg.norm, True)


@pytest.mark.parametrize("R", [1, 3, 4, 10, 32, 64, 100, 128, 200])
def test_pair_and_incidence_distances_all_dimensions(R, gpu_ctx):
  from hypergraphembedding_b200 import _native
  rng = np.random.default_rng(R)
  n, e = 500, 70
  rows = np.concatenate([rng.integers(0, n, 3000), np.arange(n), np.zeros(e, int)])
  cols = np.concatenate([rng.integers(0, e, 3000), rng.integers(0, e, n), np.arange(e)])
  A = csr_from_pairs(np.stack([rows, cols], 1), shape=(n, e))    # node 0 is in every edge
  B = A.T.tocsr()
  xn = rng.random((n, R)).astype(np.float32)
  xe = rng.random((e, R)).astype(np.float32)
  inc = _native.Incidence(gpu_ctx, n, e, A.indptr, A.indices, B.indptr, B.indices)
  try:
    coo = A.tocoo()
    want = np.sqrt(((xn[coo.row].astype(np.float64) - xe[coo.col])**2).sum(1))
    got = _native.incidence_l2(gpu_ctx, inc, xn, xe, order=0)
    assert np.allclose(got, want, rtol=RTOL, atol=ATOL)
    coo_b = B.tocoo()
    want_b = np.sqrt(((xe[coo_b.row].astype(np.float64) - xn[coo_b.col])**2).sum(1))
    got_b = _native.incidence_l2(gpu_ctx, inc, xn, xe, order=1)
    assert np.allclose(got_b, want_b, rtol=RTOL, atol=ATOL)
    w = _native.incidence_l2(gpu_ctx, inc, xn, xe, order=0, as_weight=True)
    assert np.allclose(w, (np.sqrt(R) - want) / np.sqrt(R), rtol=RTOL, atol=ATOL)
    ia = rng.integers(0, n, 4097).astype(np.int32)
    ib = rng.integers(0, e, 4097).astype(np.int32)
    want_p = np.sqrt(((xn[ia].astype(np.float64) - xe[ib])**2).sum(1))
    assert np.allclose(_native.pair_l2(gpu_ctx, xn, xe, ia, ib), want_p, rtol=RTOL, atol=ATOL)
    same = _native.pair_l2(gpu_ctx, xn, xn, ia, ia)
    assert np.all(same == 0)
  finally:
    inc.close()


def test_scale_transform_is_bit_exact_with_the_port(gpu_ctx):
  from hypergraphembedding_b200 import _native
  from oracle import port
  rng = np.random.default_rng(0)
  for n in (1, 2, 1000, 100003):
    v = (rng.random(n) * 3).astype(np.float32)
    for alpha in (0, 0.3, 1):
      want = np.asarray(port.alpha_scale(port.one_minus(port.zero_one_scale(v)), alpha), np.float32)
      got, mm = _native.scale_transform(gpu_ctx, v.copy(), alpha, want_minmax=True)
      assert np.array_equal(got, want)
      assert mm[0] == v.min() and mm[1] == v.max()
  const = np.full(17, 2.5, np.float32)
  assert np.array_equal(_native.scale_transform(gpu_ctx, const.copy(), 0.25), np.full(17, 0.25, np.float32))
  empty = np.zeros(0, np.float32)
  assert len(_native.scale_transform(gpu_ctx, empty, 0.5)) == 0


def test_device_resident_pair_weighting(gpu_ctx):
  """configs[2] shape in miniature: pairs, vectors and results stay on the device."""
  import torch
  from hypergraphembedding_b200 import _native
  rng = np.random.default_rng(1)
  xn = rng.random((2000, 64)).astype(np.float32)
  xe = rng.random((3000, 64)).astype(np.float32)
  ia = rng.integers(0, 2000, 100000).astype(np.int32)
  ib = rng.integers(0, 3000, 100000).astype(np.int32)
  d = _native.pair_l2(gpu_ctx, torch.from_numpy(xn).cuda(), torch.from_numpy(xe).cuda(),
                      torch.from_numpy(ia).cuda(), torch.from_numpy(ib).cuda())
  assert d.is_cuda
  _native.scale_transform(gpu_ctx, d, 0.1)
  want = np.sqrt(((xn[ia].astype(np.float64) - xe[ib])**2).sum(1))
  lo, hi = want.min(), want.max()
  want_w = 0.1 + 0.9 * (1 - (want - lo) / (hi - lo))
  assert np.allclose(d.cpu().numpy(), want_w, rtol=RTOL, atol=2e-6)


def test_two_step_scale_transform_equals_the_fused_one(gpu_ctx):
  """hge_scale_minmax + hge_scale_apply (the sharded form) give the bits of
  hge_scale_transform; the sharded pair weighting on one rank equals pair_l2 + scale_transform."""
  import socket
  import torch.distributed as dist
  from hypergraphembedding_b200 import _native
  from hypergraphembedding_b200 import distributed as hd
  rng = np.random.default_rng(4)
  v = (rng.random(100000) * 7).astype(np.float32)
  fused = _native.scale_transform(gpu_ctx, v.copy(), 0.3)
  lo, hi = _native.scale_minmax(gpu_ctx, v)
  assert (lo, hi) == (float(v.min()), float(v.max()))
  assert np.array_equal(_native.scale_apply(gpu_ctx, v.copy(), 0.3, lo, hi), fused)
  xa, xb = rng.random((400, 20)).astype(np.float32), rng.random((90, 20)).astype(np.float32)
  ia, ib = rng.integers(0, 400, 5000).astype(np.int32), rng.integers(0, 90, 5000).astype(np.int32)
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port_no = s.getsockname()[1]
  s.close()
  dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port_no, rank=0, world_size=1)
  try:
    got = hd.sharded_pair_weights(xa, xb, ia, ib, 0.1, ctx=gpu_ctx)
  finally:
    dist.destroy_process_group()
  want = _native.scale_transform(gpu_ctx, _native.pair_l2(gpu_ctx, xa, xb, ia, ib), 0.1)
  assert np.array_equal(got, want)
