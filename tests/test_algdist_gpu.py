"""Parity of the CUDA relaxation (through the C ABI) with the reference / the oracle port.

Bar (BASELINE.json north_star): incidence distances within 1e-5 relative of the reference's
numpy result from the same initial vectors; the reference computes in f64 and stores fp32,
the kernels compute in fp32.
"""
import numpy as np
import pytest
import scipy.sparse as sps

from conftest import (csr_from_pairs, embedding_arrays, hypergraph_from_pairs, incidence_distances,
                      load_golden)

pytestmark = pytest.mark.gpu

RTOL = 1e-5          # north star: distances and weights within 1e-5 relative, fp32
ATOL_WEIGHT = 1e-6   # absolute slack on HOBE weights in [0, 1] (they reach exactly 0)


def assert_distance_parity(A, xn, xe, ref_xn, ref_xe):
  """|d - d_ref| <= 1e-5 d_ref + 1e-6 sqrt(R) on every incidence distance, and the same bound
  on the HOBE weight w = (sqrt(R) - d) / sqrt(R) (hg2v_sample.py:537-541):
  |w - w_ref| <= 1e-5 |w_ref| + 1e-6.

  A purely relative bound is undefined for coincident pairs: the reference stores fp32, and
  that rounding alone moves its own smallest fixture distance (6.5e-7) by 9%.  The absolute
  term is 1e-6 in weight units, i.e. 1e-6 sqrt(R) in distance units (~16 ulp of a unit-range
  fp32 coordinate)."""
  R = np.asarray(ref_xn).shape[1]
  d = incidence_distances(A, xn, xe)
  dr = incidence_distances(A, ref_xn, ref_xe)
  bound = RTOL * dr + ATOL_WEIGHT * np.sqrt(R)
  worst = (np.abs(d - dr) / bound).max()
  assert worst <= 1.0, "distance error is %.2f x the allowed bound" % worst
  w = (np.sqrt(R) - d) / np.sqrt(R)
  wr = (np.sqrt(R) - dr) / np.sqrt(R)
  worst_w = (np.abs(w - wr) / (RTOL * np.abs(wr) + ATOL_WEIGHT)).max()
  assert worst_w <= 1.0, "weight error is %.2f x the allowed bound" % worst_w
  return worst, worst_w


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube"])
def test_embed_algebraic_distance_matches_reference_golden(name):
  from hypergraphembedding_b200 import EmbedAlgebraicDistance
  g = load_golden("algdist_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  np.random.seed(int(g["seed"]))
  emb = EmbedAlgebraicDistance(hg, int(g["dim"]), iterations=int(g["iters"]), run_in_parallel=False,
                               disable_pbar=True)
  assert emb.dim == int(g["dim"])
  assert emb.method_name == "AlgebraicDistance"          # algebraic_distance.py:168
  assert sorted(emb.node) == g["node_ids"].tolist()
  assert sorted(emb.edge) == g["edge_ids"].tolist()
  xn, xe = embedding_arrays(emb, g["node_ids"], g["edge_ids"])
  r = np.searchsorted(g["node_ids"], g["pairs"][:, 0])
  c = np.searchsorted(g["edge_ids"], g["pairs"][:, 1])
  A = csr_from_pairs(np.stack([r, c], 1), shape=(len(g["node_ids"]), len(g["edge_ids"])))
  assert_distance_parity(A, xn, xe, g["xn"], g["xe"])
  # the call consumed the global RNG exactly as the reference does
  assert int(np.random.get_state()[2]) == int(g["rng_pos"])


def _run(A, xn0, xe0, iters, gpu_ctx, lohi=None, device=False):
  from hypergraphembedding_b200 import algebraic_distance as ad
  inc = ad.make_incidence(A, ctx=gpu_ctx)
  try:
    if device:
      import torch
      xn = torch.from_numpy(xn0.copy()).cuda()
      xe = torch.from_numpy(xe0.copy()).cuda()
      ad.relax(inc, xn, xe, iters, lohi=lohi)
      torch.cuda.synchronize()
      return xn.cpu().numpy(), xe.cpu().numpy()
    xn, xe = xn0.copy(), xe0.copy()
    ad.relax(inc, xn, xe, iters, lohi=lohi)
    return xn, xe
  finally:
    inc.close()


def _random_graph(rng, n, e, nnz):
  rows = rng.integers(0, n, nnz)
  cols = rng.integers(0, e, nnz)
  rows = np.concatenate([rows, np.arange(n), rng.integers(0, n, e)])
  cols = np.concatenate([cols, rng.integers(0, e, n), np.arange(e)])
  return csr_from_pairs(np.stack([rows, cols], 1), shape=(n, e))


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 8, 10, 16, 31, 32, 33, 64, 100, 128, 130, 200])
def test_all_dimensions_against_oracle(R, gpu_ctx):
  from oracle import port
  rng = np.random.default_rng(R)
  A = _random_graph(rng, 700, 90, 3000)
  xn0 = rng.random((700, R)).astype(np.float32)
  xe0 = rng.random((90, R)).astype(np.float32)
  ref_xn, ref_xe = port.algdist_vectorised(A, A.T.tocsr(), xn0, xe0, 6)
  xn, xe = _run(A, xn0, xe0, 6, gpu_ctx)
  assert_distance_parity(A, xn, xe, ref_xn, ref_xe)


@pytest.mark.parametrize("device", [False, True])
def test_long_rows_and_device_pointers(device, gpu_ctx):
  """Rows that take the multi-chunk path (degree >> chunk) on both sides, plus min/max log."""
  from oracle import port
  rng = np.random.default_rng(5)
  n, e = 6000, 300
  rows = np.concatenate([np.arange(n), np.arange(n), rng.integers(0, n, 20000), np.zeros(e, int),
                         np.ones(e, int)])
  cols = np.concatenate([np.zeros(n, int), rng.integers(0, e, n), rng.integers(0, e, 20000),
                         np.arange(e), np.arange(e)])
  A = csr_from_pairs(np.stack([rows, cols], 1), shape=(n, e))   # edge 0 holds every node
  R = 32
  xn0 = rng.random((n, R)).astype(np.float32)
  xe0 = rng.random((e, R)).astype(np.float32)
  iters = 8
  ref_xn, ref_xe = port.algdist_vectorised(A, A.T.tocsr(), xn0, xe0, iters)
  lohi = np.zeros((iters, 2, R), np.float32)
  xn, xe = _run(A, xn0, xe0, iters, gpu_ctx, lohi=lohi, device=device)
  assert_distance_parity(A, xn, xe, ref_xn, ref_xe)
  assert np.all(lohi[:, 0] < lohi[:, 1])
  # after the rescale every column spans [0, 1] exactly (joint over nodes and edges)
  both = np.concatenate([xn, xe])
  assert np.allclose(both.min(0), 0, atol=1e-6) and np.allclose(both.max(0), 1, atol=1e-6)


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("both", [False, True])
def test_one_call_from_csr_equals_the_two_call_path(device, both, gpu_ctx):
  """hge_algdist_run_csr (incidence set-up + relaxation in one call; with host buffers the
  vectors' upload overlaps the set-up) gives the bits of hge_incidence_create + hge_algdist_run,
  with one orientation or both supplied, from host or device arrays; twice in a row (the staging
  block and the workspace pool are re-used); an isolated edge raises as the reference does."""
  import torch
  from hypergraphembedding_b200 import _native
  from hypergraphembedding_b200.hypergraph_util import csr_arrays
  rng = np.random.default_rng(17)
  A = _random_graph(rng, 20000, 700, 150000)
  B = A.T.tocsr()
  B.sort_indices()
  R, iters = 32, 7
  xn0 = rng.random((A.shape[0], R)).astype(np.float32)
  xe0 = rng.random((A.shape[1], R)).astype(np.float32)
  want_n, want_e = _run(A, xn0, xe0, iters, gpu_ctx)
  a_ptr, a_idx = csr_arrays(A)
  b_ptr, b_idx = csr_arrays(B)
  for rep in range(2):
    if device:
      dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
      xn, xe = dev(xn0), dev(xe0)
      args = [dev(a_ptr.astype(np.int64)), dev(a_idx.astype(np.int32))]
      kw = dict(e2n_ptr=dev(b_ptr.astype(np.int64)), e2n_idx=dev(b_idx.astype(np.int32))) if both else {}
    else:
      xn, xe = xn0.copy(), xe0.copy()
      args, kw = [a_ptr, a_idx], (dict(e2n_ptr=b_ptr, e2n_idx=b_idx) if both else {})
    lohi = np.zeros((iters, 2, R), np.float32)
    _native.algdist_run_csr(gpu_ctx, A.shape[0], A.shape[1], args[0], args[1], xn, xe, iters, lohi=lohi, **kw)
    if device:
      torch.cuda.synchronize()
      xn, xe = xn.cpu().numpy(), xe.cpu().numpy()
    assert np.array_equal(xn, want_n) and np.array_equal(xe, want_e), rep
    assert np.all(lohi[:, 0] < lohi[:, 1])
  # an edge nobody is in: 0 / 0 in the reference (algebraic_distance.py:49)
  A2 = sps.hstack([A, sps.csr_matrix((A.shape[0], 1), dtype=A.dtype)]).tocsr()
  p2, i2 = csr_arrays(A2)
  with pytest.raises(ZeroDivisionError):
    _native.algdist_run_csr(gpu_ctx, A2.shape[0], A2.shape[1], p2, i2, xn0.copy(),
                            np.zeros((A2.shape[1], R), np.float32), iters)


def test_tuning_knobs_do_not_change_results(gpu_ctx):
  rng = np.random.default_rng(11)
  A = _random_graph(rng, 3000, 200, 40000)
  xn0 = rng.random((3000, 32)).astype(np.float32)
  xe0 = rng.random((200, 32)).astype(np.float32)
  base = _run(A, xn0, xe0, 5, gpu_ctx)
  try:
    for light, chunk in ((8, 32), (255, 1024), (1, 64)):
      gpu_ctx.set_tuning(light, chunk, 2)
      xn, xe = _run(A, xn0, xe0, 5, gpu_ctx)
      assert np.abs(xn - base[0]).max() < 2e-6 and np.abs(xe - base[1]).max() < 2e-6
  finally:
    gpu_ctx.reset_tuning()      # the library's own defaults, not a copy of them


def test_short_rows_single_and_multi_chunk_long_rows_against_the_oracle(gpu_ctx):
  """A graph with short rows, single-chunk and multi-chunk long rows on both sides, R a multiple
  of 4 and not (lanes beyond the row, padding columns)."""
  from oracle import port
  rng = np.random.default_rng(17)
  n, e = 9000, 400
  rows = np.concatenate([np.arange(n), rng.integers(0, n, 50000), rng.integers(0, 300, 30000),
                         rng.integers(0, n, e)])
  cols = np.concatenate([np.zeros(n, int), rng.integers(0, e, 50000), rng.integers(0, e, 30000),
                         np.arange(e)])
  A = csr_from_pairs(np.stack([rows, cols], 1), shape=(n, e))
  for R in (32, 10):
    xn0 = rng.random((n, R)).astype(np.float32)
    xe0 = rng.random((e, R)).astype(np.float32)
    ref_xn, ref_xe = port.algdist_vectorised(A, A.T.tocsr(), xn0, xe0, 6)
    xn, xe = _run(A, xn0, xe0, 6, gpu_ctx)
    assert_distance_parity(A, xn, xe, ref_xn, ref_xe)


def test_results_are_reproducible_run_to_run(gpu_ctx):
  rng = np.random.default_rng(12)
  A = _random_graph(rng, 5000, 100, 60000)
  xn0 = rng.random((5000, 32)).astype(np.float32)
  xe0 = rng.random((100, 32)).astype(np.float32)
  a = _run(A, xn0, xe0, 10, gpu_ctx)
  b = _run(A, xn0, xe0, 10, gpu_ctx)
  assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_zero_iterations_returns_the_initial_vectors(gpu_ctx):
  rng = np.random.default_rng(1)
  A = _random_graph(rng, 50, 10, 100)
  xn0 = rng.random((50, 7)).astype(np.float32)
  xe0 = rng.random((10, 7)).astype(np.float32)
  xn, xe = _run(A, xn0, xe0, 0, gpu_ctx)
  assert np.array_equal(xn, xn0) and np.array_equal(xe, xe0)


def test_isolated_edge_raises_like_the_reference():
  """An edge key without members: the reference divides 0/0 (algebraic_distance.py:49)."""
  from hypergraphembedding_b200 import EmbedAlgebraicDistance, Hypergraph
  hg = Hypergraph()
  hg.node[0].edges.extend([0])
  hg.node[1].edges.extend([0])
  hg.edge[0].nodes.extend([0, 1])
  hg.edge[1].weight = 1.0      # present, but nobody is in it
  with pytest.raises(ZeroDivisionError):
    EmbedAlgebraicDistance(hg, 2, iterations=1, disable_pbar=True)


def test_config2_scale_properties(gpu_ctx):
  """BASELINE.json configs[1] shape at 1/10 size against the oracle, full size by properties."""
  from hypergraphembedding_b200 import synthetic
  from oracle import port
  A = synthetic.power_law_hypergraph(100000, 50000, 1000000, seed=7)
  xn0, xe0 = synthetic.legacy_initial_vectors(A.shape[0], A.shape[1], 32, seed=3)
  ref_xn, ref_xe = port.algdist_vectorised(A, A.T.tocsr(), xn0, xe0, 20)
  xn, xe = _run(A, xn0, xe0, 20, gpu_ctx)
  assert_distance_parity(A, xn, xe, ref_xn, ref_xe)


def test_node_range_tiles_of_the_edge_half(gpu_ctx):
  """Single-GPU edge half in node-range tiles (hge_ctx_set_tile_mb): same result as the oracle
  and, up to fp32 summation order, as the untiled half-sweep; deterministic."""
  from hypergraphembedding_b200 import _native, synthetic
  from hypergraphembedding_b200 import algebraic_distance as ad
  from oracle import port
  A = synthetic.power_law_hypergraph(60000, 900, 400000, seed=3, max_edge_size=30000)
  R, iters = 32, 6
  xn0, xe0 = synthetic.legacy_initial_vectors(A.shape[0], A.shape[1], R, seed=1)
  want_n, want_e = port.algdist_vectorised(A, A.T.tocsr(), xn0, xe0, iters)
  results = []
  try:
    for tile_mb in (0, 1, 1, 2):
      gpu_ctx.set_tile_mb(tile_mb, 0)      # 1 MB = 8192 rows of 32 floats -> 8 tiles
      inc = ad.make_incidence(A, ctx=gpu_ctx)
      xn, xe = xn0.copy(), xe0.copy()
      ad.relax(inc, xn, xe, iters)
      inc.close()
      results.append((xn, xe))
      assert np.abs(xn - want_n).max() < 2e-5 and np.abs(xe - want_e).max() < 2e-5, tile_mb
  finally:
    gpu_ctx.reset_tuning()
  assert np.array_equal(results[1][0], results[2][0]) and np.array_equal(results[1][1], results[2][1])
  assert np.abs(results[0][1] - results[1][1]).max() < 5e-6


@pytest.mark.parametrize("device", [False, True])
def test_edge_to_node_orientation_built_on_the_device(device, gpu_ctx):
  """hge_incidence_create with e2n == NULL: the transpose the library builds (stable radix sort of
  the (edge, node) pairs) gives the bits the caller-supplied orientation gives."""
  import torch
  from hypergraphembedding_b200 import _native
  rng = np.random.default_rng(23)
  A = _random_graph(rng, 4000, 300, 30000)
  B = A.T.tocsr()
  B.sort_indices()
  xn0 = rng.random((4000, 32)).astype(np.float32)
  xe0 = rng.random((300, 32)).astype(np.float32)
  results = []
  for with_e2n in (True, False):
    arrays = [A.indptr.astype(np.int64), A.indices.astype(np.int32)]
    if with_e2n:
      arrays += [B.indptr.astype(np.int64), B.indices.astype(np.int32)]
    if device:
      arrays = [torch.from_numpy(a).cuda() for a in arrays]
    inc = _native.Incidence(gpu_ctx, 4000, 300, *arrays)
    xn = torch.from_numpy(xn0).cuda() if device else xn0.copy()
    xe = torch.from_numpy(xe0).cuda() if device else xe0.copy()
    _native.algdist_run(gpu_ctx, inc, xn, xe, 4)
    # the weighting kernels walk the edge -> node arrays directly
    d = _native.incidence_l2(gpu_ctx, inc, xn, xe, order=1)
    inc.close()
    results.append([np.asarray(t.cpu() if device else t) for t in (xn, xe, d)])
  for a, b in zip(*results):
    assert np.array_equal(a, b)
