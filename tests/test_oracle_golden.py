"""The CPU port (oracle/port.py) reproduces the committed outputs of the unmodified reference."""
import hashlib

import numpy as np
import pytest
import scipy.sparse as sps

from conftest import csr_from_pairs, load_golden
from oracle import port


def _shape(g):
  return (len(g["node_ids"]), len(g["edge_ids"]))


def _compressed_csr(g):
  """pairs in the algdist goldens carry original ids; compress by sorted order."""
  node_ids, edge_ids = g["node_ids"], g["edge_ids"]
  pairs = g["pairs"]
  r = np.searchsorted(node_ids, pairs[:, 0])
  c = np.searchsorted(edge_ids, pairs[:, 1])
  return csr_from_pairs(np.stack([r, c], axis=1), shape=_shape(g))


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube"])
def test_algdist_port_matches_reference_output(name):
  g = load_golden("algdist_" + name)
  A = _compressed_csr(g)
  B = A.T.tocsr()
  np.random.seed(int(g["seed"]))
  xn0, xe0 = port.algdist_init(A.shape[0], A.shape[1], int(g["dim"]))
  xn, xe = port.algdist_vectorised(A, B, xn0, xe0, int(g["iters"]))
  # the reference stores fp32 (proto float); the f64 restatement differs by f64 rounding only
  assert np.abs(xn.astype(np.float32) - g["xn"]).max() <= 2e-7
  assert np.abs(xe.astype(np.float32) - g["xe"]).max() <= 2e-7
  state = np.random.get_state()
  assert int(state[2]) == int(g["rng_pos"])
  assert hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["rng_key_sha"])


def test_algdist_rowwise_port_is_bit_exact_on_small_graphs():
  for name in ("tiny", "rand25"):
    g = load_golden("algdist_" + name)
    A = _compressed_csr(g)
    np.random.seed(int(g["seed"]))
    xn0, xe0 = port.algdist_init(A.shape[0], A.shape[1], int(g["dim"]))
    xn, xe = port.algdist_rowwise(A, A.T.tocsr(), xn0, xe0, int(g["iters"]))
    assert np.array_equal(xn.astype(np.float32), g["xn"])
    assert np.array_equal(xe.astype(np.float32), g["xe"])


def _golden_csr(g, prefix):
  return sps.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]),
                        shape=tuple(g[prefix + "_shape"]))


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube"])
def test_weight_by_distance_port(name):
  g = load_golden("weights_" + name)
  A = csr_from_pairs(g["pairs"], shape=(g["xn"].shape[0], g["xe"].shape[0]))
  for alpha in (0, 0.3):
    n2e, e2n = port.weight_by_distance(A, g["xn"], g["xe"], alpha)
    ref_n2e = _golden_csr(g, "wbd_a%s_n2e" % alpha)
    ref_e2n = _golden_csr(g, "wbd_a%s_e2n" % alpha)
    assert n2e.nnz == ref_n2e.nnz and e2n.nnz == ref_e2n.nnz
    assert abs(n2e - ref_n2e).max() == 0
    assert abs(e2n - ref_e2n).max() == 0


@pytest.mark.parametrize("name", ["tiny", "rand25"])
def test_weight_by_same_type_distance_port(name):
  g = load_golden("weights_" + name)
  A = csr_from_pairs(g["pairs"], shape=(g["xn"].shape[0], g["xe"].shape[0]))
  for alpha in (0, 0.3):
    n2n, e2e = port.weight_by_same_type_distance(A, A.T.tocsr(), g["xn"], g["xe"], alpha)
    assert abs(n2n - _golden_csr(g, "wbstd_a%s_n2n" % alpha)).max() == 0
    assert abs(e2e - _golden_csr(g, "wbstd_a%s_e2e" % alpha)).max() == 0


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube"])
def test_spans_port(name):
  g = load_golden("weights_" + name)
  A = csr_from_pairs(g["pairs"], shape=(g["xn"].shape[0], g["xe"].shape[0]))
  assert np.array_equal(port.compute_span_rows(A, g["xn"], g["xe"]), g["node_span"])
  assert np.array_equal(port.compute_span_rows(A.T.tocsr(), g["xe"], g["xn"]), g["edge_span"])


def _sample_inputs(g):
  shape = (int(g["node_rows"].max()) + 1, int(g["edge_rows"].max()) + 1)
  A = csr_from_pairs(g["pairs"], shape=shape)
  return A, A.T.tocsr()


def _index_sha(arrays, keys):
  h = hashlib.sha256()
  for k in keys:
    h.update(np.ascontiguousarray(arrays[k], dtype=np.int64).tobytes())
  return h.hexdigest()


INDEX_KEYS = ("left_node", "left_edge", "right_node", "right_edge")
NEIGH_KEYS = ("neigh_node", "neigh_edge")


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube_s10", "youtube_s200"])
def test_boolean_samples_port(name):
  g = load_golden("boolean_" + name)
  A, B = _sample_inputs(g)
  np.random.seed(int(g["seed"]))
  ones_n = [1.0] * len(g["node_rows"])
  ones_e = [1.0] * len(g["edge_rows"])
  out = port.boolean_samples(A, B, g["node_rows"].tolist(), g["edge_rows"].tolist(), ones_n, ones_e,
                             int(g["k"]), int(g["num_samples"]), int(g["neg"]))
  assert len(out["left_node"]) == int(g["count"])
  assert _index_sha(out, INDEX_KEYS) == str(g["index_sha"])
  assert _index_sha(out, NEIGH_KEYS) == str(g["neigh_sha"])
  state = np.random.get_state()
  assert int(state[2]) == int(g["rng_pos"])
  assert hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["rng_key_sha"])
  if "col_left_node" in g:
    for k in INDEX_KEYS + NEIGH_KEYS:
      assert np.array_equal(out[k], g["col_" + k])


@pytest.mark.parametrize("name", ["tiny", "rand25"])
def test_hobe_samples_port(name):
  g = load_golden("hobe_" + name)
  A, B = _sample_inputs(g)
  np.random.seed(int(g["seed"]))
  out = port.algebraic_distance_samples(A, B, g["node_rows"].tolist(), g["edge_rows"].tolist(),
                                        g["xn"], g["xe"], int(g["k"]), int(g["num_samples"]))
  assert len(out["left_node"]) == int(g["count"])
  for k in INDEX_KEYS + NEIGH_KEYS:
    assert np.array_equal(out[k], g["col_" + k]), k
  for k in ("nn_prob", "ee_prob", "ne_prob"):
    assert np.array_equal(np.isnan(out[k]), np.isnan(g["col_" + k]))
    assert np.allclose(np.nan_to_num(out[k]), np.nan_to_num(g["col_" + k]), rtol=0, atol=1e-7)
  state = np.random.get_state()
  assert int(state[2]) == int(g["rng_pos"])
  assert hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["rng_key_sha"])


def test_lifo_rows_and_rng_replay_match_scipy_and_numpy():
  g = load_golden("boolean_rand25")
  A, B = _sample_inputs(g)
  AA = A * A.T
  AAA = AA * A
  for i in range(A.shape[0]):
    row = port.spgemm_row_lifo(A.indices[A.indptr[i]:A.indptr[i + 1]], B.indptr, B.indices)
    assert row == AA[i, :].nonzero()[1].tolist()
    row3 = port.spgemm_row_lifo(AA.indices[AA.indptr[i]:AA.indptr[i + 1]], A.indptr, A.indices)
    assert row3 == AAA[i, :].nonzero()[1].tolist()
  np.random.seed(77)
  rp = port.MT19937Replay()
  got = ([rp.random() for _ in range(5)], rp.permutation(17), rp.choice_replace(9, 6),
         rp.choice_replace(1, 3), rp.choice_no_replace(1000, 4))
  want = (list(np.random.random(5)), list(np.random.permutation(17)),
          list(np.random.choice(np.arange(9), 6, replace=True)),
          list(np.random.choice(np.arange(1), 3, replace=True)),
          list(np.random.choice(np.arange(1000), 4, replace=False)))
  assert got == tuple(want)
  st, st2 = rp.state_tuple(), np.random.get_state()
  assert (st[1] == st2[1]).all() and st[2] == st2[2]


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube"])
def test_c_restatement_reproduces_the_reference_bit_for_bit(name):
  """oracle/algdist_ref.c (bench.py's cpu_baseline / --impl reference) performs the reference's
  f64 operations in the reference's order: after the fp32 store its output is identical to the
  unmodified reference's, for 1 thread and for all threads."""
  from oracle import cport
  g = load_golden("algdist_" + name)
  A = _compressed_csr(g)
  B = A.T.tocsr()
  B.sort_indices()
  np.random.seed(int(g["seed"]))
  xn0, xe0 = port.algdist_init(A.shape[0], A.shape[1], int(g["dim"]))
  for threads in (1, 0):
    xn, xe = cport.algdist(A, B, xn0, xe0, int(g["iters"]), threads=threads)
    assert np.array_equal(xn.astype(np.float32), g["xn"])
    assert np.array_equal(xe.astype(np.float32), g["xe"])
