"""The host sampler of libhge_b200.so (csrc/hge_sampler.cpp) against numpy / scipy themselves:
raw MT19937 stream, bounded integers, candidate-row order of scipy's CSR product, and the
full _sample_adj_matrix / _sample_neighbors draws.  No GPU needed: this is host code."""
import numpy as np
import pytest
import scipy.sparse as sps

from conftest import csr_from_pairs, load_golden
from hypergraphembedding_b200 import _native
from oracle import port


def _graph(seed, n, e, nnz):
  rng = np.random.default_rng(seed)
  rows = np.concatenate([rng.integers(0, n, nnz), np.arange(n)])
  cols = np.concatenate([rng.integers(0, e, nnz), rng.integers(0, e, n)])
  return csr_from_pairs(np.stack([rows, cols], 1), shape=(n, e))


def test_raw_stream_and_state_roundtrip():
  np.random.seed(123)
  st = _native.LegacyRngState()
  out = np.empty(2000, dtype=np.uint32)
  _native.check(_native.load_library().hge_mt19937_random_raw(_native.ptr(st.buf), 2000,
                                                              _native.ptr(out)))
  bg = np.random.MT19937()
  legacy = np.random.get_state()
  bg.state = {"bit_generator": "MT19937", "state": {"key": legacy[1], "pos": legacy[2]}}
  assert np.array_equal(out, bg.random_raw(2000).astype(np.uint32))
  st.commit()                       # global state now sits 2000 draws further
  a = np.random.random(3)
  np.random.seed(123)
  np.random.get_state()
  bg2 = np.random.RandomState(123)
  bg2.randint(0, 2**32, size=2000, dtype=np.uint32)   # consumes exactly 2000 raw words
  assert np.array_equal(a, bg2.random_sample(3))


@pytest.mark.parametrize("mx", [0, 1, 2, 5, 31, 32, 1000, 65535, 2**31 - 1])
def test_bounded_integers_match_numpy_randint(mx):
  np.random.seed(mx % 97)
  st = _native.LegacyRngState()
  out = np.empty(500, dtype=np.uint32)
  _native.check(_native.load_library().hge_mt19937_interval(_native.ptr(st.buf), mx, 500,
                                                            _native.ptr(out)))
  want = np.random.randint(0, mx + 1, size=500)
  assert np.array_equal(out.astype(np.int64), want)
  after = np.random.get_state()
  assert np.array_equal(st.buf[:624], after[1]) and int(st.buf[624]) == int(after[2])


def test_product_rows_in_scipy_order():
  A = _graph(1, 60, 23, 150)
  At = A.T.tocsr()
  a, at = _native.CsrArrays(A), _native.CsrArrays(At)
  rows = np.arange(A.shape[0])
  for mats, ref in (((a, at), A * A.T), ((a, at, a), A * A.T * A), ((at, a), At * At.T),
                    ((at, a, at), At * At.T * At)):
    r = np.arange(ref.shape[0])
    ptr_, idx_ = _native.spgemm_rows(mats, r)
    assert np.array_equal(ptr_, ref.indptr)
    assert np.array_equal(idx_, ref.indices)            # scipy's own (unsorted) storage order
    ptr_s, idx_s = _native.spgemm_rows(mats, r, sorted_rows=True)
    srt = ref.copy()
    srt.sort_indices()
    assert np.array_equal(idx_s, srt.indices)
  del rows


@pytest.mark.parametrize("replace,negative", [(False, False), (True, False), (False, True)])
def test_sample_adj_rows_matches_the_port(replace, negative):
  A = _graph(2, 80, 17, 260)
  At = A.T.tocsr()
  a, at = _native.CsrArrays(A), _native.CsrArrays(At)
  rows = list(np.random.default_rng(0).permutation(A.shape[0]))
  per_row = [int(v) for v in np.random.default_rng(1).integers(0, 9, len(rows))]
  for mats, matrix in (((a,), A), ((a, at), A * A.T), ((a, at, a), A * A.T * A)):
    np.random.seed(5)
    want = port.sample_adj_matrix(matrix, rows, per_row, replace=replace, negative=negative)
    want_state = np.random.get_state()
    np.random.seed(5)
    st = _native.LegacyRngState()
    r, c = _native.sample_adj_rows(mats, rows, per_row, st, replace=replace, negative=negative)
    assert list(zip(r.tolist(), c.tolist())) == want
    assert np.array_equal(st.buf[:624], want_state[1]) and int(st.buf[624]) == int(want_state[2])


def test_sample_neighbors_matches_numpy_choice():
  A = _graph(3, 50, 12, 120)
  B = A.T.tocsr()
  a, b = _native.CsrArrays(A), _native.CsrArrays(B)
  rng = np.random.default_rng(4)
  coo = A.tocoo()
  pick = rng.integers(0, A.nnz, 200)
  nodes, edges = coo.row[pick], coo.col[pick]
  np.random.seed(9)
  want_e, want_n = [], []
  for n, e in zip(nodes, edges):
    want_e.append(port.sample_neighbors(n, A, 5))
    want_n.append(port.sample_neighbors(e, B, 5))
  want_state = np.random.get_state()
  np.random.seed(9)
  st = _native.LegacyRngState()
  got_e, got_n = _native.sample_neighbors(a, b, nodes, edges, 5, st)
  assert np.array_equal(got_e, np.asarray(want_e)) and np.array_equal(got_n, np.asarray(want_n))
  assert np.array_equal(st.buf[:624], want_state[1]) and int(st.buf[624]) == int(want_state[2])


def test_single_candidate_rows_consume_no_draws():
  A = sps.csr_matrix(np.eye(6, dtype=bool))
  a = _native.CsrArrays(A)
  np.random.seed(1)
  before = np.random.get_state()
  st = _native.LegacyRngState()
  r, c = _native.sample_adj_rows((a, a), list(range(6)), [3] * 6, st)
  assert r.tolist() == c.tolist() == list(range(6))
  assert np.array_equal(st.buf[:624], before[1]) and int(st.buf[624]) == int(before[2])


def test_fixture_scale_candidates_match_golden_hash():
  """All four products on the youtube fixture reproduce scipy's rows (5.69M / 138 / 2 x 34 165
  stored entries)."""
  g = load_golden("boolean_youtube_s10")
  shape = (int(g["node_rows"].max()) + 1, int(g["edge_rows"].max()) + 1)
  A = csr_from_pairs(g["pairs"], shape=shape)
  B = A.T.tocsr()
  a, b = _native.CsrArrays(A), _native.CsrArrays(B)
  for mats, ref in (((a, b), A * A.T), ((b, a), B * B.T), ((a, b, a), A * A.T * A),
                    ((b, a, b), B * B.T * B)):
    ptr_, idx_ = _native.spgemm_rows(mats, np.arange(ref.shape[0]))
    assert np.array_equal(ptr_, ref.indptr) and np.array_equal(idx_, ref.indices)


@pytest.mark.parametrize("n", [2, 15, 16, 17, 31, 32, 33, 47, 48, 100, 255, 256, 257, 271, 272, 273, 1000,
                               4095, 4096, 4097, 4112, 70000])
def test_shuffle_of_every_length_class_equals_numpy_choice(n):
  """np.random.choice(cols, k, replace=False) on a row of n candidates: the block filter of the
  draw loop (16 words at a time where the bound is far from a power of two, scalar near the
  boundaries) must consume the stream exactly as numpy's rk_interval does -- same samples, same
  final state -- for lengths on both sides of every switch-over."""
  import scipy.sparse as sps
  from hypergraphembedding_b200.hg2v_sample import _sample_adj_matrix
  cols = np.arange(0, 3 * n, 3)
  m = sps.csr_matrix((np.ones(n, dtype=bool), cols, [0, n]), shape=(1, 3 * n))
  for seed, k in ((1, 1), (2, min(n, 7)), (3, n)):
    np.random.seed(seed)
    want = np.random.choice(cols, k, replace=False)
    want_state = np.random.get_state()
    np.random.seed(seed)
    got = _sample_adj_matrix(m, [0], k)
    state = np.random.get_state()
    assert [c for _, c in got] == want.tolist()
    assert state[2] == want_state[2] and np.array_equal(state[1], want_state[1])
