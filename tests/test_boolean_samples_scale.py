"""FOBE sampling at the scale of BASELINE.json configs[3].

The sample sets are bit-exact against committed digests of the scipy / numpy oracle
(oracle/make_golden_c4.py: scipy's own product rows, numpy's own legacy RNG) on a 100 000-node
hypergraph of the config-4 family; the result is independent of how the rows are chunked and of
the number of worker threads that build candidate rows; the CSR entry point equals the proto
entry point.  Host code only (the probabilities of FOBE are all 1): no GPU needed.  The full
10M-node configuration is timed by ``bench.py --workload c4``, which also checks the induced
100 000-node digest."""
import hashlib

import numpy as np
import pytest

from conftest import hypergraph_from_pairs, load_golden
from hypergraphembedding_b200 import BooleanSamples, SampleColumns, _native, synthetic
from hypergraphembedding_b200.hg2v_sample import BooleanSamplesCsr
from oracle import port

INDEX_KEYS = ("left_node", "left_edge", "right_node", "right_edge")
NEIGH_KEYS = ("neigh_node", "neigh_edge")


def _sha(arrays, keys):
  h = hashlib.sha256()
  for k in keys:
    h.update(np.ascontiguousarray(arrays[k], dtype=np.int64).tobytes())
  return h.hexdigest()


def _same(a, b):
  for f in SampleColumns.FIELDS:
    x, y = getattr(a, f), getattr(b, f)
    assert np.array_equal(x, y, equal_nan=x.dtype.kind == "f"), f


@pytest.fixture()
def sampler_threads():
  lib = _native.load_library()
  yield lib.hge_sampler_set_threads
  lib.hge_sampler_set_threads(0)


def test_c4_family_100k_nodes_bit_exact_against_scipy_oracle_digest():
  g = load_golden("boolean_c4")
  A = synthetic.zipf_hypergraph(100000, 50000, seed=int(g["graph_seed"]))
  csr_sha = hashlib.sha256(A.indptr.astype(np.int64).tobytes() +
                           A.indices.astype(np.int32).tobytes()).hexdigest()
  assert csr_sha == str(g["standalone_csr_sha"]), "the generator no longer reproduces the fixture graph"
  np.random.seed(int(g["seed"]))
  out = BooleanSamplesCsr(A, int(g["k"]), int(g["num_samples"]))
  assert len(out) == int(g["standalone_count"])
  arrays = out.arrays()
  assert _sha(arrays, INDEX_KEYS) == str(g["standalone_index_sha"])
  assert _sha(arrays, NEIGH_KEYS) == str(g["standalone_neigh_sha"])
  state = np.random.get_state()
  assert int(state[2]) == int(g["standalone_rng_pos"])
  assert hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["standalone_rng_key_sha"])
  # every record is a first-order relation of the hypergraph with probability 1
  assert np.all(np.nan_to_num(out.nn_prob, nan=1.0) == 1.0)
  ne = out.has_neighbors
  assert np.all(np.asarray(A[out.left_node[ne], out.right_edge[ne]]).ravel())


@pytest.mark.parametrize("neg", [0, 3])
def test_live_oracle_on_a_small_graph_of_the_family(neg):
  A = synthetic.zipf_hypergraph(3000, 1500, seed=5)
  B = A.T.tocsr()
  B.sort_indices()
  np.random.seed(11)
  want = port.boolean_samples(A, B, range(A.shape[0]), range(A.shape[1]), np.ones(A.shape[0]),
                              np.ones(A.shape[1]), 4, 30, neg_samples=neg)
  want_state = np.random.get_state()
  np.random.seed(11)
  out = BooleanSamplesCsr(A, 4, 30, neg_samples=neg).arrays()
  got_state = np.random.get_state()
  for k in INDEX_KEYS + NEIGH_KEYS:
    assert np.array_equal(out[k], want[k]), k
  for k in ("nn_prob", "ee_prob", "ne_prob"):
    assert np.array_equal(out[k], want[k].astype(np.float32), equal_nan=True), k
  assert got_state[2] == want_state[2] and np.array_equal(got_state[1], want_state[1])


def test_result_does_not_depend_on_chunking_or_threads(sampler_threads):
  A = synthetic.zipf_hypergraph(20000, 10000, seed=7)
  sampler_threads(1)
  np.random.seed(2)
  one = BooleanSamplesCsr(A, 3, 50, neg_samples=2)
  end_state = np.random.get_state()
  for threads, chunk in ((4, 0), (3, 777), (1, 4096)):
    sampler_threads(threads)
    np.random.seed(2)
    parts = list(BooleanSamplesCsr(A, 3, 50, neg_samples=2, chunk_rows=chunk, stream=True))
    assert chunk == 0 or len(parts) > 10
    _same(one, SampleColumns.concatenate(parts))
    state = np.random.get_state()
    assert state[2] == end_state[2] and np.array_equal(state[1], end_state[1])


def test_csr_entry_point_equals_proto_entry_point():
  A = synthetic.zipf_hypergraph(400, 150, seed=3)
  coo = A.tocoo()
  hg = hypergraph_from_pairs(np.stack([coo.row, coo.col], axis=1))
  assert list(hg.node) == list(range(400)) and sorted(hg.edge) == list(range(150))
  edge_rows = list(hg.edge)     # proto-map order of the edges as the loader inserted them
  np.random.seed(9)
  a = BooleanSamples(hg, 2, 20, neg_samples=1)
  np.random.seed(9)
  from hypergraphembedding_b200.hg2v_sample import _Graph
  b = BooleanSamplesCsr(_Graph.from_csr(A, edge_rows=edge_rows), 2, 20, neg_samples=1)
  _same(a, b)
