"""The per-item helper functions of the reference that its drivers call through process pools
(algebraic_distance.py:54-123, hg2v_sample.py:343-392, 588-629, hg2v_weighting.py:214-233), kept
with explicit arguments for callers that use them directly; each is the one-item form of a
batched kernel and must agree with it and with the oracle."""
import numpy as np
import pytest

from conftest import csr_from_pairs, hypergraph_from_pairs, load_golden
from oracle import port

pytestmark = pytest.mark.gpu


def _emb(xn, xe):
  from hypergraphembedding_b200 import HypergraphEmbedding
  emb = HypergraphEmbedding()
  emb.dim = xn.shape[1]
  for i in range(xn.shape[0]):
    emb.node[i].values.extend(xn[i].tolist())
  for i in range(xe.shape[0]):
    emb.edge[i].values.extend(xe[i].tolist())
  return emb


def test_update_and_scale_helpers_compose_to_the_relaxation():
  from hypergraphembedding_b200.algebraic_distance import (_helper_scale_embeddings,
                                                           _helper_update_embeddings)
  g = load_golden("algdist_rand25")
  r = np.searchsorted(g["node_ids"], g["pairs"][:, 0])
  c = np.searchsorted(g["edge_ids"], g["pairs"][:, 1])
  A = csr_from_pairs(np.stack([r, c], 1), shape=(len(g["node_ids"]), len(g["edge_ids"])))
  B = A.T.tocsr()
  np.random.seed(int(g["seed"]))
  xn, xe = port.algdist_init(A.shape[0], A.shape[1], int(g["dim"]))
  xn, xe = xn.astype(np.float32), xe.astype(np.float32)
  # one sweep by hand against the row-wise oracle
  un, ue = _helper_update_embeddings(None, xn, xe, A, B)
  want_n = port.algdist_rowwise_sweep(A.indptr, A.indices, B.indptr, xn.astype(np.float64),
                                      xe.astype(np.float64))
  want_e = port.algdist_rowwise_sweep(B.indptr, B.indices, A.indptr, xe.astype(np.float64), want_n)
  assert np.abs(un - want_n).max() < 2e-6 and np.abs(ue - want_e).max() < 2e-6
  sn, se = _helper_scale_embeddings(None, un.copy(), ue.copy())
  wn, we = port.algdist_scale(want_n.copy(), want_e.copy())
  assert np.abs(sn - wn).max() < 2e-6 and np.abs(se - we).max() < 2e-6
  assert min(sn.min(), se.min()) == 0.0 and max(sn.max(), se.max()) == 1.0
  # the reference's loop (algebraic_distance.py:149-164) reproduces its committed output
  for _ in range(int(g["iters"])):
    xn, xe = _helper_update_embeddings(None, xn, xe, A, B)
    xn, xe = _helper_scale_embeddings(None, xn, xe)
  assert np.abs(xn - g["xn"]).max() < 2e-5 and np.abs(xe - g["xe"]).max() < 2e-5


def test_diff_type_distance_sample_is_the_batched_record():
  from hypergraphembedding_b200 import AlgebraicDistanceSamples
  from hypergraphembedding_b200.hg2v_sample import DiffTypeDistanceSample
  g = load_golden("hobe_rand25")
  hg = hypergraph_from_pairs(g["pairs"])
  emb = _emb(g["xn"], g["xe"])
  A = csr_from_pairs(g["pairs"])
  B = A.T.tocsr()
  ne = np.nonzero(~np.isnan(g["col_ne_prob"]))[0][:12]
  for i in ne:
    n, e = int(g["col_left_node"][i]), int(g["col_right_edge"][i])
    np.random.seed(3)
    rec = DiffTypeDistanceSample((n, e), A, B, int(g["k"]), emb)
    np.random.seed(3)
    want_e = port.sample_neighbors(n, A, int(g["k"]))
    want_n = port.sample_neighbors(e, B, int(g["k"]))
    assert rec.left_node_idx == n and rec.right_edge_idx == e
    assert np.array_equal(rec.neighbor_edge_indices, want_e)
    assert np.array_equal(rec.neighbor_node_indices, want_n)
    assert abs(rec.node_edge_prob - g["col_ne_prob"][i]) <= 1e-5 * g["col_ne_prob"][i] + 1e-6


def test_diff_type_jaccard_sample_is_the_batched_record():
  from scipy.sparse import csr_matrix
  from hypergraphembedding_b200.hg2v_sample import DiffTypeJaccardSample
  g = load_golden("jaccard_rand25_distance")
  A = csr_from_pairs(g["pairs"])
  B = A.T.tocsr()
  feats = {t: csr_matrix((g[t + "_data"], g[t + "_indices"], g[t + "_indptr"]), shape=tuple(g[t + "_shape"]))
           for t in ("n2f", "e2f")}
  ne = np.nonzero(~np.isnan(g["col_ne_prob"]))[0][:12]
  for i in ne:
    n, e = int(g["col_left_node"][i]), int(g["col_right_edge"][i])
    np.random.seed(4)
    rec = DiffTypeJaccardSample((n, e), A, B, int(g["k"]), feats["n2f"], feats["e2f"])
    assert len(rec.neighbor_edge_indices) == int(g["k"]) == len(rec.neighbor_node_indices)
    for got, key in ((rec.left_weight, "col_left_weight"), (rec.right_weight, "col_right_weight"),
                     (rec.node_edge_prob, "col_ne_prob")):
      assert abs(got - g[key][i]) <= 1e-5 * abs(g[key][i]) + 1e-6, key


def test_compute_span_one_row_equals_compute_spans():
  from hypergraphembedding_b200 import ComputeSpans
  from hypergraphembedding_b200.hg2v_weighting import _compute_span
  g = load_golden("weights_rand25")
  hg = hypergraph_from_pairs(g["pairs"])
  emb = _emb(g["xn"], g["xe"])
  A = csr_from_pairs(g["pairs"])
  node_span, edge_span = ComputeSpans(hg, emb, run_in_parallel=False, disable_pbar=True)
  for n in list(hg.node)[:10]:
    idx, span = _compute_span(n, A, emb.node, emb.edge)
    assert idx == n and abs(span - node_span[n]) < 1e-6 and abs(span - g["node_span"][n]) < 1e-6
  B = A.T.tocsr()
  for e in list(hg.edge)[:10]:
    idx, span = _compute_span(e, B, emb.edge, emb.node)
    assert abs(span - edge_span[e]) < 1e-6
  assert _compute_span(10**6, A, emb.node, emb.edge) == (10**6, 0)
