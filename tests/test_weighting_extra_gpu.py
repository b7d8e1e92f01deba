"""Exported functions of the weighting / sampling modules that the first golden set did not pin
(tests/golden/extra_*.npz, written by oracle/make_golden.py --only extra from the unmodified
reference): WeightByAlgebraicSpan (hg2v_weighting.py:170-192), the distance half of
WeightByDistanceCluster (:106-134), SameTypeDistanceSample with explicit arguments
(hg2v_sample.py:546-576) and WeightBySameTypeDistance at fixture size (:34-64, 5.69 M stored
entries on snap_youtube_tiny, compared through a digest)."""
import hashlib

import numpy as np
import pytest
import scipy.sparse as sps

from conftest import hypergraph_from_pairs, load_golden
from test_weighting_gpu import ATOL, RTOL, _assert_sparse_close, _embedding, _golden_csr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rand25", "youtube"])
def test_weight_by_algebraic_span_matches_reference(name):
  """The spans come from an embedding the function draws itself (np.random, 5 dimensions, 10
  sweeps): fp32 relaxation here, f64 in the reference, so the weights carry the relaxation's
  own 1e-6-level difference on top of the transform's -- bar 1e-5 relative + 1e-5 absolute on
  weights in [0, 1]; the global RNG must end where the reference's ends."""
  from hypergraphembedding_b200 import WeightByAlgebraicSpan
  g = load_golden("extra_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  for alpha in (0, 0.3):
    np.random.seed(int(g["seed"]))
    n2e, e2n = WeightByAlgebraicSpan(hg, alpha)
    for got, tag in ((n2e, "n2e"), (e2n, "e2n")):
      want = _golden_csr(g, "span_a%s_%s" % (alpha, tag))
      got = sps.csr_matrix(got)
      assert got.shape == want.shape and got.dtype == np.float32
      diff = abs(got - want)
      # (an entry stored on one side only -- the row or column whose scaled span is exactly 0 --
      # counts with its full value here)
      assert diff.max() <= 1e-5 * abs(want).max() + 1e-5, diff.max()
  assert np.random.get_state()[2] == int(g["span_rng_pos"])


@pytest.mark.parametrize("name", ["rand25", "youtube"])
def test_weight_by_distance_cluster(name, monkeypatch):
  """The matrix handed to sklearn's NMF is WeightByDistance's (pinned element by element); the
  factorisation is sklearn's own, so its product is compared with the reference's loosely."""
  import sklearn.decomposition
  from hypergraphembedding_b200 import WeightByDistanceCluster
  g = load_golden("extra_" + name)
  w = load_golden("weights_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  emb = _embedding(g["xn"], g["xe"])
  seen = {}
  real = sklearn.decomposition.NMF

  class Spy(real):

    def fit_transform(self, X, *a, **kw):
      seen["X"] = sps.csr_matrix(X).copy()
      return real.fit_transform(self, X, *a, **kw)

  monkeypatch.setattr(sklearn.decomposition, "NMF", Spy)
  dim = int(g["cluster_dim"])
  W, Ht = WeightByDistanceCluster(hg, 0.3, emb, np.linalg.norm, dim)
  _assert_sparse_close(seen["X"], _golden_csr(w, "wbd_a0.3_n2e"))
  assert sps.issparse(W) and sps.issparse(Ht)
  assert W.shape == g["cluster_w"].shape and Ht.shape == g["cluster_ht"].shape
  got = np.asarray((W @ Ht.T).todense())
  want = g["cluster_w"] @ g["cluster_ht"].T
  assert np.abs(got - want).max() < 5e-3


@pytest.mark.parametrize("name", ["rand25", "youtube"])
def test_same_type_distance_sample_matches_reference(name):
  from hypergraphembedding_b200 import ToCsrMatrix, ToEdgeCsrMatrix
  from hypergraphembedding_b200.hg2v_sample import SameTypeDistanceSample
  g = load_golden("extra_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  emb = _embedding(g["xn"], g["xe"])
  n2e, e2n = ToCsrMatrix(hg), ToEdgeCsrMatrix(hg)
  limit = 120 if name == "youtube" else None      # every call builds its own incidence object
  for tag, m, src, dst, is_edge in (("nn", n2e, emb.node, emb.edge, False),
                                    ("ee", e2n, emb.edge, emb.node, True)):
    left, right, prob = g["st_%s_left" % tag], g["st_%s_right" % tag], g["st_%s_prob" % tag]
    order = np.arange(len(left))
    if limit:
      # half from the random pairs (mostly no shared neighbour), half from the co-member pairs
      order = np.concatenate([order[:limit // 2], order[-limit // 2:]])
    for k in order.tolist():
      i, j = int(left[k]), int(right[k])
      rec = SameTypeDistanceSample((i, j), idx2target=m, source_half_emb=src, target_half_emb=dst,
                                   is_edge=is_edge)
      if is_edge:
        assert (rec.left_edge_idx, rec.right_edge_idx) == (i, j) and rec.left_node_idx is None
        got = rec.edge_edge_prob
      else:
        assert (rec.left_node_idx, rec.right_node_idx) == (i, j) and rec.left_edge_idx is None
        got = rec.node_node_prob
      assert abs(got - prob[k]) <= RTOL * abs(prob[k]) + ATOL, (tag, i, j, got, prob[k])


def test_weight_by_same_type_distance_at_fixture_size():
  """snap_youtube_tiny: 3 862 x 3 862 with 5.69 M stored entries (every pair of nodes that share
  an edge, diagonal included).  Pattern by hash, values by a stride-97 sample and their sum."""
  from hypergraphembedding_b200 import WeightBySameTypeDistance
  g = load_golden("extra_youtube")
  hg = hypergraph_from_pairs(g["pairs"])
  emb = _embedding(g["xn"], g["xe"])
  n2n, e2e = WeightBySameTypeDistance(hg, 0.3, emb, np.linalg.norm, True)
  for tag, m in (("n2n", n2n), ("e2e", e2e)):
    m = sps.csr_matrix(m)
    m.sort_indices()
    assert tuple(m.shape) == tuple(g["wbstd_%s_shape" % tag])
    assert m.nnz == int(g["wbstd_%s_nnz" % tag])
    sha = hashlib.sha256(m.indptr.astype(np.int64).tobytes() +
                         m.indices.astype(np.int64).tobytes()).hexdigest()
    assert sha == str(g["wbstd_%s_pattern_sha" % tag])
    want = g["wbstd_%s_data_strided" % tag]
    got = m.data[::int(g["wbstd_%s_stride" % tag])]
    assert np.all(np.abs(got - want) <= RTOL * np.abs(want) + ATOL)
    total = float(g["wbstd_%s_data_sum" % tag])
    assert abs(float(m.data.astype(np.float64).sum()) - total) <= 1e-5 * abs(total)
    assert abs(float(m.data.min()) - float(g["wbstd_%s_data_min" % tag])) <= ATOL
    assert abs(float(m.data.max()) - float(g["wbstd_%s_data_max" % tag])) <= ATOL
