"""WeightedJaccardSamples (HG2V_ADJ_JAC / HG2V_NEIGH_JAC, hg2v_sample.py:250-510): pair sets,
neighbour arrays AND the Jaccard probabilities / left / right weights bit for bit equal to the
unmodified reference (committed golden outputs, oracle/make_golden.py --only jaccard) -- the
kernels carry the two sums in the reference's own order -- plus the reference's own known-answer
vectors (tests/test_hg2v_samples.py:192-304)."""
import hashlib

import numpy as np
import pytest
import scipy.sparse as sps
from scipy.sparse import csr_matrix

from conftest import hypergraph_from_pairs, load_golden

pytestmark = pytest.mark.gpu

INDEX_KEYS = ("left_node", "left_edge", "right_node", "right_edge")
NEIGH_KEYS = ("neigh_node", "neigh_edge")
RTOL, ATOL = 1e-5, 1e-6   # relative tolerance of the north star; weights live in [0, 1]


def _features(g, tag):
  return csr_matrix((g[tag + "_data"], g[tag + "_indices"], g[tag + "_indptr"]),
                    shape=tuple(g[tag + "_shape"]))


def _float64_restatement(g, arrays):
  """SparseWeightedJaccard / CentroidFromRows / DiffTypeJaccardSample (hg2v_sample.py:250-392) in
  dense float64 for the sampled pairs."""
  NF = _features(g, "n2f").toarray().astype(np.float64)
  EF = _features(g, "e2f").toarray().astype(np.float64)
  A = np.zeros((NF.shape[0], EF.shape[0]))
  A[g["pairs"][:, 0], g["pairs"][:, 1]] = 1
  node_centroid = (A @ EF) / np.maximum(A.sum(1), 1)[:, None]        # mean of the node's edges' rows
  edge_centroid = (A.T @ NF) / np.maximum(A.sum(0), 1)[:, None]      # mean of the edge's nodes' rows

  def jac(x, y):
    hi = np.maximum(x, y).sum(1)
    return np.where(hi > 0, np.minimum(x, y).sum(1) / np.where(hi > 0, hi, 1), 0)

  ln, rn, le, re = (arrays[k].astype(np.int64) for k in ("left_node", "right_node", "left_edge",
                                                         "right_edge"))
  nn, ee, ne = (ln >= 0) & (rn >= 0), (le >= 0) & (re >= 0), (ln >= 0) & (re >= 0)
  out = {k: np.full(len(ln), np.nan) for k in ("nn_prob", "ee_prob", "ne_prob", "left_weight",
                                                "right_weight")}
  out["nn_prob"][nn] = jac(NF[ln[nn]], NF[rn[nn]])
  out["ee_prob"][ee] = jac(EF[le[ee]], EF[re[ee]])
  out["left_weight"][ne] = jac(NF[ln[ne]], edge_centroid[re[ne]])
  out["right_weight"][ne] = jac(EF[re[ne]], node_centroid[ln[ne]])
  out["ne_prob"][ne] = out["left_weight"][ne] * out["right_weight"][ne]
  return out


@pytest.mark.parametrize("name", ["tiny_uniform", "rand25_uniform", "rand25_neighborhood",
                                  "rand25_distance", "youtube_s2_neighborhood"])
def test_weighted_jaccard_samples_match_reference(name):
  from hypergraphembedding_b200 import WeightedJaccardSamples
  g = load_golden("jaccard_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  assert list(hg.node) == g["node_rows"].tolist() and list(hg.edge) == g["edge_rows"].tolist()
  np.random.seed(int(g["seed"]))
  out = WeightedJaccardSamples(hg, _features(g, "n2f"), _features(g, "e2f"), int(g["k"]),
                               int(g["num_samples"]), run_in_parallel=False, disable_pbar=True)
  state = np.random.get_state()
  assert int(state[2]) == int(g["rng_pos"])
  assert hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["rng_key_sha"])
  assert len(out) == int(g["count"])
  arrays = out.arrays()
  for k in INDEX_KEYS + NEIGH_KEYS:
    assert np.array_equal(arrays[k], g["col_" + k]), k          # bit-exact sample sets
  truth = _float64_restatement(g, arrays)
  for k in ("nn_prob", "ee_prob", "ne_prob", "left_weight", "right_weight"):
    assert np.array_equal(np.isnan(arrays[k]), np.isnan(g["col_" + k])), k
    got, want = np.nan_to_num(arrays[k]), np.nan_to_num(g["col_" + k])
    assert np.all((got >= 0) & (got <= 1 + ATOL))
    # The reference adds its minima / maxima one by one in fp32 in ascending column order
    # (hg2v_sample.py:262-271) and builds centroids with scipy's column sums (members ascending,
    # fp32); the kernels do the same, so every record is the reference's number exactly.
    assert np.array_equal(got.astype(np.float32), want.astype(np.float32)), \
        (k, int((got != want).sum()), float(np.abs(got - want).max()))
    # ... and stays as close to the float64 value of the same formula as one-by-one fp32 sums over
    # rows of thousands of non-zeros get (the reference's own accuracy: a few 1e-6 absolute)
    exact = np.nan_to_num(truth[k])
    assert np.all(np.abs(got - exact) <= RTOL * np.abs(exact) + 1e-5), (k, np.abs(got - exact).max())


def test_reference_known_answers_sparse_weighted_jaccard():
  """tests/test_hg2v_samples.py:194-208."""
  from hypergraphembedding_b200 import SparseWeightedJaccard
  assert SparseWeightedJaccard(csr_matrix([0, 1, 1, 0, 1], dtype=bool),
                               csr_matrix([1, 0, 1, 0, 1], dtype=bool)) == 0.5
  assert abs(SparseWeightedJaccard(csr_matrix([0, 3, 4, 0, 1], dtype=np.int32),
                                   csr_matrix([2, 0, 1, 0, 1], dtype=np.int32)) - 0.2) < 1e-7
  assert SparseWeightedJaccard(csr_matrix((1, 4), dtype=np.float32),
                               csr_matrix((1, 4), dtype=np.float32)) == 0     # 0 / 0 -> 0


def test_reference_known_answers_same_type_sample():
  """tests/test_hg2v_samples.py:243-281."""
  from hypergraphembedding_b200.hg2v_sample import SameTypeJaccardSample
  feats = csr_matrix([[1, 0], [1, 0], [1, 2], [0, 2]])
  recs = [SameTypeJaccardSample((0, j), feats, False) for j in (1, 2, 3)]
  assert [r.right_node_idx for r in recs] == [1, 2, 3] and all(r.left_node_idx == 0 for r in recs)
  assert np.allclose([r.node_node_prob for r in recs], [1, 1 / 3, 0], atol=1e-7)
  rec = SameTypeJaccardSample((0, 2), feats, True)
  assert rec.left_edge_idx == 0 and rec.right_edge_idx == 2 and abs(rec.edge_edge_prob - 1 / 3) < 1e-7
  assert rec.node_node_prob is None


def test_reference_known_answer_centroid():
  """tests/test_hg2v_samples.py:294-303, and the on-the-fly centroid of the kernel against it."""
  from hypergraphembedding_b200 import _native
  from hypergraphembedding_b200.hg2v_sample import CentroidFromRows
  target2features = csr_matrix([[1, 2, 3], [0, 1, 0], [1, 0, 3]], dtype=np.float32)
  idx2targets = csr_matrix([[1, 0, 1]])
  got = csr_matrix(CentroidFromRows(0, idx2targets, target2features), shape=(1, 3))
  assert np.abs(got.toarray() - [[1, 1, 3]]).max() < 1e-5
  x = csr_matrix([[2, 0, 1]], dtype=np.float32)
  # J(x, centroid) = (min(2,1) + min(0,1) + min(1,3)) / (max(2,1) + max(0,1) + max(1,3)) = 2 / 6
  ctx = _native.default_context()
  out = _native.jaccard_centroid(ctx, _native.FeatureCsr(x), _native.CsrArrays(idx2targets),
                                 _native.FeatureCsr(target2features), [0], [0])
  assert abs(out[0] - 2.0 / 6.0) < 1e-6


def test_kernels_against_a_dense_restatement_on_random_features():
  from hypergraphembedding_b200 import _native
  rng = np.random.default_rng(0)
  F = sps.random(300, 200, density=0.08, random_state=1, format="csr", dtype=np.float32)
  G = sps.random(150, 300, density=0.05, random_state=2, format="csr", dtype=np.float32)
  G.data[:] = 1
  ctx = _native.default_context()
  feat, groups = _native.FeatureCsr(F), _native.CsrArrays(G)
  pi, pj = rng.integers(0, 300, 4000), rng.integers(0, 300, 4000)
  got = _native.jaccard_rows(ctx, feat, pi, pj)
  D = F.toarray().astype(np.float64)
  lo, hi = np.minimum(D[pi], D[pj]).sum(1), np.maximum(D[pi], D[pj]).sum(1)
  want = np.where(hi > 0, lo / np.where(hi > 0, hi, 1), 0)
  assert np.abs(got - want).max() < 2e-6
  px, pg = rng.integers(0, 300, 3000), rng.integers(0, 150, 3000)
  got = _native.jaccard_centroid(ctx, feat, groups, feat, px, pg)
  cnt = np.asarray(G.sum(1)).ravel()
  C = (G.astype(np.float64) @ D) / np.where(cnt > 0, cnt, 1)[:, None]
  lo, hi = np.minimum(D[px], C[pg]).sum(1), np.maximum(D[px], C[pg]).sum(1)
  want = np.where((hi > 0) & (cnt[pg] > 0), lo / np.where(hi > 0, hi, 1), 0)
  assert np.abs(got - want).max() < 2e-6


def _reference_loop(xi, xv, yi, yv):
  """SparseWeightedJaccard's loop (hg2v_sample.py:257-275) on (sorted ids, fp32 values) rows."""
  x, y = dict(zip(xi.tolist(), xv)), dict(zip(yi.tolist(), yv))
  num = den = 0
  for c in np.union1d(xi[xv != 0], yi[yv != 0]):
    a, b = x.get(int(c), np.float32(0)), y.get(int(c), np.float32(0))
    if a < b:
      num += a
      den += b
    else:
      num += b
      den += a
  return np.float32(0) if den == 0 else np.float32(num / den)


def test_kernels_equal_the_reference_loop_bit_for_bit_negative_values_included():
  """Rows and centroids against a restatement of the reference's loop in numpy fp32 scalars:
  long rows (hundreds of non-zeros, where the summation order shows), groups of 1 .. 60 members,
  an empty group, explicit zeros and negative feature values (the reference's formula has no sign
  restriction: the smaller value goes to the numerator)."""
  from hypergraphembedding_b200 import _native
  rng = np.random.default_rng(7)
  F = sps.random(120, 900, density=0.35, random_state=3, format="csr", dtype=np.float32)
  F.data[::17] *= -1
  F.data[::29] = 0
  F.sort_indices()
  G = sps.random(40, 120, density=0.2, random_state=4, format="csr", dtype=np.float32)
  G.data[:] = 1
  G = sps.vstack([G, csr_matrix((1, 120), dtype=np.float32)]).tocsr()     # group 40 is empty
  G.sort_indices()
  ctx = _native.default_context()
  feat, groups = _native.FeatureCsr(F), _native.CsrArrays(G)
  row = lambda M, r: (M.indices[M.indptr[r]:M.indptr[r + 1]], M.data[M.indptr[r]:M.indptr[r + 1]])
  pi, pj = rng.integers(0, 120, 300), rng.integers(0, 120, 300)
  got = _native.jaccard_rows(ctx, feat, pi, pj)
  want = np.asarray([_reference_loop(*row(F, a), *row(F, b)) for a, b in zip(pi, pj)], np.float32)
  assert np.array_equal(got, want)
  px, pg = rng.integers(0, 120, 300), np.concatenate([rng.integers(0, 40, 299), [40]])
  got = _native.jaccard_centroid(ctx, feat, groups, feat, px, pg)
  want = []
  for a, grp in zip(px, pg):
    members = G.indices[G.indptr[grp]:G.indptr[grp + 1]]
    if len(members) == 0:
      want.append(np.float32(0))
      continue
    centroid = F[members].sum(axis=0) / len(members)        # CentroidFromRows, hg2v_sample.py:293
    assert centroid.dtype == np.float32
    cols = centroid.nonzero()[1]
    want.append(_reference_loop(*row(F, a), cols, np.asarray(centroid)[0, cols]))
  assert np.array_equal(got, np.asarray(want, np.float32))


def test_weighted_model_input_carries_the_jaccard_weights():
  """SamplesToModelInput(weighted=True) puts left / right weights into the features
  (hg2v_sample.py:767-770)."""
  from hypergraphembedding_b200 import SamplesToModelInput, WeightedJaccardSamples
  g = load_golden("jaccard_rand25_neighborhood")
  hg = hypergraph_from_pairs(g["pairs"])
  np.random.seed(int(g["seed"]))
  out = WeightedJaccardSamples(hg, _features(g, "n2f"), _features(g, "e2f"), int(g["k"]),
                               int(g["num_samples"]), run_in_parallel=False)
  feats, targets = SamplesToModelInput(out, int(g["k"]), weighted=True)
  k = int(g["k"])
  assert len(feats) == 6 + 4 * k
  assert np.allclose(feats[4], np.nan_to_num(g["col_left_weight"]), rtol=RTOL, atol=ATOL)
  assert np.allclose(feats[5], np.nan_to_num(g["col_right_weight"]), rtol=RTOL, atol=ATOL)
  records = list(out)[-3:]
  feats_r, _ = SamplesToModelInput(records, k, weighted=True)    # record path gives the same
  assert np.allclose(feats_r[4], feats[4][-3:]) and np.allclose(feats_r[5], feats[5][-3:])
