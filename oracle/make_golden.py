"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz`` by running the
UNMODIFIED reference (through ``oracle/ref_shim.py``) under fixed seeds.

Run in the dev container, where ``/root/reference`` exists:

    python -m oracle.make_golden [--only NAME]

The GPU box has no reference tree; tests there read only the committed vectors.
Each file records the numpy / scipy versions it was produced with.
"""
import argparse
import hashlib
import os
import random
import sys
import time

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import port  # noqa: E402
from oracle.ref_shim import REFERENCE_ROOT, load_reference  # noqa: E402

VERSIONS = np.array([np.__version__, scipy.__version__])


def incidence_pairs(hg):
  """(node_id, edge_id) pairs in node-map iteration order, then edge.nodes order
  is re-derivable because the fixtures are consistent (checked below)."""
  n2e = [(n, e) for n, d in hg.node.items() for e in d.edges]
  e2n = {(n, e) for e, d in hg.edge.items() for n in d.nodes}
  assert set(n2e) == e2n, "fixture node->edges and edge->nodes disagree"
  return np.asarray(n2e, dtype=np.int64)


def hash_columns(arrays, keys):
  h = hashlib.sha256()
  for k in keys:
    h.update(np.ascontiguousarray(arrays[k], dtype=np.int64).tobytes())
  return h.hexdigest()


INDEX_KEYS = ("left_node", "left_edge", "right_node", "right_edge")
NEIGH_KEYS = ("neigh_node", "neigh_edge")


def emb_to_arrays(emb, n, e, r):
  xn = np.zeros((n, r), np.float32)
  xe = np.zeros((e, r), np.float32)
  for i, v in emb.node.items():
    xn[i] = v.values
  for i, v in emb.edge.items():
    xe[i] = v.values
  return xn, xe


def arrays_to_emb(ref, xn, xe, method="AlgebraicDistance"):
  emb = ref.HypergraphEmbedding()
  emb.dim = xn.shape[1]
  emb.method_name = method
  for i in range(xn.shape[0]):
    emb.node[i].values.extend(xn[i])
  for i in range(xe.shape[0]):
    emb.edge[i].values.extend(xe[i])
  return emb


def graphs(ref):
  U = ref.hypergraph_util
  tiny = ref.Hypergraph()
  for n, e in [(0, 0), (1, 0), (1, 1), (2, 1), (2, 2), (3, 2)]:
    U.AddNodeToEdge(tiny, n, e)  # tests/test_embedding.py:14-22
  random.seed(2024)
  rand25 = U.CreateRandomHyperGraph(25, 25, 0.25)  # tests/test_embedding.py:58
  youtube = ref.Hypergraph()
  with open(os.path.join(REFERENCE_ROOT, "test_data",
                         "snap_youtube_tiny.hypergraph.pb"), "rb") as f:
    youtube.ParseFromString(f.read())
  return {"tiny": tiny, "rand25": rand25, "youtube": youtube}


def save(name, **kw):
  path = os.path.join(GOLDEN, name + ".npz")
  np.savez_compressed(path, versions=VERSIONS, **kw)
  print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


def gen_algdist(ref, name, hg, dim, iters, seed):
  U = ref.hypergraph_util
  np.random.seed(seed)
  t = time.time()
  emb = ref.algebraic_distance.EmbedAlgebraicDistance(
      hg, dim, iterations=iters, run_in_parallel=True, disable_pbar=True)
  dt = time.time() - t
  state = np.random.get_state()
  node_ids = sorted(hg.node)
  edge_ids = sorted(hg.edge)
  xn = np.stack([np.asarray(emb.node[i].values, np.float32) for i in node_ids])
  xe = np.stack([np.asarray(emb.edge[i].values, np.float32) for i in edge_ids])
  save("algdist_" + name, pairs=incidence_pairs(hg), dim=dim, iters=iters,
       seed=seed, xn=xn, xe=xe, node_ids=np.asarray(node_ids),
       edge_ids=np.asarray(edge_ids), rng_pos=state[2],
       rng_key_sha=hashlib.sha256(state[1].tobytes()).hexdigest(),
       method_name=emb.method_name, ref_seconds=dt)
  return xn, xe


def compressed(ref, hg):
  return ref.hypergraph_util.CompressRange(hg)[0]


def csr_parts(m):
  m = m.tocsr()
  m.sort_indices()
  return dict(indptr=m.indptr, indices=m.indices, data=m.data,
              shape=np.asarray(m.shape))


def gen_weights(ref, name, hg, xn, xe, same_type):
  hc = compressed(ref, hg)
  emb = arrays_to_emb(ref, xn, xe)
  W = ref.hg2v_weighting
  out = dict(pairs=incidence_pairs(hc), xn=xn, xe=xe)
  for alpha in (0, 0.3):
    a, b = W.WeightByDistance(hc, alpha, emb, np.linalg.norm, True)
    for tag, m in (("n2e", a), ("e2n", b)):
      for k, v in csr_parts(m).items():
        out["wbd_a%s_%s_%s" % (alpha, tag, k)] = v
    if same_type:
      a, b = W.WeightBySameTypeDistance(hc, alpha, emb, np.linalg.norm, True)
      for tag, m in (("n2n", a), ("e2e", b)):
        for k, v in csr_parts(m).items():
          out["wbstd_a%s_%s_%s" % (alpha, tag, k)] = v
  ns, es = W.ComputeSpans(hc, emb, run_in_parallel=False, disable_pbar=True)
  out["node_span"] = np.asarray([ns[i] for i in range(xn.shape[0])])
  out["edge_span"] = np.asarray([es[i] for i in range(xe.shape[0])])
  for alpha in (0, 0.3):
    a, b = W.WeightByNeighborhood(hc, alpha)
    for tag, m in (("n2e", a), ("e2n", b)):
      for k, v in csr_parts(m).items():
        out["wbn_a%s_%s_%s" % (alpha, tag, k)] = v
  save("weights_" + name, **out)


def gen_extra(ref, name, hg, xn, xe, seed, same_type_digest):
  """The exported functions no other golden pins: WeightByAlgebraicSpan (hg2v_weighting.py:170-192,
  draws its own 5-dimensional embedding from np.random), WeightByDistanceCluster (:106-134),
  SameTypeDistanceSample (hg2v_sample.py:546-576) on pairs with and without a shared neighbour,
  and -- as a digest, the matrix has 5.69 M stored entries on the youtube fixture --
  WeightBySameTypeDistance (:34-64)."""
  hc = compressed(ref, hg)
  emb = arrays_to_emb(ref, xn, xe)
  W, S, U = ref.hg2v_weighting, ref.hg2v_sample, ref.hypergraph_util
  out = dict(pairs=incidence_pairs(hc), xn=xn, xe=xe, seed=seed)
  for alpha in (0, 0.3):
    np.random.seed(seed)
    a, b = W.WeightByAlgebraicSpan(hc, alpha)
    for tag, m in (("n2e", a), ("e2n", b)):
      for k, v in csr_parts(m).items():
        out["span_a%s_%s_%s" % (alpha, tag, k)] = v
  out["span_rng_pos"] = np.random.get_state()[2]
  dim = 3
  wc, hc_t = W.WeightByDistanceCluster(hc, 0.3, emb, np.linalg.norm, dim)
  out["cluster_dim"] = dim
  out["cluster_w"] = np.asarray(wc.todense())
  out["cluster_ht"] = np.asarray(hc_t.todense())
  # same-type probabilities: all co-member pairs of a few rows plus pairs that share nothing
  n2e, e2n = U.ToCsrMatrix(hc), U.ToEdgeCsrMatrix(hc)
  rng = np.random.default_rng(seed)
  for tag, m, src, dst, is_edge in (("nn", n2e, emb.node, emb.edge, False),
                                    ("ee", e2n, emb.edge, emb.node, True)):
    rows = m.shape[0]
    ia = rng.integers(0, rows, 400)
    ib = rng.integers(0, rows, 400)
    co = (m @ m.T).tocoo()
    pick = rng.choice(co.nnz, size=min(600, co.nnz), replace=False)
    ia = np.concatenate([ia, co.row[pick]])
    ib = np.concatenate([ib, co.col[pick]])
    probs = []
    for i, j in zip(ia.tolist(), ib.tolist()):
      rec = S.SameTypeDistanceSample((i, j), idx2target=m, source_half_emb=src, target_half_emb=dst,
                                     is_edge=is_edge)
      probs.append(rec.edge_edge_prob if is_edge else rec.node_node_prob)
      assert (rec.left_edge_idx, rec.right_edge_idx) == (i, j) if is_edge else \
          (rec.left_node_idx, rec.right_node_idx) == (i, j)
    out["st_%s_left" % tag] = ia.astype(np.int32)
    out["st_%s_right" % tag] = ib.astype(np.int32)
    out["st_%s_prob" % tag] = np.asarray(probs, np.float32)
  if same_type_digest:
    t = time.time()
    n2n, e2e = W.WeightBySameTypeDistance(hc, 0.3, emb, np.linalg.norm, True)
    out["wbstd_ref_seconds"] = time.time() - t
    for tag, m in (("n2n", n2n), ("e2e", e2e)):
      p = csr_parts(m)
      out["wbstd_%s_shape" % tag] = p["shape"]
      out["wbstd_%s_nnz" % tag] = m.nnz
      out["wbstd_%s_pattern_sha" % tag] = hashlib.sha256(
          p["indptr"].astype(np.int64).tobytes() + p["indices"].astype(np.int64).tobytes()).hexdigest()
      out["wbstd_%s_stride" % tag] = 97
      out["wbstd_%s_data_strided" % tag] = p["data"][::97].astype(np.float32)
      out["wbstd_%s_data_sum" % tag] = float(p["data"].astype(np.float64).sum())
      out["wbstd_%s_data_min" % tag] = float(p["data"].min())
      out["wbstd_%s_data_max" % tag] = float(p["data"].max())
  save("extra_" + name, **out)


def gen_boolean(ref, name, hg, k, num_samples, neg, seed, full):
  hc = compressed(ref, hg)
  np.random.seed(seed)
  recs = ref.hg2v_sample.BooleanSamples(hc, k, num_samples, neg_samples=neg,
                                        disable_pbar=True)
  state = np.random.get_state()
  arr = port.records_to_arrays(recs, k)
  out = dict(pairs=incidence_pairs(hc), k=k, num_samples=num_samples, neg=neg,
             seed=seed, node_rows=np.asarray(list(hc.node)),
             edge_rows=np.asarray(list(hc.edge)), count=len(recs),
             rng_pos=state[2],
             rng_key_sha=hashlib.sha256(state[1].tobytes()).hexdigest(),
             index_sha=hash_columns(arr, INDEX_KEYS),
             neigh_sha=hash_columns(arr, NEIGH_KEYS))
  if full:
    for key, v in arr.items():
      out["col_" + key] = v.astype(np.float32) if "prob" in key else v.astype(
          np.int32)
  save("boolean_" + name, **out)


def gen_hobe(ref, name, hg, xn, xe, k, num_samples, seed, parallel, stride):
  hc = compressed(ref, hg)
  emb = arrays_to_emb(ref, xn, xe)
  np.random.seed(seed)
  t = time.time()
  recs = ref.hg2v_sample.AlgebraicDistanceSamples(
      hc, emb, k, num_samples, run_in_parallel=parallel, disable_pbar=True)
  dt = time.time() - t
  state = np.random.get_state()
  arr = port.records_to_arrays(recs, k)
  out = dict(pairs=incidence_pairs(hc), xn=xn, xe=xe, k=k,
             num_samples=num_samples, seed=seed, count=len(recs),
             node_rows=np.asarray(list(hc.node)),
             edge_rows=np.asarray(list(hc.edge)), rng_pos=state[2],
             rng_key_sha=hashlib.sha256(state[1].tobytes()).hexdigest(),
             index_sha=hash_columns(arr, INDEX_KEYS), ref_seconds=dt,
             neighbours_deterministic=not parallel, stride=stride)
  if not parallel:
    out["neigh_sha"] = hash_columns(arr, NEIGH_KEYS)
  prob = np.where(~np.isnan(arr["nn_prob"]), arr["nn_prob"],
                  np.where(~np.isnan(arr["ee_prob"]), arr["ee_prob"],
                           arr["ne_prob"])).astype(np.float32)
  if stride == 1:
    for key, v in arr.items():
      if "prob" in key:
        out["col_" + key] = v.astype(np.float32)
      elif parallel and key in NEIGH_KEYS:
        continue
      else:
        out["col_" + key] = v.astype(np.int32)
  else:
    out["prob_strided"] = prob[::stride]
    kind = np.where(~np.isnan(arr["nn_prob"]), 0,
                    np.where(~np.isnan(arr["ee_prob"]), 1, 2)).astype(np.int8)
    out["kind_counts"] = np.bincount(kind, minlength=3)
  save("hobe_" + name, **out)


def gen_jaccard(ref, name, hg, feature_kind, k, num_samples, seed, xn=None, xe=None):
  """WeightedJaccardSamples (hg2v_sample.py:398-510) with run_in_parallel=False (one worker:
  deterministic neighbour draws from a copy of the parent's RNG state)."""
  hc = compressed(ref, hg)
  W = ref.hg2v_weighting
  if feature_kind == "uniform":                    # EmbedHg2vAdjJaccard, embedding.py:339-347
    n2f, e2f = W.UniformWeight(hc)
  elif feature_kind == "neighborhood":             # EmbedHg2vNeighborhoodWeightedJaccard, :369-377
    n2f, e2f = W.WeightByNeighborhood(hc, 0.25)
  else:                                            # distance features of the HOBE weighting
    n2f, e2f = W.WeightByDistance(hc, 0.3, arrays_to_emb(ref, xn, xe), np.linalg.norm, True)
  np.random.seed(seed)
  t = time.time()
  recs = ref.hg2v_sample.WeightedJaccardSamples(hc, n2f, e2f, k, num_samples,
                                                run_in_parallel=False, disable_pbar=True)
  dt = time.time() - t
  state = np.random.get_state()
  arr = port.records_to_arrays(recs, k)
  lw = np.asarray([np.nan if r.left_weight is None else r.left_weight for r in recs], np.float32)
  rw = np.asarray([np.nan if r.right_weight is None else r.right_weight for r in recs], np.float32)
  out = dict(pairs=incidence_pairs(hc), k=k, num_samples=num_samples, seed=seed, count=len(recs),
             feature_kind=feature_kind, node_rows=np.asarray(list(hc.node)),
             edge_rows=np.asarray(list(hc.edge)), rng_pos=state[2],
             rng_key_sha=hashlib.sha256(state[1].tobytes()).hexdigest(), ref_seconds=dt,
             col_left_weight=lw, col_right_weight=rw)
  for tag, m in (("n2f", n2f), ("e2f", e2f)):
    for key, v in csr_parts(m).items():
      out["%s_%s" % (tag, key)] = v
  for key, v in arr.items():
    out["col_" + key] = v.astype(np.float32) if "prob" in key else v.astype(np.int32)
  save("jaccard_" + name, **out)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--only", default=None)
  args = ap.parse_args()
  os.makedirs(GOLDEN, exist_ok=True)
  ref = load_reference()
  g = graphs(ref)

  def want(n):
    return args.only is None or args.only == n

  cfg = {"tiny": (3, 5, 1), "rand25": (5, 10, 2), "youtube": (10, 20, 0)}
  embs = {}
  for name, (dim, iters, seed) in cfg.items():
    path = os.path.join(GOLDEN, "algdist_%s.npz" % name)
    if want("algdist") or not os.path.exists(path):
      embs[name] = gen_algdist(ref, name, g[name], dim, iters, seed)
    else:
      z = np.load(path)
      embs[name] = (z["xn"], z["xe"])
  if want("weights"):
    gen_weights(ref, "tiny", g["tiny"], *embs["tiny"], same_type=True)
    gen_weights(ref, "rand25", g["rand25"], *embs["rand25"], same_type=True)
    gen_weights(ref, "youtube", g["youtube"], *embs["youtube"], same_type=False)
  if want("extra"):
    gen_extra(ref, "rand25", g["rand25"], *embs["rand25"], seed=31, same_type_digest=False)
    gen_extra(ref, "youtube", g["youtube"], *embs["youtube"], seed=32, same_type_digest=True)
  if want("boolean"):
    gen_boolean(ref, "tiny", g["tiny"], 2, 3, 2, 5, full=True)
    gen_boolean(ref, "rand25", g["rand25"], 4, 6, 3, 6, full=True)
    gen_boolean(ref, "youtube_s10", g["youtube"], 5, 10, 0, 7, full=True)
    gen_boolean(ref, "youtube_s200", g["youtube"], 5, 200, 0, 7, full=False)
    gen_boolean(ref, "youtube_s200_neg", g["youtube"], 5, 200, 50, 8, full=False)
  if want("hobe"):
    gen_hobe(ref, "tiny", g["tiny"], *embs["tiny"], k=2, num_samples=3, seed=9,
             parallel=False, stride=1)
    gen_hobe(ref, "rand25", g["rand25"], *embs["rand25"], k=3, num_samples=5,
             seed=10, parallel=False, stride=1)
    gen_hobe(ref, "youtube_s2", g["youtube"], *embs["youtube"], k=5,
             num_samples=2, seed=11, parallel=False, stride=1)
  if want("jaccard"):
    gen_jaccard(ref, "tiny_uniform", g["tiny"], "uniform", 2, 3, 12)
    gen_jaccard(ref, "rand25_uniform", g["rand25"], "uniform", 3, 5, 13)
    gen_jaccard(ref, "rand25_neighborhood", g["rand25"], "neighborhood", 3, 5, 14)
    gen_jaccard(ref, "rand25_distance", g["rand25"], "distance", 3, 5, 15, *embs["rand25"])
    gen_jaccard(ref, "youtube_s2_neighborhood", g["youtube"], "neighborhood", 5, 2, 16)
  if want("hobe_full"):
    # BASELINE.json configs[0]: defaults of embedding.py:389-397.  Pair sets and
    # probabilities are deterministic with run_in_parallel=True; neighbour
    # draws are not and are not recorded.
    gen_hobe(ref, "youtube_s200", g["youtube"], *embs["youtube"], k=5,
             num_samples=200, seed=0, parallel=True, stride=8)


if __name__ == "__main__":
  main()
