"""TEST INFRASTRUCTURE ONLY -- digest of HOBE sampling (AlgebraicDistanceSamples,
hg2v_sample.py:632-717) at a scale where its throughput means something: the 100 000-node /
50 000-edge member of the config-4 family (synthetic.zipf_hypergraph, seed 2024), R = 10 vectors
from the legacy seeded generator (so the GPU side starts from the same bits), num_neighbors = 5,
num_samples = 20, np.random.seed(0).  Pair sets and neighbour arrays come from the scipy / numpy
oracle (``oracle.port``: scipy's own product rows, numpy's own legacy RNG; the port is pinned
against the unmodified reference by tests/golden/hobe_*.npz); probabilities are evaluated by the
port's per-record restatement on every STRIDE-th record (the per-record Python is ~1 ms each).

    python -m oracle.make_golden_hobe_scale

Writes tests/golden/hobe_scale.npz: record count, SHA-256 of the index and neighbour columns,
final RNG state, strided probabilities.
"""
import hashlib
import os
import sys
import time

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import port  # noqa: E402
from hypergraphembedding_b200 import synthetic  # noqa: E402  (generators only: pure numpy)

NODES, EDGES, GRAPH_SEED, R, VEC_SEED = 100000, 50000, 2024, 10, 5
K, NUM_SAMPLES, SEED, STRIDE = 5, 20, 0, 499


def sha(cols):
  h = hashlib.sha256()
  for c in cols:
    h.update(np.ascontiguousarray(c, dtype=np.int64).tobytes())
  return h.hexdigest()


def main():
  A = synthetic.zipf_hypergraph(NODES, EDGES, seed=GRAPH_SEED)
  B = A.T.tocsr()
  B.sort_indices()
  n, e = A.shape
  xn, xe = synthetic.legacy_initial_vectors(n, e, R, seed=VEC_SEED)
  np.random.seed(SEED)
  t = time.time()
  nn = port.sample_adj_matrix(A * A.T, range(n), NUM_SAMPLES)
  ee = port.sample_adj_matrix(B * B.T, range(e), NUM_SAMPLES)
  ne = port.sample_adj_matrix(A * A.T * A, range(n), NUM_SAMPLES)
  ne += [(a, b) for b, a in port.sample_adj_matrix(B * B.T * B, range(e), NUM_SAMPLES)]
  parent = np.random.get_state()
  nbr_e = np.empty((len(ne), K), np.int64)
  nbr_n = np.empty((len(ne), K), np.int64)
  for i, (a, b) in enumerate(ne):
    nbr_e[i] = port.sample_neighbors(a, A, K)
    nbr_n[i] = port.sample_neighbors(b, B, K)
  np.random.set_state(parent)
  print("pair sets: %d + %d + %d records in %.1f s" % (len(nn), len(ee), len(ne), time.time() - t),
        flush=True)
  nn, ee, ne = np.asarray(nn, np.int64), np.asarray(ee, np.int64), np.asarray(ne, np.int64)
  m = len(nn) + len(ee) + len(ne)
  none = lambda k: np.full(k, -1, np.int64)
  left_node = np.concatenate([nn[:, 0], none(len(ee)), ne[:, 0]])
  right_node = np.concatenate([nn[:, 1], none(len(ee)), none(len(ne))])
  left_edge = np.concatenate([none(len(nn)), ee[:, 0], none(len(ne))])
  right_edge = np.concatenate([none(len(nn)), ee[:, 1], ne[:, 1]])
  neigh_node = np.concatenate([np.full((len(nn) + len(ee), K), -1, np.int64), nbr_n])
  neigh_edge = np.concatenate([np.full((len(nn) + len(ee), K), -1, np.int64), nbr_e])
  t = time.time()
  probs = []
  for r in range(0, m, STRIDE):
    if r < len(nn):
      probs.append(port.same_type_dist_calc(int(nn[r, 0]), int(nn[r, 1]), A, xn, xe))
    elif r < len(nn) + len(ee):
      i = r - len(nn)
      probs.append(port.same_type_dist_calc(int(ee[i, 0]), int(ee[i, 1]), B, xe, xn))
    else:
      i = r - len(nn) - len(ee)
      probs.append(port.diff_type_prob(int(ne[i, 0]), int(ne[i, 1]), A, B, xn, xe))
  print("%d strided probabilities in %.1f s" % (len(probs), time.time() - t), flush=True)
  state = np.random.get_state()
  out = dict(versions=np.array([np.__version__, scipy.__version__]), nodes=n, edges=e, nnz=int(A.nnz),
             graph_seed=GRAPH_SEED, R=R, vec_seed=VEC_SEED, k=K, num_samples=NUM_SAMPLES, seed=SEED,
             count=m, kind_counts=np.asarray([len(nn), len(ee), len(ne)]),
             index_sha=sha([left_node, left_edge, right_node, right_edge]),
             neigh_sha=sha([neigh_node, neigh_edge]), rng_pos=int(state[2]),
             rng_key_sha=hashlib.sha256(state[1].tobytes()).hexdigest(), stride=STRIDE,
             prob_strided=np.asarray(probs, np.float32),
             csr_sha=hashlib.sha256(A.indptr.astype(np.int64).tobytes() +
                                    A.indices.astype(np.int32).tobytes()).hexdigest())
  np.savez(os.path.join(GOLDEN, "hobe_scale.npz"), **out)
  print({k: (v if np.ndim(v) == 0 else "array%s" % (np.shape(v),)) for k, v in out.items()})


if __name__ == "__main__":
  main()
