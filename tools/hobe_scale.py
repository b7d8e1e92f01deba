"""HOBE sampling on a 100K-node hypergraph of the config-4 family: where the time goes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hypergraphembedding_b200 as H  # noqa: E402
from hypergraphembedding_b200 import synthetic  # noqa: E402

n, e = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, int(sys.argv[2]) if len(sys.argv) > 2 else 50000
A = synthetic.zipf_hypergraph(n, e, seed=2024)
B = A.T.tocsr()
t = time.time()
hg = H.Hypergraph()
for i in range(A.shape[0]):
  hg.node[i].edges.extend(A.indices[A.indptr[i]:A.indptr[i + 1]].tolist())
for j in range(B.shape[0]):
  hg.edge[j].nodes.extend(B.indices[B.indptr[j]:B.indptr[j + 1]].tolist())
print("proto built in %.2f s (%d incidences)" % (time.time() - t, A.nnz))
np.random.seed(0)
t = time.time()
emb = H.EmbedAlgebraicDistance(hg, 10, iterations=20, disable_pbar=True)
print("EmbedAlgebraicDistance %.3f s" % (time.time() - t))
for rep in range(2):
  np.random.seed(1)
  t = time.time()
  out = H.AlgebraicDistanceSamples(hg, emb, 5, 200, disable_pbar=True)
  dt = time.time() - t
  print("AlgebraicDistanceSamples: %d records in %.3f s = %.3g samples/s" % (len(out), dt, len(out) / dt))
import cProfile, pstats
np.random.seed(1)
cProfile.run("H.AlgebraicDistanceSamples(hg, emb, 5, 200, disable_pbar=True)", "/tmp/hs.prof")
pstats.Stats("/tmp/hs.prof").sort_stats("tottime").print_stats(8)
