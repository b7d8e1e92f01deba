"""cProfile of AlgebraicDistanceSamples on the fixture (configs[0] defaults)."""
import cProfile
import os
import pstats
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hypergraphembedding_b200 as H  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "algdist_youtube.npz"))
node_ids, edge_ids = g["node_ids"], g["edge_ids"]
hg, emb = H.Hypergraph(), H.HypergraphEmbedding()
for n, e in zip(np.searchsorted(node_ids, g["pairs"][:, 0]).tolist(),
                np.searchsorted(edge_ids, g["pairs"][:, 1]).tolist()):
  hg.node[n].edges.append(e)
  hg.edge[e].nodes.append(n)
for i, v in enumerate(g["xn"]):
  emb.node[i].values.extend(v.tolist())
for i, v in enumerate(g["xe"]):
  emb.edge[i].values.extend(v.tolist())
np.random.seed(0)
H.AlgebraicDistanceSamples(hg, emb, 5, 20)
np.random.seed(0)
cProfile.run("out = H.AlgebraicDistanceSamples(hg, emb, 5, 200)", "/tmp/hobe.prof")
pstats.Stats("/tmp/hobe.prof").sort_stats("tottime").print_stats(14)
