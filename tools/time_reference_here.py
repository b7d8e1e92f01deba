"""TEST INFRASTRUCTURE -- times the UNMODIFIED reference (through oracle/ref_shim.py) on this
container's CPU cores: EmbedAlgebraicDistance on BASELINE.json configs[0] (the snap_youtube_tiny
fixture, R = 10, 20 sweeps) and on a 1/100-scale member of the config-2 family (10 000 nodes /
5 000 edges / ~100 000 incidences, R = 32, 2 sweeps), single process and with the reference's own
process pool.  /root/reference does not travel to the GPU box, so this line cannot be produced
there; the numbers go to profiles/r2_reference_cpu.md.

    python tools/time_reference_here.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
from hypergraphembedding_b200 import synthetic  # noqa: E402


def to_proto(ref, pairs):
  hg = ref.Hypergraph()
  for n, e in pairs:
    ref.hypergraph_util.AddNodeToEdge(hg, int(n), int(e))
  return hg


def main():
  ref = ref_shim.load_reference()
  g = np.load(os.path.join(ROOT, "tests", "golden", "algdist_youtube.npz"))
  cases = [("configs[0]: snap_youtube_tiny", to_proto(ref, g["pairs"].tolist()), int(g["dim"]), int(g["iters"]))]
  A = synthetic.power_law_hypergraph(10000, 5000, 100000, seed=1234).tocoo()
  cases.append(("1/100 of configs[1]: 10 000 nodes / 5 000 edges / %d incidences" % A.nnz,
                to_proto(ref, zip(A.row.tolist(), A.col.tolist())), 32, 2))
  for name, hg, dim, iters in cases:
    nnz = sum(len(n.edges) for n in hg.node.values())
    for parallel in (False, True):
      np.random.seed(0)
      t = time.perf_counter()
      ref.algebraic_distance.EmbedAlgebraicDistance(hg, dim, iterations=iters, run_in_parallel=parallel,
                                                    disable_pbar=True)
      dt = time.perf_counter() - t
      print(json.dumps({"case": name, "nnz": nnz, "R": dim, "sweeps": iters,
                        "processes": os.cpu_count() if parallel else 1, "seconds": round(dt, 2),
                        "nnz_R_iters_per_s": nnz * dim * iters / dt}), flush=True)


if __name__ == "__main__":
  main()
