"""Does id locality pay?  Times the half-sweeps on config 2 as generated (ids shuffled) and with
the ids relabelled on the host so that storage order = the order the schedule processes rows in
(descending degree, ties by id): nodes only, edges only, both.  One JSON line per labelling."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from bench import WORKLOADS, build_workload  # noqa: E402
from hypergraphembedding_b200 import _native, synthetic  # noqa: E402
from hypergraphembedding_b200 import algebraic_distance as ad  # noqa: E402


def by_degree(deg):
  order = np.lexsort((np.arange(len(deg)), -deg))   # old ids in new order
  return order


def time_it(ctx, A, R, sweeps, label):
  A = A.tocsr()
  A.sort_indices()
  B = A.T.tocsr()
  B.sort_indices()
  N, E = A.shape
  xn0, xe0 = synthetic.legacy_initial_vectors(N, E, R, seed=0)
  inc = ad.make_incidence(A, B, ctx=ctx)
  xn, xe = torch.from_numpy(xn0).cuda(), torch.from_numpy(xe0).cuda()
  st = _native.AlgDistState(ctx, inc, R, sweeps)
  times = []
  for rep in range(4):
    st.load(xn, xe)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * sweeps + 1)]
    evs[0].record()
    for t in range(sweeps):
      st.node_half(t)
      evs[2 * t + 1].record()
      st.edge_half(t)
      evs[2 * t + 2].record()
    torch.cuda.synchronize()
    if rep:
      times.append([evs[i].elapsed_time(evs[i + 1]) for i in range(2 * sweeps)])
  t = np.asarray(times)
  st.close()
  inc.close()
  print(json.dumps(dict(labelling=label, node_ms=float(t[:, 0::2].mean()), edge_ms=float(t[:, 1::2].mean()),
                        sweep_ms=float(t.sum(axis=1).mean() / sweeps))), flush=True)


def main():
  spec = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
  A, _ = build_workload(spec)
  R, sweeps = spec["R"], spec["sweeps"]
  ctx = _native.default_context(0)
  pn = by_degree(np.diff(A.indptr))
  pe = by_degree(np.bincount(A.indices, minlength=A.shape[1]))
  time_it(ctx, A, R, sweeps, "as generated (shuffled ids)")
  time_it(ctx, A[pn], R, sweeps, "nodes in processing order")
  time_it(ctx, A[:, pe], R, sweeps, "edges in processing order")
  time_it(ctx, A[pn][:, pe], R, sweeps, "both in processing order")
  rng = np.random.default_rng(0)
  time_it(ctx, A[rng.permutation(A.shape[0])][:, rng.permutation(A.shape[1])], R, sweeps, "re-shuffled")


if __name__ == "__main__":
  main()
