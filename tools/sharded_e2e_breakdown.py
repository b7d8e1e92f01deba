"""Where the host-buffer step of the sharded relaxation spends its time (torchrun, N ranks)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from bench import WORKLOADS, build_workload, init_process_group_quiet  # noqa: E402
from hypergraphembedding_b200 import _native, synthetic  # noqa: E402
from hypergraphembedding_b200 import distributed as hd  # noqa: E402


def main():
  rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
  local = int(os.environ["LOCAL_RANK"])
  torch.cuda.set_device(local)
  init_process_group_quiet(rank, world, torch.device("cuda", local))
  spec = dict(WORKLOADS["c2"], seed=WORKLOADS["c2"]["seed"] + rank)
  A, B = build_workload(spec)
  R, sweeps = spec["R"], spec["sweeps"]
  ctx = _native.default_context(local)
  xn0, xe0 = synthetic.legacy_initial_vectors(A.shape[0], A.shape[1], R, seed=rank)
  xe0 = synthetic.legacy_initial_vectors(1, A.shape[1], R, seed=10**6)[1]
  pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
  h_xn, h_xe = pin(xn0), pin(xe0)
  h_csr = [pin(A.indptr.astype(np.int64)), pin(A.indices.astype(np.int32))]   # one orientation
  marks = []

  def mark(name):
    torch.cuda.synchronize()
    marks.append((name, time.perf_counter()))

  reps = int(os.environ.get("HGE_BREAKDOWN_REPS", "4"))
  history = []
  for rep in range(reps):
    dist.barrier()
    torch.cuda.synchronize()
    marks.clear()
    mark("start")
    ops = hd.NativeOps(None, R, sweeps, 1, ctx=ctx, shape=A.shape,
                       csr_host=[t.numpy() for t in h_csr])
    mark("incidence create (sharded)")
    deg, wsum = ops.edge_sums()
    dist.all_reduce(deg)
    dist.all_reduce(wsum)
    mark("all-reduce of edge degrees / weight sums")
    bad = bool((deg == 0).any())
    mark("empty-edge check")
    ops.finish()
    mark("finish (schedules, state)")
    ops.enable_p2p(dist, None)
    mark("peer arena")
    ops.load(h_xn.numpy(), h_xe.numpy())
    mark("load (H2D)")
    for t in range(sweeps):
      ops.sweep_p2p(t)
    mark("%d sweeps" % sweeps)
    ops.store(sweeps, h_xn.numpy(), h_xe.numpy())
    mark("store (D2H)")
    ops.close()
    mark("close")
    history.append([(n, (t - marks[i][1]) * 1e3) for i, (n, t) in enumerate(marks[1:])])
    if rep >= 2 and reps <= 4:
      line = "rank %d rep %d: " % (rank, rep) + ", ".join(
          "%s %.2f ms" % (n, (t - marks[i][1]) * 1e3) for i, (n, t) in enumerate(marks[1:])) + \
          " | total %.2f ms" % ((marks[-1][1] - marks[0][1]) * 1e3)
      lines = [None] * world
      dist.all_gather_object(lines, line)
      if rank == 0:
        print("\n".join(lines), flush=True)
  if reps > 4:
    # many repetitions: the median step and every step that took 1.5 x the median, phase by phase
    totals = np.asarray([sum(v for _, v in h) for h in history[2:]])
    med = float(np.median(totals))
    lines = ["rank %d: median step %.2f ms over %d reps" % (rank, med, len(totals))]
    for k, h in enumerate(history[2:]):
      if totals[k] > 1.5 * med:
        lines.append("rank %d rep %d: %.1f ms = " % (rank, k + 2, totals[k]) +
                     ", ".join("%s %.2f" % (n, v) for n, v in h))
    every = [None] * world
    dist.all_gather_object(every, "\n".join(lines))
    if rank == 0:
      print("\n".join(every), flush=True)
  # the public path (ShardedRelaxation: the same steps plus the status agreements)
  for rep in range(3):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = hd.ShardedRelaxation(None, R, sweeps, ctx=ctx, shape=A.shape, csr_host=[t.numpy() for t in h_csr])
    t1 = time.perf_counter()
    r.run(h_xn.numpy(), h_xe.numpy())
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    r.close()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if rank == 0:
      print("ShardedRelaxation rep %d: construct %.2f ms, run %.2f ms, close %.2f ms" %
            (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
  from bench import host_link_probe
  probe = host_link_probe(world, rank)
  if rank == 0:
    import json
    print("host link: " + json.dumps(probe), flush=True)
  hd.release_peer_arenas(dist)
  dist.destroy_process_group()


if __name__ == "__main__":
  main()
