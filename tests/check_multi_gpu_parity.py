"""Parity of the sharded relaxation at N ranks (one per visible GPU) against the oracle, for
rank counts the pytest suite cannot assume (it runs 2-rank cases when 2 GPUs are visible).  A
script, not a collected test:

    python tests/check_multi_gpu_parity.py [N] [p2p|nccl] [R]
"""
import os
import socket
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.multiprocessing as mp  # noqa: E402

import dist_helpers  # noqa: E402
from oracle import port  # noqa: E402


def main():
  world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
  comm = sys.argv[2] if len(sys.argv) > 2 else "p2p"
  R = int(sys.argv[3]) if len(sys.argv) > 3 else 32
  graph_args, iters = (21, 80000, 3001, 900000), 10
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  free_port = s.getsockname()[1]
  s.close()
  with tempfile.TemporaryDirectory() as tmp:
    mp.spawn(dist_helpers.worker,
             args=(world, free_port, "nccl", graph_args, R, iters, 1, tmp, True, comm),
             nprocs=world, join=True)
    A = dist_helpers.make_graph(*graph_args)
    rng = np.random.default_rng(123)
    xn0 = rng.random((A.shape[0], R)).astype(np.float32)
    xe0 = rng.random((A.shape[1], R)).astype(np.float32)
    ref_xn, ref_xe = port.algdist_vectorised(A, A.T.tocsr(), xn0, xe0, iters)
    got = np.zeros_like(ref_xn)
    worst_e = 0.0
    for r in range(world):
      z = np.load(os.path.join(tmp, "rank%d.npz" % r))
      got[int(z["r0"]):int(z["r1"])] = z["xn"]
      worst_e = max(worst_e, float(np.abs(z["xe"] - ref_xe).max()))
    worst_n = float(np.abs(got - ref_xn).max())
  ok = worst_n < 2e-5 and worst_e < 2e-5
  print("check_multi world=%d comm=%s R=%d: max |xn - oracle| = %.3g, max |xe - oracle| = %.3g -> %s"
        % (world, comm, R, worst_n, worst_e, "OK" if ok else "FAIL"))
  sys.exit(0 if ok else 1)


if __name__ == "__main__":
  main()
