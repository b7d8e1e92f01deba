"""Where the host-buffer (e2e) step of bench.py spends its time on config 2 (1 GPU)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from bench import WORKLOADS, build_workload  # noqa: E402
from hypergraphembedding_b200 import _native, synthetic  # noqa: E402


def main():
  spec = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
  A, B = build_workload(spec)
  N, E = A.shape
  R, sweeps = spec["R"], spec["sweeps"]
  ctx = _native.default_context(0)
  xn0, xe0 = synthetic.legacy_initial_vectors(N, E, R, seed=0)
  pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
  h = [pin(A.indptr.astype(np.int64)), pin(A.indices.astype(np.int32)),
       pin(B.indptr.astype(np.int64)), pin(B.indices.astype(np.int32))]
  h_xn, h_xe = pin(xn0), pin(xe0)
  d = torch.empty(h_xn.numel() + h_xe.numel(), dtype=torch.float32, device="cuda")
  for name, fn in (("h2d 192 MB", lambda: (d[:h_xn.numel()].copy_(h_xn.view(-1), non_blocking=True),
                                           d[h_xn.numel():].copy_(h_xe.view(-1), non_blocking=True))),
                   ("d2h 192 MB", lambda: (h_xn.view(-1).copy_(d[:h_xn.numel()], non_blocking=True),
                                           h_xe.view(-1).copy_(d[h_xn.numel():], non_blocking=True)))):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    print("%-28s %.2f ms" % (name, (time.perf_counter() - t) * 1e3))
  for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    inc = _native.Incidence(ctx, N, E, *[x.numpy() for x in h])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    _native.algdist_run(ctx, inc, h_xn.numpy(), h_xe.numpy(), sweeps)
    t3 = time.perf_counter()
    inc.close()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print("rep %d: incidence create %.2f ms (+%.2f to drain), run(host) %.2f ms, close %.2f ms, total %.2f ms"
          % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3))


if __name__ == "__main__":
  main()
