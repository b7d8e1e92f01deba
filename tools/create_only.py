"""Incidence creation alone (host buffers), for an ncu launch list of the set-up kernels."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from bench import WORKLOADS, build_workload  # noqa: E402
from hypergraphembedding_b200 import _native  # noqa: E402

spec = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
A, B = build_workload(spec)
N, E = A.shape
ctx = _native.default_context(0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h = [pin(A.indptr.astype(np.int64)), pin(A.indices.astype(np.int32)),
     pin(B.indptr.astype(np.int64)), pin(B.indices.astype(np.int32))]
for rep in range(3):
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  inc = _native.Incidence(ctx, N, E, *[x.numpy() for x in h])
  t1 = time.perf_counter()
  torch.cuda.synchronize()
  t2 = time.perf_counter()
  print("create: host %.2f ms, drained after %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
  inc.close()
