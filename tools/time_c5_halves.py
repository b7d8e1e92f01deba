"""Per-half-sweep timing of config 5 (or c5mini) on one GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from bench import WORKLOADS, community_shard_device  # noqa: E402
from hypergraphembedding_b200 import _native  # noqa: E402

spec = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c5"]
dev = torch.device("cuda", 0)
ctx = _native.default_context(0)
if len(sys.argv) >= 5:
  ctx.set_tuning(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
g = community_shard_device(spec, 0, 1, dev)
R, sweeps = spec["R"], 4
inc = _native.Incidence(ctx, g["n_loc"], g["E"], g["a_ptr"], g["a_idx"], g["b_ptr"], g["b_idx"])
xn = torch.rand((g["n_loc"], R), device=dev)
xe = torch.rand((g["E"], R), device=dev)
st = _native.AlgDistState(ctx, inc, R, sweeps)
st.load(xn, xe)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * sweeps + 1)]
evs[0].record()
for t in range(sweeps):
  st.node_half(t)
  evs[2 * t + 1].record()
  st.edge_half(t)
  evs[2 * t + 2].record()
torch.cuda.synchronize()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(2 * sweeps)]
nnz, N, E = g["nnz"], g["n_loc"], g["E"]
node_b = nnz * (4 * R + 4) + 2 * N * 4 * R
edge_b = nnz * (4 * R + 4) + 2 * E * 4 * R
print("node half %.2f ms (%.0f GB/s algorithmic), edge half %.2f ms (%.0f GB/s)" % (
    np.mean(ms[2::2]), node_b / np.mean(ms[2::2]) / 1e6, np.mean(ms[3::2]), edge_b / np.mean(ms[3::2]) / 1e6))
deg_e = (g["b_ptr"][1:] - g["b_ptr"][:-1]).cpu().numpy()
deg_n = (g["a_ptr"][1:] - g["a_ptr"][:-1]).cpu().numpy()
print("edge sizes: max %d, >64: %d rows holding %.1f%% of nnz; node degrees: mean %.1f max %d, >64: %d" % (
    deg_e.max(), (deg_e > 64).sum(), 100.0 * deg_e[deg_e > 64].sum() / nnz, deg_n.mean(), deg_n.max(),
    (deg_n > 64).sum()))
