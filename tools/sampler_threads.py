"""Edge-edge FOBE candidate sampling on the 1M-node config-4 family graph vs worker threads."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypergraphembedding_b200 import _native, synthetic  # noqa: E402
from hypergraphembedding_b200.hg2v_sample import _Graph  # noqa: E402

lib = _native.load_library()
A = synthetic.zipf_hypergraph(1000000, 500000, seed=2024)
gr = _Graph.from_csr(A)
for th in [int(v) for v in sys.argv[1:]] or [1, 2, 4, 8]:
  lib.hge_sampler_set_threads(th)
  st = _native.LegacyRngState()
  t = time.time()
  r, c = _native.sample_adj_rows((gr.b, gr.bt), gr.edge_rows, np.full(len(gr.edge_rows), 200, np.int32), st)
  print("threads %d: %.2f s, %d samples" % (th, time.time() - t, len(r)), flush=True)
