#!/bin/bash
# A/B of the hot-row variant of the node half on config 2 (run on the GPU box):
#   bash tools/r2_hot_ab.sh > gpurun_out/r2_hot_ab.log 2>&1
# Every line is bench.py's own JSON (full-size parity against the C port included), reduced to
# the keys that matter here.
set -u
cd "$(dirname "$0")/.."
pick='import json,sys
for l in sys.stdin:
  if l.startswith("{"):
    d=json.loads(l); r=d["roofline"]
    print(json.dumps({"ms_per_step": d["ms_per_step"], "node_half_ms": r["node_half_ms"], "edge_half_ms": r["edge_half_ms"], "parity_ok": d["parity"]["ok"], "max_dist_err_over_bound": d["parity"]["max_dist_err_over_bound"], "e2e_ms": d["e2e"]["ms_per_step"]}))'
run() { # name, env...
  local name=$1; shift
  echo "== $name"
  env "$@" timeout 600 python bench.py --no-extras --steps 5 --warmup 3 2>/tmp/hot_ab.err | python -c "$pick" || tail -5 /tmp/hot_ab.err
}
python -m pytest tests/test_algdist_gpu.py -x -q -m gpu 2>&1 | tail -3
run plain HGE_HOT_ROWS=0
run hot_auto X=1
run hot_512 HGE_HOT_ROWS=512
run hot_auto_waves1 HGE_SWEEP_WAVES=1
run hot_auto_waves3 HGE_SWEEP_WAVES=3
run plain_waves3 HGE_HOT_ROWS=0 HGE_SWEEP_WAVES=3
