#!/bin/bash
# pipelined (dynamic gather) vs unpipelined peer-memory sweep at N ranks: bench lines with the phase split
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-2}
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 5 --warmup 3 --no-extras > $O/r2d_n${N}_$name.json 2> $O/r2d_n${N}_$name.err
  echo "== $name rc=$?"
  python - <<PY
import json
for l in open("$O/r2d_n${N}_$name.json"):
  if l.startswith("{"):
    d = json.loads(l)
    ph = d["roofline"].get("sweep_phases_ms") or {}
    print("  ms_per_step %.3f sweep %.4f parity %s" % (d["ms_per_step"], d["roofline"]["ms_per_sweep"], d["parity"]["ok"]), {k: round(v, 4) for k, v in ph.items() if k != "note"})
PY
}
run s1 HGE_P2P_SLICES=1
run s4 HGE_P2P_SLICES=4
run s2 HGE_P2P_SLICES=2 HGE_DYN_PIECES_PER_WARP=4
