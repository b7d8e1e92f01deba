#!/bin/bash
# pipelined peer-memory sweep at N ranks: tests, then the bench line per slice count
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-2}
[ "${SKIP_TESTS:-0}" = 1 ] || timeout 900 python -m pytest tests/test_distributed_gpu.py -x -q > $O/r2c_gputests_n$N.log 2>&1
echo "pytest rc=$?"; tail -5 $O/r2c_gputests_n$N.log
for s in 1 2 4; do
  HGE_P2P_SLICES=$s timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + s)) bench.py --gpus $N --steps 10 --warmup 3 --no-extras > $O/r2c_bench_n${N}_s$s.json 2> $O/r2c_bench_n${N}_s$s.err
  echo "slices $s rc=$?"
  python - <<PY
import json
for l in open("$O/r2c_bench_n${N}_s$s.json"):
  if l.startswith("{"):
    d = json.loads(l)
    print("  ms_per_step %.3f sweep %.4f parity %s phases %s" % (d["ms_per_step"], d["roofline"]["ms_per_sweep"], d["parity"]["ok"], d["roofline"].get("sweep_phases_ms")))
PY
done
