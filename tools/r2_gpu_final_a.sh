#!/bin/bash
# final evidence, part A: GPU tests, ncu launch list and --set full capture of k_sweep
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2f_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2f_gputests.log
python bench.py --no-extras --steps 2 --warmup 3 > $O/r2f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2f_launches.csv \
  python bench.py --no-extras --steps 2 --warmup 3 > $O/r2f_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 8 -c 4 -o $O/r2f_prof -f \
  python bench.py --no-extras --steps 2 --warmup 3 > $O/r2f_ncu_full.log 2>&1
echo "full capture rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2f_smoke.log
