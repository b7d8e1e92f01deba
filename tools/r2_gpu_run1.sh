#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, the bench line, the ncu launch list and one
# --set full capture of the half-sweep kernel, compute-sanitizer over the multi-chunk / tile tests.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2b_gputests.log 2>&1; echo "pytest rc=$?"
tail -3 $O/r2b_gputests.log
python bench.py > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?"
python bench.py --no-extras --steps 2 --warmup 3 > $O/r2b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2b_launches.csv \
  python bench.py --no-extras --steps 2 --warmup 3 > $O/r2b_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 8 -c 4 -o $O/r2b_prof -f \
  python bench.py --no-extras --steps 2 --warmup 3 > $O/r2b_ncu_full.log 2>&1
echo "full capture rc=$?"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest tests -m gpu -x -q -k "long_rows or tiles or one_rank or chunk" > $O/r2b_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"
  tail -5 $O/r2b_sanitizer_$tool.log
done
