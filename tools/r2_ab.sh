#!/bin/bash
# Round-2 A/B of the half-sweep kernel variants on config 2 (run on the GPU box):
#   bash tools/r2_ab.sh > gpurun_out/r2_ab.log 2>&1
set -u
cd "$(dirname "$0")/.."
V=hypergraphembedding_b200/_variants
run() { # name, env..., -- tuning args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "== $name"
  env "${envs[@]}" timeout 600 python tools/time_variant.py c2 "$@" 2>&1 | tail -1
}
run default X=1 --
run default_again X=1 --
for f in $V/libhge_*.so; do
  [ -e "$f" ] || continue
  run "$(basename $f)" HGE_LIB_PATH=$f --
done
