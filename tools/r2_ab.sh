#!/bin/bash
# Round-2 A/B of the half-sweep kernels on config 2 (run on the GPU box):
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
run stream X=1 --
run items HGE_KERNEL=items --
for f in $V/libhge_*.so; do
  [ -e "$f" ] || continue
  run "$(basename $f)" HGE_LIB_PATH=$f --
done
run stream_uc1 HGE_UNIT_COST=1 --
run stream_uc6 HGE_UNIT_COST=6 --
run stream_2waves X=1 -- 128 1024 8
run stream_4waves X=1 -- 128 1024 16
run stream_l64 X=1 -- 64 1024 0
run stream_l255_c2048 X=1 -- 255 2048 0
