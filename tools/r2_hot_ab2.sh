#!/bin/bash
# second A/B of the hot-row variant: how many rows, generic vs predicated loads
set -u
cd "$(dirname "$0")/.."
V=hypergraphembedding_b200/_variants
run() { local name=$1; shift; echo "== $name"; env "$@" timeout 600 python tools/time_variant.py c2 2>&1 | tail -1; }
run plain HGE_HOT_ROWS=0
for n in 8 128 512 -1; do run generic_$n HGE_HOT_ROWS=$n; done
for n in 8 128 512 -1; do run pred_$n HGE_HOT_ROWS=$n HGE_LIB_PATH=$V/libhge_hotpred.so; done
