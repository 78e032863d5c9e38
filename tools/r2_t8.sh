#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-t8}
{
  echo "base"; SDIMB_TILE_GLB=1 timeout 300 python tools/probe_small_breakdown.py planes
  for v in "$@"; do
    [ "$v" = "$T" ] && continue
    echo "variant $v"; SDIMB_LIB=$PWD/variants/libsdimb_$v.so SDIMB_TILE_GLB=1 timeout 300 python tools/probe_small_breakdown.py planes
  done
} 2>&1 | grep -v Warning > gpurun_out/${T}_breakdown.txt
cat gpurun_out/${T}_breakdown.txt
