#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-t6}
( SDIMB_TILE_GLB=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 200 -k "tile or golden or config" 2>&1 | tail -4 ) > gpurun_out/${T}_parity.log
cat gpurun_out/${T}_parity.log
{
  echo "shared-memory images"; timeout 300 python tools/probe_small_breakdown.py planes
  echo "global images, 20 CTAs/SM"; SDIMB_TILE_GLB=1 timeout 300 python tools/probe_small_breakdown.py planes
  echo "global images, 24 CTAs/SM"; SDIMB_LIB=$PWD/variants/libsdimb_glb24.so SDIMB_TILE_GLB=1 timeout 300 python tools/probe_small_breakdown.py planes
  echo "global images, 32 CTAs/SM"; SDIMB_LIB=$PWD/variants/libsdimb_glb32.so SDIMB_TILE_GLB=1 timeout 300 python tools/probe_small_breakdown.py planes
} 2>&1 | grep -v Warning > gpurun_out/${T}_breakdown.txt
cat gpurun_out/${T}_breakdown.txt
