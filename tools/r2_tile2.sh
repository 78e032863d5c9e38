#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/probe_small_breakdown.py planes > gpurun_out/t2_breakdown_tile.txt 2>&1
python tools/probe_small_breakdown.py planes-warp > gpurun_out/t2_breakdown_warp.txt 2>&1
cat gpurun_out/t2_breakdown_tile.txt gpurun_out/t2_breakdown_warp.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interp_tile -s 2 -c 1 -o gpurun_out/t2_config4_tile \
    python tools/probe_small_breakdown.py planes "config 4" > gpurun_out/t2_ncu.log 2>&1
tail -3 gpurun_out/t2_ncu.log
