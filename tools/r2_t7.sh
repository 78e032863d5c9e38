#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-t7}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:interp_tile -s 2 -c 1 -o gpurun_out/${T}_config4_tile \
    python tools/probe_small_breakdown.py planes "config 4" > gpurun_out/${T}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:interp_tile -s 2 -c 1 -o gpurun_out/${T}_config3_tile \
    python tools/probe_small_breakdown.py planes "config 3" >> gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
