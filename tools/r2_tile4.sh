#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-t4}
( timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --timeout 100 -k "golden or philox or config3 or config4 or config2" 2>&1 | tail -4 ) > gpurun_out/${T}_parity.log
tail -2 gpurun_out/${T}_parity.log
timeout 300 python tools/probe_small_breakdown.py planes > gpurun_out/${T}_breakdown_tile.txt 2>&1
cat gpurun_out/${T}_breakdown_tile.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:interp_tile -s 2 -c 1 -o gpurun_out/${T}_config4_tile \
    python tools/probe_small_breakdown.py planes "config 4" > gpurun_out/${T}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:interp_tile -s 2 -c 1 -o gpurun_out/${T}_config3_tile \
    python tools/probe_small_breakdown.py planes "config 3" >> gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
