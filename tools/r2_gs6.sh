#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs7}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gate_stream' -s 1 -c 1 -o gpurun_out/${T}_front \
    python tools/run_case.py 256 3 16384 auto headline 2 > gpurun_out/${T}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'run_tail' -s 1 -c 1 -o gpurun_out/${T}_tail \
    python tools/run_case.py 256 3 16384 auto headline 2 >> gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
