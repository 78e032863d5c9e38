#!/bin/bash
# ncu --set full of run_tail_kernel (and the interpreter in front of it) on the headline workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=${1:-}
[ -n "$L" ] && export SDIMB_LIB=$PWD/$L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_tail -s 1 -c 1 -f -o gpurun_out/r2_tail_kernel \
    python tools/run_case.py 256 3 16384 auto headline 2 > gpurun_out/r2_ncu_tail.log 2>&1
tail -2 gpurun_out/r2_ncu_tail.log
ls -la gpurun_out/r2_tail_kernel.ncu-rep
