#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python benchmarks/sweep.py > gpurun_out/r2_sweep.json 2> gpurun_out/r2_sweep.err
tail -3 gpurun_out/r2_sweep.err
