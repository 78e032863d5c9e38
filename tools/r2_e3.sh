#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "lane_kernels_on_the_reference or lanes_tail" 2>&1 | tail -12 ) > gpurun_out/e3_parity.log
cat gpurun_out/e3_parity.log
