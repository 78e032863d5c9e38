#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_program.py tests/test_gpu_frames.py -x -q --timeout 100 2>&1 | tail -6 ) > gpurun_out/t3_parity.log
tail -2 gpurun_out/t3_parity.log
true
true
