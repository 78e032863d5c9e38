#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_program.py tests/test_gpu_frames.py tests/test_integration_stub.py -x -q --timeout 300 2>&1 | tail -15 ) > gpurun_out/e2_parity.log
cat gpurun_out/e2_parity.log
