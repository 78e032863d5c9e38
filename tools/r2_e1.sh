#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "edge_cases" 2>&1 | tail -15 ) > gpurun_out/e1_parity.log
cat gpurun_out/e1_parity.log
