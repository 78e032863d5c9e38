#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-l8f}
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 120 -k "lanes_tail or edge_cases" 2>&1 | tail -8 ) > gpurun_out/${T}_parity.log
tail -8 gpurun_out/${T}_parity.log
{
  timeout 120 python tools/probe_gate_stream.py 256 5 4096 --check
  SDIMB_TAIL8_NO_TMA=1 timeout 120 python tools/probe_gate_stream.py 256 5 4096 --check
  timeout 120 python tools/probe_gate_stream.py 128 5 4096 --check
  SDIMB_TAIL8_NO_TMA=1 timeout 120 python tools/probe_gate_stream.py 128 5 4096
  timeout 120 python tools/probe_gate_stream.py 500 5 1024 --check
  SDIMB_TAIL8_NO_TMA=1 timeout 120 python tools/probe_gate_stream.py 500 5 1024
  timeout 120 python tools/probe_gate_stream.py 256 7 16384 --check
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
