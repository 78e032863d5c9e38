#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs1}
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "gate_stream or tail_run" 2>&1 | tail -6 ) > gpurun_out/${T}_parity.log
cat gpurun_out/${T}_parity.log
{
  SDIMB_NO_GATE_STREAM=1 python tools/probe_gate_stream.py 256 3 16384 --check
  python tools/probe_gate_stream.py 256 3 16384 --check
  SDIMB_GS_WARPS=4 python tools/probe_gate_stream.py 256 3 16384 --check
  SDIMB_GS_WARPS=6 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_GS_WARPS=3 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_GS_GLOBAL=1 python tools/probe_gate_stream.py 256 3 16384 --check
  SDIMB_NO_GATE_STREAM=1 python tools/probe_gate_stream.py 400 2 8192
  python tools/probe_gate_stream.py 400 2 8192 --check
  SDIMB_NO_GATE_STREAM=1 python tools/probe_gate_stream.py 500 3 4096
  python tools/probe_gate_stream.py 500 3 4096 --check
  SDIMB_NO_GATE_STREAM=1 python tools/probe_gate_stream.py 160 3 16384
  python tools/probe_gate_stream.py 160 3 16384 --check
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
