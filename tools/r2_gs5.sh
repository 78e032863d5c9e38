#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs6}
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "gate_stream or tail_run" 2>&1 | tail -3 ) > gpurun_out/${T}_parity.log
cat gpurun_out/${T}_parity.log
{
  python tools/probe_gate_stream.py 256 3 16384 --check
  SDIMB_GS_NO_FUSE=1 python tools/probe_gate_stream.py 256 3 16384 --check
  python tools/probe_gate_stream.py 128 3 16384 --check
  SDIMB_GS_NO_FUSE=1 python tools/probe_gate_stream.py 128 3 16384 --check
  python tools/probe_gate_stream.py 400 2 8192 --check
  SDIMB_GS_NO_FUSE=1 python tools/probe_gate_stream.py 400 2 8192 --check
  python tools/probe_gate_stream.py 80 2 16384 --check
  SDIMB_GS_NO_FUSE=1 python tools/probe_gate_stream.py 80 2 16384 --check
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
