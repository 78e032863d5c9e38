#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs5}
( timeout 1400 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -12 ) > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
{
  python tools/probe_gate_stream.py 256 3 16384 --check
  SDIMB_GS_WARPS=4 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_GS_WARPS=6 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_GS_GLOBAL=1 python tools/probe_gate_stream.py 256 3 16384
  python tools/probe_gate_stream.py 128 3 16384 --check
  python tools/probe_gate_stream.py 96 3 16384 --check
  python tools/probe_gate_stream.py 96 3 16384 --check --mode=planes
  python tools/probe_gate_stream.py 80 3 16384 --check
  python tools/probe_gate_stream.py 80 3 16384 --check --mode=planes
  python tools/probe_gate_stream.py 64 3 16384 --check --mode=planes-global
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
