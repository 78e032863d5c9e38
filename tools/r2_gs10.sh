#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs17}
{
  python tools/probe_gate_stream.py 256 3 16384 --check
  SDIMB_RUN_LPS=32 python tools/probe_gate_stream.py 256 3 16384 --check
  SDIMB_GS_WARPS=6 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_GS_WARPS=7 python tools/probe_gate_stream.py 256 3 16384
  python tools/probe_gate_stream.py 256 3 65536
  python tools/probe_gate_stream.py 256 2 16384 --check
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
