#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs4}
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "gate_stream" 2>&1 | tail -3 ) > gpurun_out/${T}_parity.log
cat gpurun_out/${T}_parity.log
{
  python tools/probe_gate_stream.py 256 3 16384 --check
  SDIMB_GS_WARPS=4 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_GS_WARPS=6 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_GS_GLOBAL=1 python tools/probe_gate_stream.py 256 3 16384
  python tools/probe_gate_stream.py 160 3 16384 --check
  python tools/probe_gate_stream.py 64 3 16384 --check
  python tools/probe_gate_stream.py 64 3 16384 --check --mode=planes-global
  python tools/probe_gate_stream.py 128 3 16384 --check
  python tools/probe_gate_stream.py 128 3 16384 --check --mode=planes-global
  python tools/probe_gate_stream.py 100 2 16384 --check
  python tools/probe_gate_stream.py 100 2 16384 --check --mode=planes-global
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gate_stream' -s 1 -c 1 -o gpurun_out/${T}_front \
    python tools/run_case.py 256 3 16384 auto headline 2 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
