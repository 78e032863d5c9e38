#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-l8c}
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 -k "lanes_tail" 2>&1 | tail -5 ) > gpurun_out/${T}_parity.log
tail -5 gpurun_out/${T}_parity.log
{
  timeout 300 python tools/probe_gate_stream.py 256 5 4096 --check
  timeout 300 python tools/probe_gate_stream.py 256 7 4096 --check
  timeout 300 python tools/probe_gate_stream.py 128 5 4096 --check --mode=global
  SDIMB_NO_TAIL8=1 timeout 300 python tools/probe_gate_stream.py 128 5 4096 --mode=global
  timeout 300 python tools/probe_gate_stream.py 128 5 4096
  timeout 300 python tools/probe_gate_stream.py 500 5 1024 --check
  SDIMB_NO_TAIL8=1 timeout 300 python tools/probe_gate_stream.py 500 5 1024
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
bash tools/r2_l8b.sh ${T} > /dev/null 2>&1
