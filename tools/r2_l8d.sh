#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-l8d}
{
  for n in 48 64 96 160 200; do
    timeout 300 python tools/probe_gate_stream.py $n 5 4096
    timeout 300 python tools/probe_gate_stream.py $n 5 4096 --mode=global
  done
  timeout 300 python tools/probe_gate_stream.py 96 7 4096
  timeout 300 python tools/probe_gate_stream.py 96 7 4096 --mode=global
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
