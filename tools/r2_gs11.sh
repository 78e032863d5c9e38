#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs21}
{
  for cfg in "500 3 4096" "512 2 4096" "320 3 8192" "192 3 16384" "440 3 4096"; do
    python tools/probe_gate_stream.py $cfg --check
    SDIMB_LIB=$PWD/variants/libsdimb_presmem.so python tools/probe_gate_stream.py $cfg
  done
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
