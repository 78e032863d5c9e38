#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs8}
{
  python tools/probe_gate_stream.py 256 3 16384 --check
  for v in "$@"; do
    [ "$v" = "$T" ] && continue
    echo "variant $v"
    SDIMB_LIB=$PWD/variants/libsdimb_$v.so python tools/probe_gate_stream.py 256 3 16384 --check
    SDIMB_LIB=$PWD/variants/libsdimb_$v.so python tools/probe_gate_stream.py 400 2 8192
  done
  python tools/probe_gate_stream.py 400 2 8192
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
