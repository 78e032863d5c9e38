#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs10}
{
  for s in 9472 16384 18944 28416 37888; do python tools/probe_gate_stream.py 256 3 $s; done
  SDIMB_RUN_CTAS_PER_SM=7 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_RUN_CTAS_PER_SM=6 python tools/probe_gate_stream.py 256 3 16384
  SDIMB_RUN_CTAS_PER_SM=7 python tools/probe_gate_stream.py 256 3 16576
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
