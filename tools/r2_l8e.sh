#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-l8e}
( timeout 1400 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | grep -E "passed|failed|Error|assert" | tail -8 ) > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
{
  timeout 300 python tools/probe_gate_stream.py 256 5 4096 --check
  timeout 300 python tools/probe_gate_stream.py 256 7 16384 --check
  timeout 300 python tools/probe_gate_stream.py 128 5 4096 --check
  timeout 300 python tools/probe_gate_stream.py 500 5 1024 --check
} 2>&1 | grep -v Warning > gpurun_out/${T}_probe.txt
cat gpurun_out/${T}_probe.txt
