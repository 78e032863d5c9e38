#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 120 -k "host_entry or host" 2>&1 | tail -4 ) > gpurun_out/h2_parity.log
cat gpurun_out/h2_parity.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-configs --no-gate-update > gpurun_out/h2_bench.json 2> gpurun_out/h2_bench.err
tail -2 gpurun_out/h2_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/h2_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e'],'ms',d['ms_per_step'])"
