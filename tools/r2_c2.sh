#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-c2}
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 200 -k "tile" 2>&1 | tail -3 ) > gpurun_out/${T}_parity.log
cat gpurun_out/${T}_parity.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-gate-update > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'])
for c in d['configs']: print(c['name'][:40], c['kernel'], '%.3e'%c['value'], c['records_match_oracle'], round(c['ms_per_pass'],2))"
