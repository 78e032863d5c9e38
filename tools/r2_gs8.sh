#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-gs9}
( timeout 1400 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | grep -E "passed|failed|Error" | tail -5 ) > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-configs --no-gate-update > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'P'
import json,sys
d=json.loads(open(sys.argv[1] if len(sys.argv)>1 else "gpurun_out/%s_bench.json" % "T").read().strip().splitlines()[-1]) if False else None
P
tail -c 1500 gpurun_out/${T}_bench.json | head -c 100
python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['records_match_device_path'],'ms',d['ms_per_step'])"
