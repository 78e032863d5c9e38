#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in 16384 32768 65536; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-configs --no-gate-update --shots $s > gpurun_out/b64_$s.json 2> gpurun_out/b64_$s.err
python -c "
import json
d=json.loads(open('gpurun_out/b64_$s.json').read().strip().splitlines()[-1])
print($s, 'value %.4e e2e %.4e'%(d['value'],d['e2e']['value']),'ms',round(d['ms_per_step'],3), d['e2e']['records_match_device_path'], d['roofline']['step_kernels']['headline_front']['launch_ms'], d['roofline']['step_kernels']['headline_tail']['launch_ms'])"
done
