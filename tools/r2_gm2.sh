#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail_run" ) > gpurun_out/gm2_pytest_gm.log 2>&1
tail -5 gpurun_out/gm2_pytest_gm.log
python tools/probe_breakdown.py 256 3 16384 > gpurun_out/gm2_breakdown.txt 2>&1; cat gpurun_out/gm2_breakdown.txt
SDIMB_GM_MIN_RUN=100000 python tools/probe_breakdown.py 256 3 16384 > gpurun_out/gm2_breakdown_nogm.txt 2>&1; cat gpurun_out/gm2_breakdown_nogm.txt
for c in 2 4 6; do echo "run ctas/sm $c"; SDIMB_RUN_CTAS_PER_SM=$c python tools/probe_breakdown.py 256 3 16384 2>&1 | grep -E "gates\+meas|headline"; done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/gm2_bench.json 2> gpurun_out/gm2_bench.err; tail -1 gpurun_out/gm2_bench.json | cut -c1-300
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/gm2_pytest.log 2>&1
tail -3 gpurun_out/gm2_pytest.log
