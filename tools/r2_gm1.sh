#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "generator_major" ) > gpurun_out/gm1_pytest_gm.log 2>&1
tail -5 gpurun_out/gm1_pytest_gm.log
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/gm1_pytest.log 2>&1
tail -3 gpurun_out/gm1_pytest.log
python tools/probe_breakdown.py 256 3 16384 > gpurun_out/gm1_breakdown.txt 2>&1; cat gpurun_out/gm1_breakdown.txt
SDIMB_GM_MIN_RUN=100000 python tools/probe_breakdown.py 256 3 16384 > gpurun_out/gm1_breakdown_nogm.txt 2>&1; cat gpurun_out/gm1_breakdown_nogm.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/gm1_bench.json 2> gpurun_out/gm1_bench.err; tail -1 gpurun_out/gm1_bench.json | cut -c1-400
