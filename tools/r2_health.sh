#!/bin/bash
# Round-2 health check on the GPU box: the -m gpu suite, smoke(), both bench arms.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2h_pytest.log 2>&1
tail -3 gpurun_out/r2h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2h_smoke.log 2>&1; tail -1 gpurun_out/r2h_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -1 gpurun_out/r2h_bench.json | cut -c1-1500
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
tail -1 gpurun_out/r2h_bench_ref.json | cut -c1-600
