#!/bin/bash
# Round-2 evidence of the shipped binary (run on the GPU box through gpurun): GPU tests, bench line, ncu launch list of
# the bench command, full captures of the headline's kernels.  Summarised into profiles/ by tools/summarise_profiles.py r2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r2}
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${R}_pytest.log 2>&1
tail -3 gpurun_out/${R}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
tail -1 gpurun_out/${R}_bench_n1.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/${R}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'interp_planes|gate_stream' -s 1 -c 1 -o gpurun_out/${R}_headline_front \
    python tools/run_case.py 256 3 16384 auto headline 2 > gpurun_out/${R}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:run_tail -s 1 -c 1 -o gpurun_out/${R}_headline_tail \
    python tools/run_case.py 256 3 16384 auto headline 2 >> gpurun_out/${R}_ncu.log 2>&1
ls -la gpurun_out/${R}_*
( timeout 600 python __graft_entry__.py smoke ) > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
