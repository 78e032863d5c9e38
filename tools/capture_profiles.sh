#!/bin/bash
# Run on the GPU box (through gpurun): ncu launch list of the bench command, full captures of the two
# headline kernels, sanitizer runs.  Outputs land in gpurun_out/ and are summarised into profiles/ by
# tools/summarise_profiles.py on the CPU box.
set +x
R=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --shots 16384 --no-cpu > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:interp -s 1 -c 1 -o gpurun_out/${R}_headline_planes \
    python tools/run_case.py 256 3 16384 auto headline 2 > gpurun_out/${R}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:interp -s 1 -c 1 -o gpurun_out/${R}_headline_lanes_global \
    python tools/run_case.py 256 3 4096 global headline 2 >> gpurun_out/${R}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:interp -s 1 -c 1 -o gpurun_out/${R}_gates_stream \
    python tools/run_case.py 256 3 9472 global gates 2 >> gpurun_out/${R}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cluster -s 1 -c 1 -o gpurun_out/${R}_config5_cluster \
    python tools/run_config5.py 4096 >> gpurun_out/${R}_ncu.log 2>&1
{
  echo "## racecheck planes d=3 n=256 (4 warps/shot)"; compute-sanitizer --tool racecheck --racecheck-report analysis python tools/probe_race.py 3 256 600 planes 2>&1 | tail -2
  echo "## racecheck lanes-resident d=5 n=70";          compute-sanitizer --tool racecheck --racecheck-report analysis python tools/probe_race.py 5 70 300 resident 2>&1 | tail -2
  echo "## racecheck lanes-resident d=2 n=97";          compute-sanitizer --tool racecheck --racecheck-report analysis python tools/probe_race.py 2 97 400 resident 2>&1 | tail -2
  echo "## memcheck planes d=3 n=256";                  compute-sanitizer --tool memcheck python tools/probe_race.py 3 256 600 planes 2>&1 | tail -1
  echo "## memcheck planes d=2 n=300";                  compute-sanitizer --tool memcheck python tools/probe_race.py 2 300 600 planes 2>&1 | tail -1
  echo "## memcheck lanes-global d=7 n=300";            compute-sanitizer --tool memcheck python tools/probe_race.py 7 300 300 global 2>&1 | tail -1
  echo "## memcheck cluster interpreter (8 CTAs) d=7 n=300";   compute-sanitizer --tool memcheck python tools/probe_race.py 7 300 300 cluster 2>&1 | tail -1
  echo "## memcheck cluster interpreter (16 CTAs) d=5 n=600";  SDIMB_CLUSTER_SIZE=16 compute-sanitizer --tool memcheck python tools/probe_race.py 5 600 300 cluster 2>&1 | tail -1
  echo "## memcheck lanes-global streaming shape (gates only) d=5 n=256"; compute-sanitizer --tool memcheck python tools/run_case.py 256 5 64 global gates 1 2>&1 | tail -1
} > gpurun_out/${R}_sanitizer.txt 2>&1
python bench.py --steps 5 --warmup 3 --shots 16384 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
tail -1 gpurun_out/${R}_bench_n1.json | cut -c1-300
python tools/probe_cluster.py 4096 5 1 auto,global-cta > gpurun_out/${R}_config5_breakdown.txt 2>&1
python benchmarks/configs.py > gpurun_out/${R}_configs.json 2> gpurun_out/${R}_configs.err
