#!/bin/bash
# compute-sanitizer over the round-2 kernels -> gpurun_out/<tag>_sanitizer.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2e}
O=gpurun_out/${T}_sanitizer.txt
: > $O
run() {  # tool, title, env, args...
  tool=$1; title=$2; envs=$3; shift 3
  echo "## $tool $title" >> $O
  env $envs timeout 900 compute-sanitizer --tool $tool python tools/probe_race2.py "$@" 2>&1 | grep -E "^ran|SUMMARY|ERROR SUMMARY|Error|hazard" | grep -v "^=========     " | head -8 >> $O
}
run memcheck  "gate_stream_kernel (smem image, fused transposition) + run_tail_kernel, d=3 n=256" "X=1" 3 256 400 auto 0.0 5
run racecheck "gate_stream_kernel + run_tail_kernel, d=3 n=130" "X=1" 3 130 300 auto 0.0 3
run memcheck  "gate_stream_kernel + run_tail_kernel, d=2 n=300" "X=1" 2 300 400 auto 0.0 5
run memcheck  "gate_stream_kernel on the global image (n=500) + run_tail_kernel with its own transposition" "X=1" 3 500 300 auto 0.0 3
run memcheck  "lane interpreter + run_tail8_kernel, d=5 n=256" "X=1" 5 256 400 global 0.05 5
run racecheck "lane interpreter + run_tail8_kernel, d=7 n=100" "X=1" 7 100 300 global 0.05 3
run memcheck  "run_tail8_kernel with TMA staging, d=5 n=256" "SDIMB_TAIL8_TMA=1" 5 256 400 global 0.05 5
run memcheck  "tile interpreter, images in scratch, d=2 n=97" "SDIMB_TILE_GLB=1" 2 97 400 auto 0.1 9
run racecheck "tile interpreter, images in scratch, d=3 n=49" "SDIMB_TILE_GLB=1" 3 49 300 auto 0.1 9
cat $O
