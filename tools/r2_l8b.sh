#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-l8b}
cat > /tmp/run_d5.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import noisy_random_clifford
prog = compile_circuits([noisy_random_clifford(256, 2000, 5)])
eng = TableauEngine(prog)
rec = torch.empty((4096, prog.n_meas), dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.run(4096, 0, 1, records=rec)
torch.cuda.synchronize()
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'run_tail8' -s 1 -c 1 -o gpurun_out/${T}_tail8 python /tmp/run_d5.py > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
