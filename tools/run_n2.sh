#!/bin/bash
# dev helper: 2-GPU bench under torchrun (run through gpurun --gpus 2)
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/n$N.err | tail -1 > gpurun_out/bench_r1_n$N.json
python - <<PY
import json
j = json.load(open("gpurun_out/bench_r1_n$N.json"))
print(j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["value"], j["gpu_launches"], j["clocks"])
PY
tail -3 gpurun_out/n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-160
