#!/bin/bash
# first GPU run of the tile interpreter: parity (goldens + Philox vs C oracle + config tests), then configs timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or philox or config or planes" 2>&1 | tail -15 ) > gpurun_out/t1_parity.log
tail -3 gpurun_out/t1_parity.log
( timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/t1_all.log
tail -2 gpurun_out/t1_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-gate-update > gpurun_out/t1_bench.json 2> gpurun_out/t1_bench.err
SDIMB_NO_TILE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-gate-update > gpurun_out/t1_bench_notile.json 2> gpurun_out/t1_bench_notile.err
python - <<'P'
import json
for f in ("gpurun_out/t1_bench.json", "gpurun_out/t1_bench_notile.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["value"], [(c["name"][:8], c.get("kernel"), round(c["value"] / 1e9, 2), c.get("records_match_oracle")) for c in j["configs"]])
    except Exception as e:
        print(f, "failed", e)
P
