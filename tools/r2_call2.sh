#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_a.log 2>&1
tail -3 gpurun_out/r2_pytest_a.log
( time timeout 600 python bench.py ) > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 600 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
rm -f gpurun_out/ab_variants.json
timeout 600 python tools/ab_variants.py variants/libsdimb_w4c8.so variants/libsdimb_w2c16.so variants/libsdimb_w1c32.so variants/libsdimb_w2c12.so variants/libsdimb_w1c24.so variants/libsdimb_w8c4.so > gpurun_out/r2_ab3.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/r2_ab_warps.json
python - <<'P'
import json
r = json.load(open("gpurun_out/r2_ab_warps.json"))
for k, v in r.items():
    print(k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else str(v)[:600])
P
python tools/probe_meas_cost.py 256 3 16384 > gpurun_out/r2_meas_cost.txt 2>&1; cat gpurun_out/r2_meas_cost.txt
