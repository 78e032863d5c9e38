#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_b.log 2>&1
tail -3 gpurun_out/r2_pytest_b.log
rm -f gpurun_out/ab_variants.json
timeout 600 python tools/ab_variants.py variants/libsdimb_w4c8.so variants/libsdimb_new_all.so variants/libsdimb_no_il.so variants/libsdimb_no_detrun.so variants/libsdimb_no_merge.so variants/libsdimb_none.so > gpurun_out/r2_ab4.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/r2_ab_meas.json
python - <<'P'
import json
r = json.load(open("gpurun_out/r2_ab_meas.json"))
for k, v in r.items():
    print(k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else str(v)[:600])
P
python tools/probe_meas_cost.py 256 3 16384 > gpurun_out/r2_meas_cost_b.txt 2>&1; cat gpurun_out/r2_meas_cost_b.txt
