#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest_c.log 2>&1
tail -3 gpurun_out/r2_pytest_c.log
rm -f gpurun_out/ab_variants.json
timeout 900 python tools/ab_variants.py variants/libsdimb_base.so variants/libsdimb_il.so variants/libsdimb_il_w8.so variants/libsdimb_il_w8c3.so variants/libsdimb_il_w6c5.so variants/libsdimb_il_nounroll.so > gpurun_out/r2_ab7.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/r2_ab_il.json
python - <<'P'
import json
r = json.load(open("gpurun_out/r2_ab_il.json"))
for k, v in r.items():
    print(k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else str(v)[:600])
P
