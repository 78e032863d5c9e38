#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_variants.json
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail_run" ) > gpurun_out/gm3_pytest_gm.log 2>&1
tail -3 gpurun_out/gm3_pytest_gm.log
timeout 900 python tools/ab_variants.py variants/libsdimb_v0.so variants/libsdimb_v1.so variants/libsdimb_v2.so variants/libsdimb_v3.so variants/libsdimb_v4.so variants/libsdimb_v5.so variants/libsdimb_v6.so variants/libsdimb_v7.so > gpurun_out/gm3_ab.log 2>&1
python - <<'P'
import json
r = json.load(open("gpurun_out/ab_variants.json"))
for k, v in r.items():
    print(k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else str(v)[:600])
P
cp gpurun_out/ab_variants.json gpurun_out/gm3_ab.json
