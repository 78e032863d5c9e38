#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_variants.json
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail_run" ) > gpurun_out/gm4_pytest_gm.log 2>&1
tail -3 gpurun_out/gm4_pytest_gm.log
timeout 900 python tools/ab_variants.py variants/libsdimb_v0.so variants/libsdimb_t1.so variants/libsdimb_t2.so variants/libsdimb_t3.so variants/libsdimb_t4.so variants/libsdimb_t5.so variants/libsdimb_t6.so > gpurun_out/gm4_ab.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/gm4_ab.json
SDIMB_RUN_LPS=32 timeout 300 python tools/ab_variants.py variants/libsdimb_t1.so variants/libsdimb_t3.so > gpurun_out/gm4_ab32.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/gm4_ab_lps32.json
python - <<'P'
import json
for f in ("gm4_ab.json", "gm4_ab_lps32.json"):
    r = json.load(open("gpurun_out/" + f))
    for k, v in r.items():
        print(f, k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else str(v)[:600])
P
