#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_variants.json
timeout 900 python tools/ab_variants.py variants/libsdimb_base.so variants/libsdimb_v_none.so variants/libsdimb_v_il.so variants/libsdimb_v_merge.so variants/libsdimb_v_nodetrun.so variants/libsdimb_v_none_unroll.so variants/libsdimb_v_nodetrun_unroll.so > gpurun_out/r2_ab6.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/r2_ab_meas3.json
python - <<'P'
import json
r = json.load(open("gpurun_out/r2_ab_meas3.json"))
for k, v in r.items():
    print(k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else str(v)[:600])
P
