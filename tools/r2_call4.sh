#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_variants.json
timeout 900 python tools/ab_variants.py variants/libsdimb_base.so variants/libsdimb_u_none.so variants/libsdimb_u_il.so variants/libsdimb_u_merge.so variants/libsdimb_u_nodetrun.so variants/libsdimb_u_detrun_only.so variants/libsdimb_u_all.so > gpurun_out/r2_ab5.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/r2_ab_meas2.json
python - <<'P'
import json
r = json.load(open("gpurun_out/r2_ab_meas2.json"))
for k, v in r.items():
    print(k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else str(v)[:600])
P
for v in u_nodetrun u_all; do echo "== $v"; SDIMB_LIB=$PWD/variants/libsdimb_$v.so python tools/probe_meas_cost.py 256 3 16384; done > gpurun_out/r2_meas_cost_c.txt 2>&1; cat gpurun_out/r2_meas_cost_c.txt
