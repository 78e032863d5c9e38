#!/bin/bash
# Round-2 GPU call: the three A/B runs parked at the end of round 1 (NOUNROLL, NOUNROLL + HOSTGEO on the global-image
# set; PR_COMPACT on the resident set).  Results: gpurun_out/r2_ab_global.json, gpurun_out/r2_ab_resident.json.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_variants.json
timeout 300 python tools/ab_variants.py variants/libsdimb_base.so variants/libsdimb_nounroll.so variants/libsdimb_nounroll_hostgeo.so > gpurun_out/r2_ab1.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/r2_ab_global.json
AB_SET=resident timeout 300 python tools/ab_variants.py variants/libsdimb_base.so variants/libsdimb_prcompact.so variants/libsdimb_prcompact_nounroll.so variants/libsdimb_nounroll.so > gpurun_out/r2_ab2.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/r2_ab_resident.json
python tools/probe_breakdown.py 256 3 16384 > gpurun_out/r2_breakdown_base.txt 2>&1
python - <<'P'
import json
for f in ("r2_ab_global", "r2_ab_resident"):
    r = json.load(open(f"gpurun_out/{f}.json"))
    for k, v in r.items():
        print(f, k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else v)
P
cat gpurun_out/r2_breakdown_base.txt | tail -12
