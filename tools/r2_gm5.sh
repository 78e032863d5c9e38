#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_variants.json
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail_run" ) > gpurun_out/gm5_pytest_gm.log 2>&1
tail -3 gpurun_out/gm5_pytest_gm.log
timeout 900 python tools/ab_variants.py variants/libsdimb_g8.so variants/libsdimb_g10.so variants/libsdimb_g12.so variants/libsdimb_g6.so > gpurun_out/gm5_ab.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/gm5_ab.json
SDIMB_NO_GATES_ONLY=1 timeout 300 python tools/ab_variants.py variants/libsdimb_g8.so > gpurun_out/gm5_ab_nogo.log 2>&1
mv gpurun_out/ab_variants.json gpurun_out/gm5_ab_nogo.json
python - <<'P'
import json
for f in ("gm5_ab.json", "gm5_ab_nogo.json"):
    r = json.load(open("gpurun_out/" + f))
    for k, v in r.items():
        print(f, k, [(x["d"], x["n"], round(x["ms_min"], 3), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else str(v)[:600])
P
