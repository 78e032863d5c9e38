#!/bin/bash
# One GPU call: A/B the library variants (tools/build_variants.sh), run the parity tests and bench.py on the winner.
# usage: tools/ab_call.sh name1 name2 ...   (variants/libsdimb_<name>.so); everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_variants.json
libs=""
for v in "$@"; do libs="$libs variants/libsdimb_$v.so"; done
timeout 240 python tools/ab_variants.py $libs > gpurun_out/ab1.log 2>&1
best=$(python tools/ab_pick.py)
echo "best=$best" | tee gpurun_out/ab_best.txt
if [ -n "$best" ]; then
  SDIMB_LIB=$best timeout ${AB_PYTEST_TIMEOUT:-150} python -m pytest tests/test_gpu_parity.py -x -q -m gpu ${AB_PYTEST_ARGS} > gpurun_out/ab_pytest.log 2>&1
  echo "pytest rc=$?" | tee -a gpurun_out/ab_best.txt
  SDIMB_LIB=$best timeout 120 python bench.py --steps 5 --warmup 3 > gpurun_out/ab_bench_best.json 2> gpurun_out/ab_bench_best.err
fi
tail -3 gpurun_out/ab_pytest.log
python - <<'P'
import json
r = json.load(open("gpurun_out/ab_variants.json"))
for k, v in r.items():
    print(k, [(x["d"], x["n"], round(x["ms_min"], 2), x["records_equal_oracle"]) for x in v] if isinstance(v, list) else v)
P
