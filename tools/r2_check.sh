#!/bin/bash
# full GPU suite + bench line (no CPU legs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-chk}
( timeout 900 python -m pytest tests -x -q -m gpu --timeout 200 2>&1 | tail -5 ) > gpurun_out/${T}_pytest.log
tail -2 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<P
import json
j = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print(j["value"], j["e2e"]["value"], j["roofline"]["kernel"], j["roofline"]["frac"], j["roofline"].get("frac_dram"))
print(json.dumps(j["roofline"].get("step_kernels"))[:600])
print([(c["name"][:8], c.get("kernel"), round(c["value"] / 1e9, 3), c.get("records_match_oracle")) for c in j["configs"]])
P
