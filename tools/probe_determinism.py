import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdim_b200.engine import TableauEngine, simulate_host
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import noisy_random_clifford
from oracle import c_oracle
shots = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
prog = compile_circuits([noisy_random_clifford(256, 2000, 3)])
eng = TableauEngine(prog)
tab = eng.alloc_tableau(shots)
a = eng.run(shots, 0, 2026, tableau=tab).cpu().numpy()
b = eng.run(shots, 0, 2026, tableau=tab).cpu().numpy()
h, _ = simulate_host(prog, shots, 0, 2026)
want = c_oracle.run_philox(prog, shots, 0, 2026)
for name, x in (("run1", a), ("run2", b), ("host", h)):
    bad = np.argwhere(x != want)
    print(name, "mismatching records:", len(bad), "shots:", len(set(bad[:, 0])) if len(bad) else 0,
          "first:", bad[:5].tolist())
    if len(bad):
        s, k = bad[0]
        print("   got", x[s, max(0,k-2):k+3], "want", want[s, max(0,k-2):k+3])
