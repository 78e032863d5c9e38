"""Dev probe: run one configuration once or twice (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200 import generate_random_clifford_circuit
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import noisy_random_clifford
n, d, shots = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "auto" else None
kind = sys.argv[5] if len(sys.argv) > 5 else "headline"
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
c = {"headline": lambda: noisy_random_clifford(n, 2000, d),
     "gates": lambda: generate_random_clifford_circuit(n, 2000, d, measurement_rounds=0, seed=1),
     "gates+meas": lambda: generate_random_clifford_circuit(n, 2000, d, measurement_rounds=1, seed=1)}[kind]()
prog = compile_circuits([c])
eng = TableauEngine(prog)
tab = eng.alloc_tableau(shots) if eng.plan(mode)[1] else None
rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    eng.run(shots, 0, 1, mode=mode, tableau=tab, records=rec)
torch.cuda.synchronize()
print("done", n, d, shots, mode, kind)
