"""Dev probe: time variants of the headline circuit to see which op class dominates."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200 import Circuit, generate_random_clifford_circuit
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import noisy_random_clifford

def timeit(prog, shots, mode=None, reps=3):
    eng = TableauEngine(prog)
    tab = eng.alloc_tableau(shots) if eng.plan(mode)[1] else None
    rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.run(shots, 0, 1, mode=mode, tableau=tab, records=rec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.run(shots, 0, 1, mode=mode, tableau=tab, records=rec)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n, d, g = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 3, 2000
shots = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
mode = sys.argv[4] if len(sys.argv) > 4 else None
variants = {
    "gates_only": generate_random_clifford_circuit(n, g, d, measurement_rounds=0, seed=1),
    "gates+meas": generate_random_clifford_circuit(n, g, d, measurement_rounds=1, seed=1),
    "noisy_no_meas": noisy_random_clifford(n, g, d, measurement_rounds=0),
    "headline": noisy_random_clifford(n, g, d),
    "empty": Circuit(n, d),
}
for name, c in variants.items():
    prog = compile_circuits([c])
    ms = timeit(prog, shots, mode)
    print(f"{name:14s} ops={prog.n_ops:5d} meas={prog.n_meas:4d} shots={shots} {ms:9.3f} ms  "
          f"{shots * max(prog.n_user_gates,1) / ms * 1e3:10.3e} shot*gates/s")
