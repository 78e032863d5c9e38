import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import qudit_repetition_code, rotated_surface_code, noisy_random_clifford
from sdim_b200 import generate_random_clifford_circuit
cases = {"rep": (qudit_repetition_code(25, 25, 3, prob=1e-2), 400000), "surface": (rotated_surface_code(7, 7, prob=1e-3), 400000),
         "cfg2": (generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1), 40000),
         "headline": (noisy_random_clifford(256, 2000, 3), 16384)}
for name, (circ, shots) in cases.items():
    prog = compile_circuits([circ]); eng = TableauEngine(prog)
    rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
    for _ in range(2): eng.run(shots, 0, 1, records=rec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.run(shots, 0, 1, records=rec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name:9s} {ms:9.3f} ms  {shots*prog.n_user_gates/ms/1e-3:.3e} shot*gates/s", flush=True)
