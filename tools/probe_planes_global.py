"""d = 2 / 3 tableaus beyond the shared-memory limit: bit planes on a global image vs uint8 lanes on the HBM store."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdim_b200 import generate_random_clifford_circuit
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import rotated_surface_code
shots = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cases = [("surface code d=2 distance 21 (881 qubits), 3 rounds", compile_circuits([rotated_surface_code(21, 3, prob=1e-3)])),
         ("random Clifford d=3 n=600 depth 4000 + M", compile_circuits([generate_random_clifford_circuit(600, 4000, 3, measurement_rounds=1, seed=1)]))]
for name, prog in cases:
    eng = TableauEngine(prog)
    out = {}
    for mode in (None, "lanes"):
        need = eng.plan(mode)[1]
        tab = eng.alloc_tableau(shots) if need else None
        for _ in range(2):
            rec = eng.run(shots, 0, 1, mode=mode, tableau=tab)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rec = eng.run(shots, 0, 1, mode=mode, tableau=tab); e1.record(); torch.cuda.synchronize()
        out[mode] = rec.cpu().numpy()
        ms = e0.elapsed_time(e1)
        print(f"{name}: n={prog.num_qudits} ops={prog.n_ops} meas={prog.n_meas} shots={shots} kernel={eng.plan(mode)[0]} "
              f"{ms:.2f} ms  {shots * prog.n_user_gates / ms * 1e3:.3e} shot*gates/s", flush=True)
        del tab
    print("  records identical:", np.array_equal(out[None], out["lanes"]))
