"""Configs 2-4 (small tableaus, many shots): time with and without the measurement / noise ops."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200 import Circuit, generate_random_clifford_circuit
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import qudit_repetition_code, rotated_surface_code


def strip(circ, drop):
    out = Circuit(circ.num_qudits, circ.dimension)
    for op in circ.operations:
        if op.name in drop:
            continue
        if op.target_index is None:
            out.add_gate(op.name, op.qudit_index, **(op.params or {}))
        else:
            out.add_gate(op.name, op.qudit_index, op.target_index)
    return out


MODE = sys.argv[1] if len(sys.argv) > 1 else None          # e.g. planes (tile interpreter where it fits) / planes-warp


def timeit(prog, shots, reps=3):
    eng = TableauEngine(prog)
    rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.run(shots, 0, 1, records=rec, mode=MODE)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.run(shots, 0, 1, records=rec, mode=MODE)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cases = [("config 2 random Clifford d=3 n=64", generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1), 100000),
         ("config 3 surface code d=2 distance 7", rotated_surface_code(7, 7, prob=1e-3), 200000),
         ("config 4 qutrit repetition code", qudit_repetition_code(25, 25, 3, prob=1e-2), 200000)]
only = sys.argv[2] if len(sys.argv) > 2 else ""
for name, circ, shots in cases:
    if only and only not in name:
        continue
    full = compile_circuits([circ])
    gates = compile_circuits([strip(circ, {"M", "M_X", "RESET", "N1"})])
    nonoise = compile_circuits([strip(circ, {"N1"})])
    t_full, t_gates, t_nn = timeit(full, shots), timeit(gates, shots), timeit(nonoise, shots)
    print(f"{name}: shots={shots} ops={full.n_ops} meas={full.n_meas} noise={full.n_noise}  full {t_full:.2f} ms  "
          f"no noise {t_nn:.2f} ms  gates only ({gates.n_ops} ops) {t_gates:.2f} ms", flush=True)
