"""Cluster interpreter breakdown on a config-5 shaped circuit: gates only / gates + measurements, per mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdim_b200 import generate_random_clifford_circuit
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 5
shots = int(sys.argv[3]) if len(sys.argv) > 3 else 1
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["auto", "global-cta"]
reps = 3


def timed(eng, tab, mode):
    for _ in range(2):
        eng.run(shots, 0, 3, tableau=tab, keep_tableau=True, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        rec = eng.run(shots, 0, 3, tableau=tab, keep_tableau=True, mode=mode)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, rec


for rounds in (0, 1):
    prog = compile_circuits([generate_random_clifford_circuit(n, 2 * n, d, measurement_rounds=rounds, seed=1)])
    eng = TableauEngine(prog)
    tab = eng.alloc_tableau(shots)
    for mode in modes:
        m = None if mode == "auto" else mode
        ms, rec = timed(eng, tab, m)
        extra = ""
        if rounds:
            r = rec.cpu().numpy()
            extra = f" deterministic={int((r[0] & 0x80 != 0).sum())}/{prog.n_meas}"
        print(f"n={n} d={d} shots={shots} ops={prog.n_ops} meas={prog.n_meas} mode={mode} cluster={eng.cluster_size(shots, m)} "
              f"{ms:.3f} ms{extra}", flush=True)
