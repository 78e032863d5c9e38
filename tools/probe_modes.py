"""A/B of interpreter choices on the sweep workload: python tools/probe_modes.py d n shots mode[,mode...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import noisy_random_clifford
d, n, shots = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["auto"]
prog = compile_circuits([noisy_random_clifford(n, 8 * n, d)])
eng = TableauEngine(prog)
base = None
for mode in modes:
    m = None if mode == "auto" else mode
    kernel, need = eng.plan(m)
    tab = eng.alloc_tableau(shots) if need else None
    rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
    eng.run(shots, 0, 1, tableau=tab, records=rec, mode=m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        eng.run(shots, 0, 1, tableau=tab, records=rec, mode=m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    r = rec.cpu().numpy()
    same = "" if base is None else f" same_records={np.array_equal(r, base)}"
    base = r if base is None else base
    print(f"d={d} n={n} shots={shots} mode={mode} kernel={kernel} cluster={eng.cluster_size(shots, m) if kernel == 'lanes-global' else 0} "
          f"{ms:.2f} ms  {shots * prog.n_user_gates / ms * 1e3:.3e} shot*gates/s{same}", flush=True)
    del tab
