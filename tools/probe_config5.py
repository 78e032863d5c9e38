"""Config 5 probe: one large tableau (random Clifford, d = 5 / 7, n = 4096), parity vs the C oracle + timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdim_b200 import generate_random_clifford_circuit
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from oracle import c_oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gates = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * n
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else [None, "global-cta"]   # None = auto (cluster for few shots)
for d in (5, 7):
    prog = compile_circuits([generate_random_clifford_circuit(n, gates, d, measurement_rounds=1, seed=1)])
    eng = TableauEngine(prog)
    tab = eng.alloc_tableau(1)
    t0 = time.time(); want, fin = c_oracle.run(n, d, prog.ops, 1, 0, 3, want_final=True); cpu = time.time() - t0
    for mode in modes:
        mode = None if mode in (None, "auto") else mode
        for _ in range(2):
            rec = eng.run(1, 0, 3, tableau=tab, keep_tableau=True, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rec = eng.run(1, 0, 3, tableau=tab, keep_tableau=True, mode=mode); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        got = rec.cpu().numpy()
        arrs = eng.export(tab, 0)
        ok = np.array_equal(got, want) and all(np.array_equal(arrs[k], fin[k]) for k in fin)
        print(f"d={d} n={n} ops={prog.n_ops} mode={mode or 'auto'} cluster={eng.cluster_size(1, mode)} gpu={ms:.2f} ms  "
              f"c_oracle(1 thread)={cpu*1e3:.0f} ms  parity={'OK' if ok else 'MISMATCH'}  {prog.n_user_gates/ms*1e3:.3e} gates/s",
              flush=True)
