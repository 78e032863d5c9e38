"""Dev probe: cost of the measurement rounds of the headline circuit (d = 3, n = 256): round 1 = 159 random + 97
deterministic measurements, rounds 2.. = 256 deterministic ones each (the state is an eigenstate by then)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200 import generate_random_clifford_circuit
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits

n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 3
shots = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
mode = sys.argv[4] if len(sys.argv) > 4 else None
prev = None
for rounds in (0, 1, 2, 3):
    prog = compile_circuits([generate_random_clifford_circuit(n, 2000, d, measurement_rounds=rounds, seed=1)])
    eng = TableauEngine(prog)
    rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.run(shots, 0, 1, mode=mode, records=rec)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(shots, 0, 1, mode=mode, records=rec); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    det = int(((rec[0] & 0x80) != 0).sum()) if prog.n_meas else 0
    print(f"rounds={rounds} ops={prog.n_ops} meas={prog.n_meas} det={det} {ms:8.3f} ms" + (f"  (+{ms - prev:.3f} ms for this round)" if prev is not None else ""), flush=True)
    prev = ms
