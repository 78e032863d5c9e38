"""Dev probe: front / tail kernel times of the headline (or n d shots) for one form of the front kernel.
The form comes from the environment (SDIMB_NO_GATE_STREAM, SDIMB_GS_GLOBAL, SDIMB_GS_WARPS): one process per form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdim_b200 import _native as N
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import noisy_random_clifford

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shots = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
mode = next((a.split("=")[1] for a in sys.argv if a.startswith("--mode=")), None)
prog = compile_circuits([noisy_random_clifford(n, 2000, d)])
eng = TableauEngine(prog)
rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.run(shots, 0, 1, records=rec, mode=mode)
torch.cuda.synchronize()
fr, tl, tot = [], [], []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run(shots, 0, 1, records=rec, time_kernels=True, mode=mode)
    e1.record(); torch.cuda.synchronize()
    kt = N.kernel_times()
    tot.append(e0.elapsed_time(e1))
    if kt: fr.append(kt[0]); tl.append(kt[1])
form = {k: v for k, v in os.environ.items() if k.startswith("SDIMB_")}
chk = ""
if "--check" in sys.argv:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import c_oracle
    want = c_oracle.run_philox(prog, 64, 0, 1)
    chk = f" oracle_match={bool(np.array_equal(rec[:64].cpu().numpy(), want))}"
print(f"n={n} d={d} shots={shots} mode={mode} form={form} gate_stream={'yes' if eng.gate_stream is not None else 'no'} "
      f"front={np.median(fr) if fr else float('nan'):.3f} ms tail={np.median(tl) if tl else float('nan'):.3f} ms "
      f"total={np.median(tot):.3f} ms{chk}")
