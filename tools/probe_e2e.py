import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sdim_b200.engine import TableauEngine, simulate_host
_, prog = bench.build_workload()
shots = 16384
simulate_host(prog, 256, 0, 1)
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); rec, ms = simulate_host(prog, shots, 0, 2026); dt = time.perf_counter() - t0
    print(f"wall {dt*1e3:.2f} ms  device(e0..e1) {ms:.2f} ms  -> {shots*prog.n_user_gates/dt:.3e} shot*gates/s")
eng = TableauEngine(prog)
recd = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); eng.run(shots, 0, 2026, records=recd); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"device path wall {dt*1e3:.2f} ms")
