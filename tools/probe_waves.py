"""Wall time of Program.simulate_records for a large shot count (config 4 shape): one wave vs pipelined waves."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200 import Program
from sdim_b200.workloads import qudit_repetition_code
shots = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
circ = qudit_repetition_code(25, 25, 3, prob=0.01)
for label, wave_bytes in (("pipelined 256 MiB waves", 256 << 20), ("one wave", 1 << 40), ("pipelined 256 MiB waves", 256 << 20)):
    prog = Program(circ)
    prog.WAVE_RECORD_BYTES = wave_bytes
    prog.simulate_records(1000, seed=1)                      # warm-up: engine, library, context
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table = prog.simulate_records(shots, seed=1)
    dt = time.perf_counter() - t0
    n_gates = len(circ.operations) if hasattr(circ, "operations") else 0
    print(f"{label:26s} shots={shots} wall={dt:.3f} s  checksum={int(table.values[::997].sum())}", flush=True)
