"""Dev probe (compute-sanitizer target): a few shots through one of the round-2 two-kernel paths / the tile interpreter."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from make_cases import random_program
from oracle import c_oracle
from sdim_b200.engine import TableauEngine
d, n, depth, mode, pmeas = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], float(sys.argv[5])
shots = int(sys.argv[6]) if len(sys.argv) > 6 else 5
prog = random_program(seed=1000 * d + n, n=n, d=d, depth=depth, p_meas=pmeas)
eng = TableauEngine(prog)
got = eng.run(shots, 0, 2026 + d, mode=None if mode == "auto" else mode).cpu().numpy()
want = c_oracle.run_philox(prog, shots, 0, 2026 + d)
print("ran", eng.plan(None if mode == "auto" else mode)[0], "gate_stream" if eng.gate_stream is not None else "-", "match", bool(np.array_equal(got, want)))
