import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from make_cases import random_program
from oracle import c_oracle
from sdim_b200.engine import TableauEngine
d, n, depth = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prog = random_program(seed=1000 * d + n, n=n, d=d, depth=depth)
shots, seed = 96, 2026 + d
want = c_oracle.run_philox(prog, shots, 0, seed)
for mode in ("resident", "global", "planes", "resident", "global"):
    try:
        got = TableauEngine(prog).run(shots, 0, seed, mode=mode).cpu().numpy()
    except ValueError as e:
        print(mode, "n/a", e); continue
    bad = np.argwhere(got != want)
    print(mode, "bad records", len(bad), "bad shots", len(set(bad[:,0])) if len(bad) else 0, "first", bad[:3].tolist())
    if len(bad):
        s, k = bad[0]
        ops_m = [i for i,o in enumerate(prog.ops) if o[0] in (14,15,16)]
        print("   op index of first bad meas:", ops_m[k], "op", prog.ops[ops_m[k]].tolist(), "got", got[s,k], "want", want[s,k])
