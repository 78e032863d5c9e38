"""Per-phase clock breakdown of the cluster interpreter's measurements (needs the -DSDIMB_PHASE_CLOCKS build:
SDIMB_LIB=build/libsdimb_prof.so python tools/probe_phases.py [n] [d])."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200 import generate_random_clifford_circuit, _native as N
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prog = compile_circuits([generate_random_clifford_circuit(n, 2 * n, d, measurement_rounds=1, seed=1)])
eng = TableauEngine(prog); tab = eng.alloc_tableau(1)
buf = (C.c_ulonglong * 32)()
lib = N.lib()
eng.run(1, 0, 3, tableau=tab); lib.sdimb_debug_phase_clocks(buf)
eng.run(1, 0, 3, tableau=tab); lib.sdimb_debug_phase_clocks(buf)
names = {0: "stage row + pivot", 1: "det: factor list", 2: "det: generator columns", 3: "det: block sums", 4: "det: remote adds",
         5: "det: cluster barrier B2", 6: "rnd: factors + prefetch", 7: "rnd: cluster barrier B1", 8: "rnd: column walk + sum",
         9: "rnd: rank-1 update", 10: "rnd: column writes + remote adds", 11: "rnd: cluster barrier B2", 12: "tail (phases, record)"}
tot = sum(buf[i] for i in range(16))
for i in range(13):
    if buf[16 + i]:
        print(f"{names[i]:34s} calls {buf[16+i]:6d}  cycles/call {buf[i]/buf[16+i]:8.0f}  share {100*buf[i]/tot:5.1f}%")
print("total cycles in measurements", tot)
