import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
from sdim_b200.workloads import qudit_repetition_code, rotated_surface_code
which = sys.argv[1]; shots = int(sys.argv[2])
circ = rotated_surface_code(7, 7, prob=1e-3) if which == "surface" else qudit_repetition_code(25, 25, 3, prob=1e-2)
prog = compile_circuits([circ]); eng = TableauEngine(prog)
for _ in range(2):
    eng.run(shots, 0, 1)
torch.cuda.synchronize(); print("done", eng.plan(None))
