import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdim_b200 import generate_random_clifford_circuit
from sdim_b200.engine import TableauEngine
from sdim_b200.ir import compile_circuits
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prog = compile_circuits([generate_random_clifford_circuit(n, 2 * n, 5, measurement_rounds=1, seed=1)])
eng = TableauEngine(prog); tab = eng.alloc_tableau(1)
for _ in range(2):
    eng.run(1, 0, 3, tableau=tab)
torch.cuda.synchronize(); print("done")
