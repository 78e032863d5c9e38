import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from make_cases import random_program
from sdim_b200.engine import TableauEngine
d, n, depth, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
prog = random_program(seed=1000 * d + n, n=n, d=d, depth=depth)
TableauEngine(prog).run(2, 0, 2026 + d, mode=mode).cpu()
print("ran")
