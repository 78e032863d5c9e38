"""Multi-GPU check of the real N > 1 path (run under torchrun on a multi-GPU box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_nccl_check.py

Every rank simulates its contiguous shot range on its own GPU, the packed records are all-gathered over NCCL, and
every rank checks the gathered matrix against the C oracle (records depend on the global shot id only)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from make_cases import random_circuit
from oracle import c_oracle
from sdim_b200 import Program
from sdim_b200.ir import compile_circuits


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for d, n, shots in ((3, 40, 1001), (2, 97, 640), (5, 12, 333)):
        circ = random_circuit(17 + d, n, d, 20 * n)
        prog = compile_circuits([circ])
        table = Program(circ, device=f"cuda:{local}").simulate_records(shots, seed=5, distributed=True)
        rec = table.values | (table.deterministic.astype(np.uint8) << 7)
        want = c_oracle.run_philox(prog, shots, 0, 5)
        same = bool(np.array_equal(rec, want))
        ok = ok and same
        print(f"rank {rank}/{world}: d={d} n={n} shots={shots} gathered {rec.shape} match_oracle={same}", flush=True)
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
