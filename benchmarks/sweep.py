"""Throughput across tableau sizes and dimensions on one B200: which interpreter `auto` picks and what it delivers.
Workload per point: generate_random_clifford_circuit(n, 8 n, d, seed = 1) + N1(p = 1e-3, 'd') after every gate + M on
all qudits (the headline recipe at other sizes), one tableau per shot.  Prints one JSON document."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sdim_b200.engine import TableauEngine  # noqa: E402
from sdim_b200.ir import compile_circuits  # noqa: E402
from sdim_b200.workloads import noisy_random_clifford  # noqa: E402


def point(n, d, shots, reps=2):
    prog = compile_circuits([noisy_random_clifford(n, 8 * n, d)])
    eng = TableauEngine(prog)
    mode = eng._auto_mode(None, shots)               # what `auto` resolves to for this shot count
    kernel, need_tab = eng.plan(mode)
    csize = eng.cluster_size(shots, mode) if kernel == "lanes-global" else 0
    tab = eng.alloc_tableau(shots) if need_tab else None
    rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
    eng.run(shots, 0, 1, tableau=tab, records=rec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.run(shots, 0, 1, tableau=tab, records=rec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    path = kernel
    if kernel == "planes-global" and eng.tail_run_len and n <= 512:
        path = ("gate_stream_kernel" if eng.gate_stream is not None else "interp_planes_kernel") + " + run_tail_kernel"
    elif kernel == "lanes-global" and eng.tail_run_len_raw and not csize and n <= 512 and d <= 127:
        path = "lane interpreter + run_tail8_kernel"
    return {"n": n, "d": d, "shots": shots, "ops": prog.n_ops, "gates_per_shot": prog.n_user_gates,
            "kernel": kernel + (f" ({csize}-CTA clusters)" if csize else ""), "path": path, "ms": ms,
            "shot_gates_per_sec": shots * prog.n_user_gates / ms * 1e3}


if __name__ == "__main__":
    rows = []
    for d in (2, 3, 5, 7):
        for n, shots in ((16, 65536), (64, 32768), (128, 16384), (256, 8192), (512, 2048), (1024, 512), (2048, 8)):
            try:
                rows.append(point(n, d, shots))
            except Exception as exc:                     # keep the sweep going, report the failure
                rows.append({"n": n, "d": d, "shots": shots, "error": str(exc)[:200]})
            print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "rows": rows}, indent=1))
