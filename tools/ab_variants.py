"""Same-box A/B of prebuilt library variants (tools/build_variants.sh) on the headline and two neighbours.

    python tools/ab_variants.py variants/libsdimb_base.so variants/libsdimb_mlp.so ...
    AB_SET=resident python tools/ab_variants.py ...      # shapes that run the shared-memory instantiation

One subprocess per library (SDIMB_LIB is read at import).  Each prints, per workload, the device time of a launch and
whether the records equal the C oracle's on the first shots; results land in gpurun_out/ab_variants.json.
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKLOADS = [  # (d, n, depth or None for 8n, shots, oracle_shots)
    (3, 256, None, 16384, 48),   # headline shape (bench.py builds the same circuit)
    (2, 400, None, 8192, 24),    # d = 2 on the slab image
    (3, 500, None, 2048, 12),    # beyond the shared-memory limit
]
if os.environ.get("AB_SET") == "resident":   # shared-memory image: 1-warp CTAs (n = 64), 4-warp CTAs (n = 160, 192)
    WORKLOADS = [(3, 64, None, 32768, 64), (2, 97, None, 32768, 64), (3, 160, None, 16384, 48), (2, 256, None, 8192, 32)]


def worker():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import noisy_random_clifford
    out = []
    for d, n, depth, shots, oshots in WORKLOADS:
        circ = noisy_random_clifford(n, 2000 if (d, n) == (3, 256) else 8 * n, d)
        prog = compile_circuits([circ])
        eng = TableauEngine(prog)
        kernel, need = eng.plan(None)
        rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            eng.run(shots, 0, 2026, records=rec)
        torch.cuda.synchronize()
        times = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(shots, 0, 2026, records=rec)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        got = rec[:oshots].cpu().numpy()
        want = c_oracle.run_philox(prog, oshots, 0, 2026)
        row = {"lib_md5": hashlib.md5(open(os.environ["SDIMB_LIB"], "rb").read()).hexdigest()[:8], "d": d, "n": n, "shots": shots, "kernel": kernel, "ms_min": min(times), "ms_med": sorted(times)[2],
               "shot_gates_per_sec": shots * prog.n_user_gates / min(times) * 1e3,
               "records_equal_oracle": bool(np.array_equal(got & 0xFF, want)),
               "records_crc": int(np.frombuffer(rec.cpu().numpy().tobytes(), dtype=np.uint8).astype(np.uint64).sum())}
        print(json.dumps(row), flush=True)
        out.append(row)
    return out


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--worker":
        worker()
        sys.exit(0)
    results = {}
    out_path = os.path.join(ROOT, "gpurun_out", "ab_variants.json")
    if os.path.exists(out_path):                       # later calls add to the same table
        with open(out_path) as f:
            results = json.load(f)
    for lib in sys.argv[1:]:
        env = dict(os.environ, SDIMB_LIB=os.path.abspath(lib))
        proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env, capture_output=True,
                              text=True, timeout=240)
        rows = [json.loads(l) for l in proc.stdout.splitlines() if l.startswith("{")]
        results[os.path.basename(lib)] = rows if proc.returncode == 0 else {"error": proc.stderr[-2000:], "rows": rows}
        print(os.path.basename(lib), json.dumps(results[os.path.basename(lib)])[:1500], flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(out_path, "w") as f:
            json.dump(results, f, indent=1)
