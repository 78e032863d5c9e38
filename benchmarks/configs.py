#!/usr/bin/env python
"""Throughput of every BASELINE.json config on one GPU (bench.py covers only the headline metric).

    python benchmarks/configs.py [--quick] > gpurun_out/configs.json

For each config: device-timed shot*gates/s of the per-shot tableau path (CUDA events, 3 timed launches after 2
warm-ups), which interpreter ran, the Pauli-frame sampler where it applies, and the C restatement of the reference
algorithm on the host cores for a bounded sample of the same workload (a reported baseline, not a target).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import c_oracle  # noqa: E402
from sdim_b200 import generate_random_clifford_circuit, read_circuit  # noqa: E402
from sdim_b200.engine import TableauEngine  # noqa: E402
from sdim_b200.ir import compile_circuits  # noqa: E402
from sdim_b200.workloads import noisy_random_clifford, qudit_repetition_code, rotated_surface_code  # noqa: E402


def time_tableau(prog, shots, reps=3):
    eng = TableauEngine(prog)
    kernel, need_tab = eng.plan(None)
    tab = eng.alloc_tableau(shots) if need_tab else None
    rec = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.run(shots, 0, 1, tableau=tab, records=rec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.run(shots, 0, 1, tableau=tab, records=rec)
    e1.record()
    torch.cuda.synchronize()
    csize = eng.cluster_size(shots) if kernel == "lanes-global" else 0
    if csize:
        kernel += f", {csize}-CTA cluster per shot"
    return eng, kernel, e0.elapsed_time(e1) / reps


def time_frames(eng, prog, shots, reps=3):
    quiet = torch.zeros((1, prog.n_noise, 2), dtype=torch.uint8) if prog.n_noise else None
    ref = eng.run(1, 0, 1, None, quiet)
    eng.run_frames(shots, ref[0], 1, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.run_frames(shots, ref[0], 1, 1)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def cpu_rate(prog, target_s=3.0):
    if not c_oracle.available():
        return None, 0
    probe = 8
    t0 = time.perf_counter(); c_oracle.run_philox(prog, probe, 0, 1); dt = time.perf_counter() - t0
    shots = int(min(max(probe, target_s / max(dt / probe, 1e-7)), 1 << 18))
    t0 = time.perf_counter(); c_oracle.run_philox(prog, shots, 0, 1); dt = time.perf_counter() - t0
    return shots * prog.n_user_gates / dt, shots


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    q = 8 if args.quick else 1
    configs = [
        ("1a epr.chp d=2 n=2, 1k shots", read_circuit("circuits/epr.chp"), 1000),
        ("1b css_steane_final.chp d=2 n=13, 1k shots", read_circuit("circuits/css_steane_final.chp"), 1000),
        ("2 random Clifford d=3 n=64 depth 2k + M, 1e4 shots",
         generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1), 10000),
        ("3 surface code d=2 distance 7 (97 qubits), 7 rounds, depolarising p=1e-3, 1e6 shots",
         rotated_surface_code(7, 7, prob=1e-3), 1000000 // q),
        ("4 qutrit repetition code distance 25, 25 rounds, flip noise p=1e-2 + reset, 1e7 shots (one wave of 2e6)",
         qudit_repetition_code(25, 25, 3, prob=1e-2), 2000000 // q),
        ("headline noisy random Clifford d=3 n=256, 16384 shots", noisy_random_clifford(256, 2000, 3), 16384),
        ("5 single tableau random Clifford d=5 n=4096, 1 shot",
         generate_random_clifford_circuit(4096, 8192, 5, measurement_rounds=1, seed=1), 1),
        ("5 single tableau random Clifford d=7 n=4096, 1 shot",
         generate_random_clifford_circuit(4096, 8192, 7, measurement_rounds=1, seed=1), 1),
    ]
    out = []
    for name, circ, shots in configs:
        prog = compile_circuits([circ])
        eng, kernel, ms = time_tableau(prog, shots)
        row = {"config": name, "n": prog.num_qudits, "d": prog.dimension, "ops": prog.n_user_gates,
               "n_meas": prog.n_meas, "shots": shots, "kernel": kernel, "ms_per_launch": ms,
               "shot_gates_per_sec": shots * prog.n_user_gates / (ms * 1e-3)}
        if shots > 1:
            fms = time_frames(eng, prog, shots)
            row["frame_sampler_shot_gates_per_sec"] = shots * prog.n_user_gates / (fms * 1e-3)
        rate, cshots = cpu_rate(prog)
        row["cpu_port_shot_gates_per_sec"] = rate
        row["cpu_port_sample_shots"] = cshots
        row["cpu_threads"] = c_oracle.threads() if c_oracle.available() else 0
        out.append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
        del eng
        torch.cuda.empty_cache()
    json.dump({"gpu": torch.cuda.get_device_name(0), "rows": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
