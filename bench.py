#!/usr/bin/env python
"""Benchmark of the tableau hot path — contract in the task prompt (section 4 / "Maintain bench.py").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shots S]

Workload (BASELINE.json metric, SURVEY 8d "headline"): d = 3, n = 256 noisy random Clifford:
generate_random_clifford_circuit(256, 2000, 3, seed=1) with `N1 prob=1e-3 noise_channel='d'` after every
gate on the qudit(s) it touched, then M on all 256 qudits; one full stabilizer tableau per shot, Philox seed
2026.  A step = one pass of the whole circuit over `--shots` shots per GPU.  Metric: shot*gates/s with
gates = len(circuit.operations) as the user wrote them.

`value`  : device-timed (CUDA events on the launch stream), op stream already resident in HBM.
`e2e`    : the same metric through the host-buffer C-ABI call sdimb_simulate_host (host op stream in,
           host records out; H2D/D2H and scratch allocation inside the timed region).
`--impl reference` times the CPU oracle port of the reference algorithm (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "shot_gates_per_sec"
UNIT = "shot*gates/s"
WORKLOAD = dict(n=256, d=3, gates=2000, circuit_seed=1, noise_prob=1e-3, noise_channel="d", philox_seed=2026)


def build_workload():
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import noisy_random_clifford
    w = WORKLOAD
    circ = noisy_random_clifford(w["n"], w["gates"], w["d"], seed=w["circuit_seed"], prob=w["noise_prob"],
                                 channel=w["noise_channel"])
    return circ, compile_circuits([circ])


def algorithmic_bytes_per_shot(prog, det_flags, meas_nnz=None) -> float:
    """SURVEY 8d per-op byte table (un-fused streaming model), N = 2n lanes, 1 byte per entry/phase for odd d,
    1 bit / 2 bits for d = 2.  det_flags[k]: measurement k was deterministic (shot-invariant).

    meas_nnz[k] (from the oracle) = generators with a non-zero factor in measurement k.  With it, a measurement
    counts only the columns the reference itself touches (it skips zero factors, tableau_prime.py:308,315,351):
    random: row q (N) + pivot column (2n) + nnz * 4n (X,Z columns read+write) + nnz * 2 phases + column writes 4n;
    deterministic: row q destab half (n) + nnz * (2n + 1).  Without it the dense 8d figure is used."""
    n, d = prog.num_qudits, prog.dimension
    we, wp = (1.0, 1.0) if d != 2 else (1.0 / 8, 2.0 / 8)
    N = 2 * n
    pauli = N * we + 2 * N * wp
    table = {1: pauli, 2: pauli, 3: pauli, 4: pauli,
             5: 4 * N * we + 2 * N * wp, 6: 4 * N * we + 2 * N * wp,
             7: 3 * N * we + 2 * N * wp, 8: 3 * N * we + 2 * N * wp,
             9: 6 * N * we, 10: 6 * N * we,
             11: 6 * N * we + 2 * N * wp, 12: 6 * N * we + 2 * N * wp,
             13: 8 * N * we}
    total = 0.0
    for op, _a, _b, slot in prog.ops:
        op = int(op)
        if op in table:
            total += table[op]
        elif op in (14, 15, 16):
            if op == 15:
                total += table[6]
            k = int(slot)
            if meas_nnz is None:
                total += (2 * n * n * we + n * we + n * wp + 1) if det_flags[k] else (4 * N * n * we + 2 * N * wp + 1)
            elif det_flags[k]:
                total += n * we + meas_nnz[k] * (2 * n * we + wp) + 1
            else:
                total += N * we + 2 * n * we + meas_nnz[k] * (4 * n * we + 2 * wp) + 4 * n * we + 1
            if op == 16:
                total += pauli * (d - 1) / d          # X^k correction, k != 0 with prob (d-1)/d on random outcomes
        elif op == 17:
            pr = float(prog.noise_prob[int(slot)])
            ch = int(prog.noise_channel[int(slot)])
            if ch == 0:                               # 'd': both exponents non-zero in (d-1)^2 of d^2-1 cases
                both = (d - 1) ** 2 / (d * d - 1)
                total += pr * (both * (2 * N * we + 2 * N * wp) + (1 - both) * pauli)
            else:
                total += pr * pauli
    return total


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in self.rows:
            if len(row) < 7:
                continue
            try:
                sm.append(float(row[0])); smax.append(float(row[1]))
            except ValueError:
                continue
            for name, cell in zip(names, row[3:7]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_cpu_oracle(prog, shots: int, seed: int):
    """Time the oracle port on `shots` shots of the workload; returns (seconds, cores, kind, records)."""
    from oracle import c_oracle
    if c_oracle.available():
        t0 = time.perf_counter()
        rec = c_oracle.run_philox(prog, shots, 0, seed)
        return time.perf_counter() - t0, c_oracle.threads(), "port", rec
    from oracle.tableau_oracle import run_shots
    from sdim_b200.rng import measurement_draws, noise_draws
    ids = np.arange(shots)
    md = measurement_draws(seed, prog.dimension, ids, prog.n_meas)
    nd = noise_draws(seed, prog.dimension, ids, prog.noise_thresh24, prog.noise_channel) if prog.n_noise else None
    t0 = time.perf_counter()
    rec, _ = run_shots(prog.num_qudits, prog.dimension, prog.ops, shots, md, nd)
    return time.perf_counter() - t0, 1, "port", rec


def cpu_sample_size(prog, seed: int, target_s: float, requested: int) -> int:
    """Shots whose CPU simulation takes about `target_s` seconds on this box (calibrated on a short run)."""
    if requested > 0:
        return requested
    probe = 64
    dt, _, _, _ = run_cpu_oracle(prog, probe, seed)
    dt2, _, _, _ = run_cpu_oracle(prog, probe, seed)
    per_shot = max(min(dt, dt2), 1e-6) / probe
    return int(min(max(probe, target_s / per_shot), 1 << 16))


def bench_reference(args):
    """--impl reference: the CPU restatement of the reference algorithm on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    _, prog = build_workload()
    gates = prog.n_user_gates
    sample = cpu_sample_size(prog, WORKLOAD["philox_seed"], 4.0, args.cpu_shots)    # ~4 s of CPU work per step
    for _ in range(args.warmup):
        run_cpu_oracle(prog, max(1, sample // 8), WORKLOAD["philox_seed"])
    times = []
    cores = 1
    for _ in range(args.steps):
        dt, cores, kind, _ = run_cpu_oracle(prog, sample, WORKLOAD["philox_seed"])
        times.append(dt)
    total = sum(times)
    value = sample * args.steps * gates / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(), "shots_per_step": sample, "gates_per_shot": gates},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} shots x {gates} ops per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name():
    w = WORKLOAD
    return (f"noisy random Clifford d={w['d']} n={w['n']}: generate_random_clifford_circuit({w['n']},{w['gates']},"
            f"{w['d']},seed={w['circuit_seed']}) + N1(p={w['noise_prob']},'{w['noise_channel']}') after every gate "
            f"+ M on all qudits; one tableau per shot")


def bench_ours(args):
    import torch
    import torch.distributed as dist
    from sdim_b200 import _native as N
    from sdim_b200.engine import TableauEngine, simulate_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    _, prog = build_workload()
    gates = prog.n_user_gates
    shots = args.shots                      # per GPU (weak scaling: per-GPU work fixed)
    seed = WORKLOAD["philox_seed"]
    engine = TableauEngine(prog, dev)
    L = engine.layout
    kernel_name, need_tab = engine.plan(args.mode)
    tab = engine.alloc_tableau(shots) if need_tab else None
    records = torch.empty((shots, prog.n_meas), dtype=torch.uint8, device=dev)
    lo = rank * shots                       # global shot ids of this rank
    stream = torch.cuda.current_stream(dev)

    def step():
        engine.run(shots, lo, seed, mode=args.mode, tableau=tab, records=records)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def flush_l2():
        flush_buf.fill_(1)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = N.lib().sdimb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        for e0, e1 in ev:
            flush_l2()                      # untimed: evict the previous step's lines from L2
            e0.record(stream)
            step()
            e1.record(stream)
        barrier()
        time.sleep(0.25)
    launches = N.lib().sdimb_launch_count() - launches0
    kernel_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = float(sum(kernel_ms))        # exactly K timed steps, device time on the launch stream
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * shots * args.steps * gates / (total_ms * 1e-3)

    # ---- end to end through the host-buffer C ABI call -------------------------------------------
    e2e_shots = shots
    e2e_times = []
    simulate_host(prog, e2e_shots, lo, seed, mode=args.mode)   # warm-up: sizes the library's reusable workspace
    for _ in range(max(1, min(args.steps, 3))):
        barrier()
        t0 = time.perf_counter()
        rec_host, _ms = simulate_host(prog, e2e_shots, lo, seed, mode=args.mode)
        e2e_times.append(time.perf_counter() - t0)
    e2e_t = float(np.mean(e2e_times))
    if world > 1:
        t = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    e2e_value = world * e2e_shots * gates / e2e_t
    up_rows = N.schedule(prog.num_qudits, prog.ops).shape[0] if kernel_name.startswith("planes") else prog.n_ops
    h2d = up_rows * 16 + prog.noise_thresh24.nbytes + prog.noise_channel.nbytes      # what the call uploads
    d2h = e2e_shots * prog.n_meas

    # the device-buffer path and the host-buffer path must agree bit for bit (same Philox counters)
    same = bool(np.array_equal(records.cpu().numpy(), rec_host))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the interpreter launch = one step) -------------------------
    det_flags = (records[0].cpu().numpy() & 0x80) != 0
    # accounting only (never on the measured path): how many generators each measurement really touches
    from oracle import c_oracle
    meas_nnz = c_oracle.measurement_factor_counts(prog, seed) if c_oracle.available() else None
    alg_bytes = algorithmic_bytes_per_shot(prog, det_flags, meas_nnz) * shots
    alg_bytes_dense = algorithmic_bytes_per_shot(prog, det_flags, None) * shots
    launch_s = float(np.mean(kernel_ms)) * 1e-3
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as fh:
            peak, peak_src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj.get("shots"):
            traffic = tj["dram_bytes_per_launch"] * (shots / tj["shots"])

    # ---- BASELINE.json's second metric: gate-update HBM GB/s vs peak.  The gates of the same circuit (no noise,
    # no measurement) streamed over one uint8 tableau per shot in HBM by the lane interpreter ("global" mode):
    # every gate reads/writes its dense rows, SURVEY 8d bytes, no residency.  N = 1 only.
    gate_update = None
    if world == 1 and not args.no_gate_update:
        from sdim_b200 import generate_random_clifford_circuit
        from sdim_b200.ir import compile_circuits
        w = WORKLOAD
        gprog = compile_circuits([generate_random_clifford_circuit(w["n"], w["gates"], w["d"], 0, w["circuit_seed"])])
        geng = TableauEngine(gprog, dev)
        gshots = 148 * 32 * 2        # two full waves of the streaming launch shape (32 one-warp CTAs per SM)
        gtab = geng.alloc_tableau(gshots)
        geng.init_tableau(gtab)
        grec = torch.empty((gshots, 0), dtype=torch.uint8, device=dev)
        gtimes = []
        for it in range(3 + 3):
            flush_l2()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            geng.run(gshots, 0, seed, mode="global", tableau=gtab, fresh=False, keep_tableau=True, records=grec)
            g1.record(stream)
            torch.cuda.synchronize(dev)
            if it >= 3:
                gtimes.append(g0.elapsed_time(g1))
        gbytes = algorithmic_bytes_per_shot(gprog, [], None) * gshots
        gms = float(np.mean(gtimes))
        gate_update = {"kernel": "interp_kernel_stream (uint8 lanes, 16 lanes per thread, one tableau per shot in HBM)",
                       "workload": f"{gprog.n_ops} gates of the headline circuit, {gshots} shots, "
                                   f"{gshots * L.shot_bytes / 2**30:.2f} GiB store, tableaus evolve across launches",
                       "achieved": gbytes / (gms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": gbytes / (gms * 1e-3) / 1e9 / peak, "launch_ms": gms,
                       "shot_gates_per_sec": gshots * gprog.n_ops / (gms * 1e-3)}
        del gtab

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu_shots = min(cpu_sample_size(prog, seed, 15.0, args.cpu_shots), shots)     # ~15 s of CPU work
        dt, cores, kind, cpu_rec = run_cpu_oracle(prog, cpu_shots, seed)
        cpu_ok = bool(np.array_equal(cpu_rec, rec_host[:cpu_shots]))
        cpu = {"value": cpu_shots * gates / dt, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"first {cpu_shots} shots of the same circuit/seed, C restatement of the reference "
                         f"algorithm with OpenMP over shots ({dt:.1f} s)",
               "records_match_gpu": cpu_ok}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(), "shots_per_gpu_per_step": shots, "gates_per_shot": gates,
                   "ops_in_stream": prog.n_ops, "n_meas": prog.n_meas, "n_noise": prog.n_noise,
                   "mode": args.mode or "auto", "kernel": kernel_name,
                   "l2": "L2 flushed (256 MiB fill) between timed steps, outside the timed region",
                   "tableau_store_mib_per_step": (shots * L.shot_bytes / 2**20) if need_tab else 0.0,
                   "parallelism": f"shots sharded over {world} GPU(s), no data-path collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "records_match_device_path": same},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "interp_planes_kernel" if kernel_name.startswith("planes") else "interp_kernel", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": launch_s * 1e3,
                     "accounting": "SURVEY 8d per-op bytes; measurements count only generators with non-zero factor "
                                   "(the reference's own skip rule), see DESIGN.md section 4",
                     "achieved_dense": alg_bytes_dense / launch_s / 1e9},
        "gate_update_hbm": gate_update,
        "cpu_baseline": cpu,
        "clocks": clocks.summary(),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shots", type=int, default=16384, help="shots per GPU per step")
    ap.add_argument("--cpu-shots", type=int, default=0, help="shots in the CPU baseline sample (0 = calibrate)")
    ap.add_argument("--mode", default=None, choices=[None, "auto", "global", "resident", "lanes", "planes"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gate-update", action="store_true")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched bare with --gpus N: become the torchrun launch the contract describes (one rank per GPU)
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
