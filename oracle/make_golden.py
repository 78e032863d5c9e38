"""TEST INFRASTRUCTURE ONLY: regenerate tests/golden/*.json from the UNMODIFIED reference.

    python oracle/make_golden.py            # needs /root/reference (this container only)

Every case is one shot of the reference's own Program.simulate on a seeded
random circuit over ALL tableau-path gate kinds (13 unitaries, M, M_X, RESET,
and N1 replayed as explicit Paulis).  Stored per case: the op list, the noise
draws, the chronological records (whose values double as the replay vector for
random measurements) and the six final arrays.  Cases where the reference
overflows int64 in its lazy-reduction window (detected by
ref_harness.ref_run_eager_modulo disagreeing) are dropped, and the count of
dropped cases is stored in the file header.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

UNITARY_1 = [1, 2, 3, 4, 5, 6, 7, 8]
UNITARY_2 = [9, 10, 11, 12, 13]


def random_ops(rng: random.Random, n: int, depth: int, d: int, p_meas=0.12, p_noise=0.08):
    """Random op list over every opcode; returns (ops, n_meas, n_noise)."""
    ops, k, j = [], 0, 0
    for _ in range(depth):
        u = rng.random()
        if u < p_meas:
            op = rng.choice([14, 14, 15, 16])
            ops.append([op, rng.randrange(n), -1, k])
            k += 1
        elif u < p_meas + p_noise:
            ops.append([17, rng.randrange(n), -1, j])
            j += 1
        elif n >= 2 and rng.random() < 0.4:
            a, b = rng.sample(range(n), 2)
            ops.append([rng.choice(UNITARY_2), a, b, -1])
        else:
            ops.append([rng.choice(UNITARY_1 + [0]), rng.randrange(n), -1, -1])
    return ops, k, j


def final_measure_all(ops, n, k):
    for q in range(n):
        ops.append([14, q, -1, k])
        k += 1
    return k


def make_case(seed: int, n: int, d: int, depth: int):
    rng = random.Random(seed)
    ops, k, j = random_ops(rng, n, depth, d)
    k = final_measure_all(ops, n, k)
    noise = [[rng.randrange(d), rng.randrange(d)] if rng.random() < 0.7 else [0, 0] for _ in range(j)]
    noise_arr = np.array(noise, dtype=np.int64).reshape(-1, 2)
    recs, arrs = rh.ref_run(n, d, ops, noise_arr, draw_seed=seed)
    recs2, arrs2 = rh.ref_run_eager_modulo(n, d, ops, noise_arr, draw_seed=seed)
    overflow = recs != recs2 or any(not np.array_equal(arrs[key], arrs2[key]) for key in arrs)
    if overflow:
        return None
    return {
        "seed": seed, "n": n, "d": d, "ops": ops, "noise_ab": noise,
        "records": [[q, int(det), m] for q, det, m in recs],
        "final": {key: val.tolist() for key, val in arrs.items()},
    }


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    plan = []
    seed = 1000
    for d in (2, 3, 5, 7, 11, 13):
        for n, depth, reps in ((1, 24, 3), (2, 40, 4), (3, 60, 4), (5, 90, 4), (8, 140, 3), (13, 200, 2)):
            for _ in range(reps):
                plan.append((seed, n, d, depth))
                seed += 1
    cases, dropped = [], 0
    for seed, n, d, depth in plan:
        case = make_case(seed, n, d, depth)
        if case is None:
            dropped += 1
        else:
            cases.append(case)
    out = {"generator": "oracle/make_golden.py", "reference": "events555/sdim @ /root/reference",
           "dropped_for_reference_int64_overflow": dropped, "cases": cases}
    path = os.path.join(GOLDEN_DIR, "random_circuits.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print(f"wrote {len(cases)} cases ({dropped} dropped) -> {path} ({os.path.getsize(path)} bytes)")

    # The two shipped circuits (SURVEY section 4 golden vectors), through the reference's own reader.
    sdim = rh.load_reference()
    shipped = {}
    for name in ("circuits/css_steane_final.chp", "circuits/epr.chp"):
        circ = sdim.read_circuit(name)
        import sdim.tableau.tableau_prime as tp
        saved = tp.random
        tp.random = rh._ChoiceFeed(random.Random(7))
        try:
            prog = sdim.Program(circ)
            res = prog.simulate(shots=1)
        finally:
            tp.random = saved
        shipped[name] = {
            "num_qudits": circ.num_qudits, "dimension": circ.dimension,
            "ops": [[op.gate_id, op.qudit_index, -1 if op.target_index is None else op.target_index]
                    for op in circ.operations],
            "flat_results": [[r.qudit_index, int(r.deterministic), int(r.measurement_value)] for r in res],
            "final": {k: v.tolist() for k, v in rh._final_arrays(prog.stabilizer_tableau).items()},
        }
    path = os.path.join(GOLDEN_DIR, "shipped_circuits.json")
    with open(path, "w") as fh:
        json.dump({"generator": "oracle/make_golden.py", "circuits": shipped}, fh, separators=(",", ":"))
    print(f"wrote {path}")


def main_large():
    """Large primes (uint8-lane limit d = 127 included): pins the oracles where the GPU parity tests lean on the
    C oracle alone.  Shallower circuits: the reference's lazy reduction multiplies entries by up to d per CNOT
    (SURVEY Appendix B-1), cases it overflows in are dropped like above."""
    plan, seed = [], 5000
    for d in (17, 31, 61, 127):
        for n, depth, reps in ((1, 20, 2), (2, 30, 3), (3, 45, 3), (5, 60, 3), (8, 80, 2)):
            for _ in range(reps):
                plan.append((seed, n, d, depth))
                seed += 1
    cases, dropped = [], 0
    for seed, n, d, depth in plan:
        case = make_case(seed, n, d, depth)
        if case is None:
            dropped += 1
        else:
            cases.append(case)
    out = {"generator": "oracle/make_golden.py --large", "reference": "events555/sdim @ /root/reference",
           "dropped_for_reference_int64_overflow": dropped, "cases": cases}
    path = os.path.join(GOLDEN_DIR, "large_primes.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print(f"wrote {len(cases)} cases ({dropped} dropped) -> {path} ({os.path.getsize(path)} bytes)")


def main_wide():
    """Primes above 127 (the uint16-lane path of the CUDA store, csrc/wide.cuh): same construction as --large.
    Shallow circuits for the same overflow reason; 32749 is the largest prime the store accepts."""
    plan, seed = [], 7000
    for d in (131, 251, 257, 1031, 32749):
        for n, depth, reps in ((1, 16, 2), (2, 24, 3), (3, 36, 3), (5, 48, 3), (8, 60, 2)):
            for _ in range(reps):
                plan.append((seed, n, d, depth))
                seed += 1
    cases, dropped = [], 0
    for seed, n, d, depth in plan:
        case = make_case(seed, n, d, depth)
        if case is None:
            dropped += 1
        else:
            cases.append(case)
    out = {"generator": "oracle/make_golden.py --wide", "reference": "events555/sdim @ /root/reference",
           "dropped_for_reference_int64_overflow": dropped, "cases": cases}
    path = os.path.join(GOLDEN_DIR, "wide_primes.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print(f"wrote {len(cases)} cases ({dropped} dropped) -> {path} ({os.path.getsize(path)} bytes)")


def config_cases():
    """(name, n, d, ops, noise_ab) of the BASELINE.json config-size shots: circuits from the product's workload
    builders (sdim_b200/workloads.py, same constructions as SURVEY 8d), N1 events as the (a, b) draws of Philox
    seed 2026, global shot 0 — i.e. exactly shot 0 of a free-running GPU run — except that configs 3 and 4 also get a
    second shot drawn at 20x the noise probability so that several events fire."""
    from sdim_b200.ir import compile_circuits
    from sdim_b200.random_circuit import generate_random_clifford_circuit
    from sdim_b200.rng import noise_draws, prob_to_thresh24
    from sdim_b200.workloads import noisy_random_clifford, qudit_repetition_code, rotated_surface_code
    out = []

    def add(name, circ, boost=1.0, shot=0):
        prog = compile_circuits([circ])
        noise = np.zeros((0, 2), dtype=np.int64)
        if prog.n_noise:
            thr = prog.noise_thresh24 if boost == 1.0 else np.array(
                [prob_to_thresh24(min(1.0, boost * p)) for p in prog.noise_prob], dtype=np.uint32)
            noise = noise_draws(2026, prog.dimension, np.array([shot]), thr, prog.noise_channel)[0].astype(np.int64)
        out.append((name, prog.num_qudits, prog.dimension, prog.ops.tolist(), noise))

    add("config2_random_clifford_n64_d3", generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1))
    add("config3_surface_code_d7_n97_d2", rotated_surface_code(7, 7, 1e-3))
    add("config3_surface_code_d7_n97_d2_noisy", rotated_surface_code(7, 7, 1e-3), boost=20.0, shot=1)
    add("config4_repetition_code_d25_n49_d3", qudit_repetition_code(25, 25, 3, 1e-2, "f"))
    add("config4_repetition_code_d25_n49_d3_noisy", qudit_repetition_code(25, 25, 3, 1e-2, "f"), boost=20.0, shot=1)
    add("headline_noisy_random_clifford_n256_d3", noisy_random_clifford(256, 2000, 3, seed=1, prob=1e-3, channel="d"))
    return out


def lanes_cases():
    """(name, n, d, ops, noise_ab) of one-shot cases for the uint8-LANE kernels at multi-word sizes (d = 5, 7, 11): the
    headline circuit family (noisy random Clifford + M on all qudits: the stream shape run_tail8_kernel takes) and
    streams with mid-circuit M / M_X / RESET in front of the final measurement.  Noise as in config_cases (Philox seed
    2026, shot 0, probabilities boosted so that events fire)."""
    from sdim_b200.ir import compile_circuits
    from sdim_b200.rng import noise_draws, prob_to_thresh24
    from sdim_b200.workloads import noisy_random_clifford
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN_DIR)))
    from make_cases import random_circuit
    out = []

    def add(name, circ, boost=1.0):
        prog = compile_circuits([circ])
        noise = np.zeros((0, 2), dtype=np.int64)
        if prog.n_noise:
            thr = np.array([prob_to_thresh24(min(1.0, boost * p)) for p in prog.noise_prob], dtype=np.uint32)
            noise = noise_draws(2026, prog.dimension, np.array([0]), thr, prog.noise_channel)[0].astype(np.int64)
        out.append((name, prog.num_qudits, prog.dimension, prog.ops.tolist(), noise))

    add("headline_shape_n100_d5", noisy_random_clifford(100, 700, 5, seed=3, prob=1e-3, channel="d"), boost=30.0)
    add("headline_shape_n160_d7", noisy_random_clifford(160, 900, 7, seed=4, prob=1e-3, channel="d"), boost=30.0)
    add("headline_shape_n256_d5", noisy_random_clifford(256, 1200, 5, seed=1, prob=1e-3, channel="d"), boost=30.0)
    add("mixed_stream_n97_d7", random_circuit(seed=71, n=97, d=7, depth=500))
    add("mixed_stream_n130_d11", random_circuit(seed=72, n=130, d=11, depth=500))
    return out


def main_configs(cases=None, filename="config_sizes.npz"):
    """One reference shot per BASELINE.json config size (n = 64, 97, 49, 256): pins both oracles, and through them
    every GPU mode, where round 1's goldens (n <= 13) could not see — multi-word lane rows, the headline shape.
    Stored compactly (uint8 / int32 arrays in one .npz); records rows are (qudit, deterministic, value).
    `--lanes` writes lanes_sizes.npz from lanes_cases() in the same format."""
    import time
    blob, names = {}, []
    for name, n, d, ops, noise in (cases if cases is not None else config_cases()):
        t0 = time.time()
        recs, arrs = rh.ref_run(n, d, ops, noise, draw_seed=2026)
        recs2, arrs2 = rh.ref_run_eager_modulo(n, d, ops, noise, draw_seed=2026)
        if recs != recs2 or any(not np.array_equal(arrs[k], arrs2[k]) for k in arrs):
            raise SystemExit(f"{name}: reference int64 overflow (lazy vs eager modulo disagree)")
        names.append(name)
        blob[name + "/nd"] = np.array([n, d], dtype=np.int32)
        blob[name + "/ops"] = np.array(ops, dtype=np.int32).reshape(-1, 4)
        blob[name + "/noise_ab"] = noise.astype(np.uint8).reshape(-1, 2)
        blob[name + "/records"] = np.array([[q, int(det), m] for q, det, m in recs], dtype=np.int16).reshape(-1, 3)
        for k, v in arrs.items():
            assert v.min() >= 0 and v.max() < 2 * d
            blob[name + "/" + k] = v.astype(np.uint8)
        n_rand = sum(1 for r in recs if not r[1])
        print(f"{name}: n={n} d={d} ops={len(ops)} noise fired={int((noise.sum(axis=1) > 0).sum())} "
              f"records={len(recs)} ({n_rand} random) {time.time() - t0:.1f}s")
    blob["names"] = np.array(names)
    path = os.path.join(GOLDEN_DIR, filename)
    np.savez_compressed(path, **blob)
    print(f"wrote {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    if "--configs" in sys.argv:
        main_configs()
    elif "--lanes" in sys.argv:
        main_configs(lanes_cases(), "lanes_sizes.npz")
    elif "--large" in sys.argv:
        main_large()
    elif "--wide" in sys.argv:
        main_wide()
    else:
        main()
