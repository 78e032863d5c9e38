"""CPU: both oracles (numpy restatement and C restatement) against the reference's golden vectors.

The golden files were produced by the unmodified reference (oracle/make_golden.py); these tests are what
pins the oracle.  When /root/reference is mounted the oracle is additionally cross-checked live.
"""
import random

import numpy as np
import pytest

from oracle import c_oracle, ref_harness
from oracle.tableau_oracle import run_shot, run_shots

KEYS = ("x", "z", "p", "dx", "dz", "dp")


def _want(case):
    return np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in case["records"]], dtype=np.uint8)


def test_golden_file_is_substantial(golden_random):
    assert len(golden_random) >= 100
    dims = {c["d"] for c in golden_random}
    assert {2, 3, 5, 7, 11, 13} <= dims
    n_rand = sum(1 for c in golden_random for r in c["records"] if not r[1])
    n_det = sum(1 for c in golden_random for r in c["records"] if r[1])
    assert n_rand > 500 and n_det > 500
    kinds = {op[0] for c in golden_random for op in c["ops"]}
    assert kinds == set(range(18)), "every opcode (I..N1) must appear in the goldens"


def test_numpy_oracle_matches_reference_goldens(golden_random):
    for case in golden_random:
        draws = [r[2] for r in case["records"]]
        noise = np.array(case["noise_ab"], dtype=np.int64).reshape(-1, 2)
        recs, t = run_shot(case["n"], case["d"], case["ops"], lambda k: draws[k], noise)
        assert recs == [(q, bool(det), m) for q, det, m in case["records"]], case["seed"]
        for key, arr in zip(KEYS, t.arrays()):
            assert np.array_equal(arr, np.array(case["final"][key])), (case["seed"], key)


@pytest.mark.skipif(not c_oracle.available(), reason="oracle/liboracle.so not built (run `make -C oracle`)")
def test_c_oracle_matches_reference_goldens(golden_random):
    for case in golden_random:
        want = _want(case)
        noise = np.array(case["noise_ab"], dtype=np.uint8).reshape(1, -1, 2)
        rec, fin = c_oracle.run(case["n"], case["d"], case["ops"], 1, replay_meas=(want & 0x7F)[None, :],
                                replay_noise=noise, want_final=True)
        assert np.array_equal(rec[0], want), case["seed"]
        for key in KEYS:
            assert np.array_equal(fin[key], np.array(case["final"][key])), (case["seed"], key)


def test_oracles_match_reference_goldens_at_large_primes(golden_large_primes):
    """d in {17, 31, 61, 127} (127 = the uint8-lane limit of the CUDA store): both restatements against outputs of the
    unmodified reference (`oracle/make_golden.py --large`).  The GPU parity tests at these dimensions compare with the
    C oracle, which this test pins."""
    assert {c["d"] for c in golden_large_primes} == {17, 31, 61, 127} and len(golden_large_primes) >= 48
    assert sum(1 for c in golden_large_primes for r in c["records"] if not r[1]) > 80
    for case in golden_large_primes:
        draws = [r[2] for r in case["records"]]
        noise64 = np.array(case["noise_ab"], dtype=np.int64).reshape(-1, 2)
        recs, t = run_shot(case["n"], case["d"], case["ops"], lambda k: draws[k], noise64)
        assert recs == [(q, bool(det), m) for q, det, m in case["records"]], case["seed"]
        for key, arr in zip(KEYS, t.arrays()):
            assert np.array_equal(arr, np.array(case["final"][key])), (case["seed"], key)
        if c_oracle.available():
            want = _want(case)
            rec, fin = c_oracle.run(case["n"], case["d"], case["ops"], 1, replay_meas=(want & 0x7F)[None, :],
                                    replay_noise=noise64.astype(np.uint8).reshape(1, -1, 2), want_final=True)
            assert np.array_equal(rec[0], want), case["seed"]
            for key in KEYS:
                assert np.array_equal(fin[key], np.array(case["final"][key])), (case["seed"], key)


def test_oracles_match_reference_goldens_above_127(golden_wide_primes):
    """d in {131, 251, 257, 1031, 32749}: the dimensions of the uint16-lane store (records and replay arrays are
    uint16 there, bit 15 = deterministic).  Both restatements against the unmodified reference
    (`oracle/make_golden.py --wide`); the GPU tests of csrc/wide.cuh compare with these goldens and the C oracle."""
    assert {c["d"] for c in golden_wide_primes} == {131, 251, 257, 1031, 32749} and len(golden_wide_primes) >= 55
    assert sum(1 for c in golden_wide_primes for r in c["records"] if not r[1]) > 80
    assert max(r[2] for c in golden_wide_primes for r in c["records"]) > 255        # values that need the second byte
    for case in golden_wide_primes:
        draws = [r[2] for r in case["records"]]
        noise64 = np.array(case["noise_ab"], dtype=np.int64).reshape(-1, 2)
        recs, t = run_shot(case["n"], case["d"], case["ops"], lambda k: draws[k], noise64)
        assert recs == [(q, bool(det), m) for q, det, m in case["records"]], case["seed"]
        for key, arr in zip(KEYS, t.arrays()):
            assert np.array_equal(arr, np.array(case["final"][key])), (case["seed"], key)
        want = np.array([(m & 0x7FFF) | (0x8000 if det else 0) for _, det, m in case["records"]], dtype=np.uint16)
        packed, _ = run_shots(case["n"], case["d"], case["ops"], 1, np.array(draws)[None, :], noise64[None])
        assert packed.dtype == np.uint16 and np.array_equal(packed[0], want), case["seed"]
        if c_oracle.available():
            rec, fin = c_oracle.run(case["n"], case["d"], case["ops"], 1, replay_meas=(want & 0x7FFF)[None, :],
                                    replay_noise=noise64.astype(np.uint16).reshape(1, -1, 2), want_final=True)
            assert rec.dtype == np.uint16 and np.array_equal(rec[0], want), case["seed"]
            for key in KEYS:
                assert np.array_equal(fin[key], np.array(case["final"][key])), (case["seed"], key)


def test_oracles_match_reference_goldens_at_config_sizes(golden_config_sizes):
    """One reference shot per BASELINE.json config size — config 2 (n = 64, d = 3), the distance-7 surface code
    (n = 97, d = 2), the distance-25 qutrit repetition code (n = 49), the headline (n = 256, d = 3; N1 events = shot 0
    of Philox seed 2026): records and all six final arrays of both restatements.  This is what pins the C oracle at the
    sizes where the GPU parity tests lean on it (multi-word lane rows, 4-warp CTAs, the slab image)."""
    names = [c["name"] for c in golden_config_sizes]
    assert len(names) == 6 and {c["n"] for c in golden_config_sizes} == {49, 64, 97, 256}
    assert sum(int((c["records"][:, 1] == 0).sum()) for c in golden_config_sizes) > 300
    for case in golden_config_sizes:
        n, d, ops = case["n"], case["d"], case["ops"]
        want = np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in case["records"]], dtype=np.uint8)
        noise = case["noise_ab"]
        if c_oracle.available():
            rec, fin = c_oracle.run(n, d, ops, 1, replay_meas=(want & 0x7F)[None, :],
                                    replay_noise=noise.reshape(1, -1, 2), want_final=True)
            assert np.array_equal(rec[0], want), case["name"]
            for key in KEYS:
                assert np.array_equal(fin[key], case["final"][key]), (case["name"], key)
        if n <= 97:       # the numpy restatement at n = 256 is covered through the C oracle (pinned to it at n <= 33 and here)
            draws = [int(r[2]) for r in case["records"]]
            recs, t = run_shot(n, d, ops.tolist(), lambda k: draws[k], noise.astype(np.int64))
            assert recs == [(int(q), bool(det), int(m)) for q, det, m in case["records"]], case["name"]
            for key, arr in zip(KEYS, t.arrays()):
                assert np.array_equal(arr, case["final"][key]), (case["name"], key)


def test_oracles_match_reference_goldens_for_the_lane_kernels(golden_lanes_sizes):
    """One reference shot per case at d = 5, 7, 11 and n = 97 ... 256 (oracle/make_golden.py --lanes): the headline
    circuit family in the dimensions the uint8 lanes serve, and streams with mid-circuit M / M_X / RESET.  Pins the C
    oracle (and the numpy one at n <= 100) where the GPU tests of the lane interpreter and of run_tail8_kernel lean on it."""
    assert {(c["n"], c["d"]) for c in golden_lanes_sizes} == {(100, 5), (160, 7), (256, 5), (97, 7), (130, 11)}
    assert sum(int((c["records"][:, 1] == 0).sum()) for c in golden_lanes_sizes) > 300      # random measurements
    for case in golden_lanes_sizes:
        n, d, ops = case["n"], case["d"], case["ops"]
        want = np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in case["records"]], dtype=np.uint8)
        noise = case["noise_ab"]
        assert int((noise.sum(axis=1) > 0).sum()) >= 10
        if c_oracle.available():
            rec, fin = c_oracle.run(n, d, ops, 1, replay_meas=(want & 0x7F)[None, :],
                                    replay_noise=noise.reshape(1, -1, 2), want_final=True)
            assert np.array_equal(rec[0], want), case["name"]
            for key in KEYS:
                assert np.array_equal(fin[key], case["final"][key]), (case["name"], key)
        if n <= 100:
            draws = [int(r[2]) for r in case["records"]]
            recs, t = run_shot(n, d, ops.tolist(), lambda k: draws[k], noise.astype(np.int64))
            assert recs == [(int(q), bool(det), int(m)) for q, det, m in case["records"]], case["name"]
            for key, arr in zip(KEYS, t.arrays()):
                assert np.array_equal(arr, case["final"][key]), (case["name"], key)


def test_shipped_circuit_goldens(golden_shipped):
    """circuits/css_steane_final.chp -> 1,1,0,1,1,0 all deterministic; circuits/epr.chp -> qudit 1 random."""
    st = golden_shipped["circuits/css_steane_final.chp"]
    assert st["num_qudits"] == 13 and st["dimension"] == 2
    assert [r[2] for r in st["flat_results"]] == [1, 1, 0, 1, 1, 0]
    assert all(r[1] == 1 for r in st["flat_results"])
    for name, g in golden_shipped.items():
        ops, k = [], 0
        for op, a, b in g["ops"]:
            slot = -1
            if op in (14, 15, 16):
                slot, k = k, k + 1
            ops.append([op, a, b, slot])
        draws = {}
        # flat results are ordered by (qudit, round); every qudit is measured at most once in these files
        for q, det, m in g["flat_results"]:
            draws[q] = m
        meas_q = [o[1] for o in ops if o[0] in (14, 15, 16)]
        recs, t = run_shot(g["num_qudits"], g["dimension"], ops, lambda kk: draws[meas_q[kk]])
        got = sorted((q, int(det), m) for q, det, m in recs)
        assert got == sorted(tuple(r) for r in g["flat_results"]), name
        for key, arr in zip(KEYS, t.arrays()):
            assert np.array_equal(arr, np.array(g["final"][key])), (name, key)


@pytest.mark.skipif(not c_oracle.available(), reason="liboracle.so not built")
@pytest.mark.parametrize("d,n", [(2, 20), (3, 33), (5, 17), (7, 12), (13, 9)])
def test_numpy_and_c_oracle_agree_free_running(d, n):
    """Independent implementations, Philox draws: records of several shots must coincide."""
    from make_cases import random_program
    prog = random_program(seed=100 + d, n=n, d=d, depth=30 * n)
    from sdim_b200.rng import measurement_draws, noise_draws
    shots = 4
    ids = np.arange(shots)
    md = measurement_draws(77, d, ids, prog.n_meas)
    nd = noise_draws(77, d, ids, prog.noise_thresh24, prog.noise_channel)
    a, _ = run_shots(n, d, prog.ops, shots, md, nd)
    b = c_oracle.run_philox(prog, shots, 0, 77)
    assert np.array_equal(a, b)
    assert ((a & 0x80) == 0).any() and ((a & 0x80) != 0).any()


@pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not mounted")
def test_oracle_against_live_reference():
    """Fresh random circuits through the real reference, beyond the committed goldens."""
    import warnings
    warnings.simplefilter("ignore")
    from oracle.make_golden import random_ops, final_measure_all
    rng = random.Random(424242)
    checked = 0
    for d in (2, 3, 5, 7):
        for n in (2, 4, 6):
            ops, k, j = random_ops(rng, n, 25 * n, d)
            k = final_measure_all(ops, n, k)
            noise = np.array([[rng.randrange(d), rng.randrange(d)] for _ in range(j)], dtype=np.int64).reshape(-1, 2)
            recs, arrs = ref_harness.ref_run(n, d, ops, noise, draw_seed=d * 100 + n)
            recs2, arrs2 = ref_harness.ref_run_eager_modulo(n, d, ops, noise, draw_seed=d * 100 + n)
            if recs != recs2 or any(not np.array_equal(arrs[key], arrs2[key]) for key in arrs):
                continue    # reference int64 overflow
            draws = [r[2] for r in recs]
            got, t = run_shot(n, d, ops, lambda kk: draws[kk], noise)
            assert got == recs
            for key, arr in zip(KEYS, t.arrays()):
                assert np.array_equal(arr, arrs[key])
            checked += 1
    assert checked >= 10


def test_frame_oracle_matches_reference_simulate_frame():
    """oracle/frame_oracle.py vs outputs of the unmodified reference simulate_frame (sdim/program.py:45-165),
    its np.random draws replayed (oracle/make_golden_frames.py)."""
    import json
    import os
    from conftest import GOLDEN_DIR
    from oracle.frame_oracle import simulate_frames
    cases = json.load(open(os.path.join(GOLDEN_DIR, "frame_cases.json")))["cases"]
    assert len(cases) >= 16 and {c["d"] for c in cases} >= {2, 3, 5, 7}
    n_reset = 0
    for c in cases:
        shots = len(c["z0"])
        noise = np.array(c["noise_ab"], dtype=np.int64).reshape(shots, -1, 2) if c["noise_ab"] else None
        got = simulate_frames(c["n"], c["d"], c["ops"], c["reference"], c["z0"], np.array(c["zm"]), noise,
                              reset_records="reference")
        assert np.array_equal(got, np.array(c["records"])), c["seed"]
        n_reset += sum(1 for o in c["ops"] if o[0] == 16)
    assert n_reset > 20
