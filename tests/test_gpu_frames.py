"""GPU: the Pauli-frame sampler (sdimb_frames) — the reference's default multi-shot path (sdim/program.py:45-165,
244-265) — against the reference's golden vectors, the frame oracle and the per-shot tableau path."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import GOLDEN_DIR
from helpers import circuit_from_ops


def _engine(n, d, ops):
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    prog = compile_circuits([circuit_from_ops(n, d, ops)])
    return prog, TableauEngine(prog)


def test_frames_replay_against_reference_goldens():
    """Replayed draws: every M / M_X record equals the reference's simulate_frame output; RESET records are the
    physical outcome (frame oracle, reset_records='physical') instead of the reference's stale value (B-5)."""
    import torch
    from oracle.frame_oracle import simulate_frames
    cases = json.load(open(os.path.join(GOLDEN_DIR, "frame_cases.json")))["cases"]
    for c in cases:
        n, d, ops = c["n"], c["d"], c["ops"]
        prog, eng = _engine(n, d, ops)
        shots = len(c["z0"])
        noise = np.array(c["noise_ab"], dtype=np.uint8).reshape(shots, -1, 2) if c["noise_ab"] else None
        got = eng.run_frames(shots, torch.tensor(c["reference"], dtype=torch.uint8), 1, 0,
                             torch.tensor(c["z0"], dtype=torch.uint8), torch.tensor(c["zm"], dtype=torch.uint8),
                             None if noise is None else torch.from_numpy(noise)).cpu().numpy()
        ref = np.array(c["records"], dtype=np.uint8)
        is_reset = np.array([o[0] == 16 for o in ops if o[0] in (14, 15, 16)])
        assert np.array_equal(got[:, ~is_reset], ref[:, ~is_reset]), c["seed"]
        phys = simulate_frames(n, d, ops, c["reference"], c["z0"], np.array(c["zm"]),
                               None if noise is None else noise.astype(np.int64), reset_records="physical")
        assert np.array_equal(got, phys), c["seed"]


@pytest.mark.parametrize("d,n,depth", [(2, 30, 600), (3, 64, 1500), (5, 21, 500), (13, 7, 200)])
def test_frames_philox_mode_matches_frame_oracle(d, n, depth):
    from make_cases import random_program
    from oracle.frame_oracle import simulate_frames
    from sdim_b200.engine import TableauEngine
    from sdim_b200.rng import frame_z0_draws, frame_zm_draws, noise_draws
    import torch
    prog = random_program(seed=7 * d + n, n=n, d=d, depth=depth)
    eng = TableauEngine(prog)
    seed, shots = 99, 500
    quiet = torch.zeros((1, prog.n_noise, 2), dtype=torch.uint8) if prog.n_noise else None
    ref = eng.run(1, 0, seed, None, quiet)
    got = eng.run_frames(shots, ref[0], 1, seed).cpu().numpy()
    ids = np.arange(1, shots + 1)
    want = simulate_frames(n, d, prog.ops, ref[0].cpu().numpy(), frame_z0_draws(seed, d, ids, n),
                           frame_zm_draws(seed, d, ids, prog.n_meas),
                           noise_draws(seed, d, ids, prog.noise_thresh24, prog.noise_channel).astype(np.int64)
                           if prog.n_noise else None)
    assert np.array_equal(got, want)


def test_program_frame_method_shapes_and_reference_shot():
    from sdim_b200 import Circuit, MeasurementResult, Program
    c = Circuit(3, 3)
    c.add_gate("H", 0); c.add_gate("CNOT", 0, 1); c.add_gate("N1", 2, prob=1.0, noise_channel="f")
    c.add_gate("M", [0, 1, 2])
    p = Program(c)
    res = p.simulate(shots=2000, method="frame", seed=4)
    assert len(res) == 3 and len(res[0]) == 1 and len(res[0][0]) == 2000
    # shot 0 is the noiseless reference shot (program.py:245-247, B-6): qudit 2 reads 0 there, never in the others
    assert res[2][0][0] == MeasurementResult(2, True, 0)
    assert all(r.measurement_value != 0 for r in res[2][0][1:])
    # EPR correlations survive in every frame shot; qudit 0 uniform
    v0 = np.array([r.measurement_value for r in res[0][0]]); v1 = np.array([r.measurement_value for r in res[1][0]])
    assert np.array_equal(v0, v1)
    assert np.abs(np.bincount(v0, minlength=3) / 2000 - 1 / 3).max() < 0.05
    # force_tableau overrides the frame method
    t = p.simulate(shots=50, method="frame", force_tableau=True, seed=4)
    assert all(r.measurement_value != 0 for r in t[2][0])


def test_frame_and_tableau_methods_agree_in_distribution():
    """Noisy qutrit repetition-code round: per-record marginals of the frame sampler match the per-shot tableaus
    (RESET records included, thanks to the physical RESET record)."""
    from sdim_b200 import Program
    from sdim_b200.workloads import qudit_repetition_code
    circ = qudit_repetition_code(5, 3, 3, prob=0.15)
    shots = 60000
    a = Program(circ).simulate_records(shots, seed=1, method="tableau")
    b = Program(circ).simulate_records(shots, seed=2, method="frame")
    for k in range(a.values.shape[1]):
        ha = np.bincount(a.values[:, k], minlength=3) / shots
        hb = np.bincount(b.values[1:, k], minlength=3) / (shots - 1)
        assert np.abs(ha - hb).max() < 0.012, k
    # a pair correlation too: syndromes of neighbouring ancillas in round 0
    ja = np.bincount(a.values[:, 0] * 3 + a.values[:, 1], minlength=9) / shots
    jb = np.bincount(b.values[1:, 0] * 3 + b.values[1:, 1], minlength=9) / (shots - 1)
    assert np.abs(ja - jb).max() < 0.012


def test_reference_reset_and_noise_tests_through_frames():
    """reference tests/test_reset.py and tests/test_noise_and_io.py:142-163 run on its frame path; same here."""
    from sdim_b200 import Circuit, Program
    for d in (3, 5, 7):
        k = d - 1
        c = Circuit(dimension=d, num_qudits=1)
        c.add_gate("N1", 0, prob=(d * d - 1) / (d * d), noise_channel="d")
        c.add_gate("RESET", 0)
        for _ in range(k):
            c.add_gate("X", 0)
        c.add_gate("M", 0)
        t = Program(c).simulate_records(100000, seed=d, method="frame")
        assert (t.values[:, 1] == k).all()
    d, p = 5, 0.4
    c = Circuit(dimension=d, num_qudits=1)
    c.add_gate("N1", 0, prob=p, noise_channel="d"); c.add_gate("M", 0)
    t = Program(c).simulate_records(100000, seed=3, method="frame")
    emp = np.bincount(t.values[1:, 0], minlength=d) / 99999
    ideal = [(1 - p) + (d - 1) * p / (d * d - 1)] + [d * p / (d * d - 1)] * (d - 1)
    assert np.abs(emp - np.array(ideal)).max() < 0.01
