"""GPU: the drop-in API (Program.simulate) against the reference's own known-answer tests.

Cases restate reference tests/test_circuit.py (deterministic golden vectors through the tableau path),
tests/test_program.py:14-47 (result container shapes), tests/test_noise_and_io.py:101-200 and
tests/test_reset.py (noise statistics — here on the per-shot tableau path)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from sdim_b200 import Circuit, MeasurementResult, Program, read_circuit


def MR(q, det, v):
    return MeasurementResult(q, det, v)


def test_measurement_format():
    """reference tests/test_program.py:14-47."""
    c = Circuit(dimension=3, num_qudits=2); c.add_gate("M", [0, 1])
    p = Program(c)
    assert p.simulate(shots=1) == [MR(0, True, 0), MR(1, True, 0)]
    assert p.measurement_results == [[[MR(0, True, 0)]], [[MR(1, True, 0)]]]
    c = Circuit(dimension=3, num_qudits=2); c.add_gate("M", [0, 1]); c.add_gate("M", [0, 1])
    p = Program(c)
    assert p.simulate(shots=1) == [MR(0, True, 0), MR(0, True, 0), MR(1, True, 0), MR(1, True, 0)]
    assert p.measurement_results == [[[MR(0, True, 0)], [MR(0, True, 0)]], [[MR(1, True, 0)], [MR(1, True, 0)]]]
    c = Circuit(dimension=3, num_qudits=2); c.add_gate("M", [0, 1])
    p = Program(c)
    r = p.simulate(shots=2, force_tableau=True)
    assert r == [[[MR(0, True, 0), MR(0, True, 0)]], [[MR(1, True, 0), MR(1, True, 0)]]] == p.measurement_results


def test_phase_kickback_and_hpph():
    """tests/test_circuit.py:9-39."""
    c = Circuit(2, 2)
    c.add_gate("H", 0); c.add_gate("X", 1); c.add_gate("H", 1); c.add_gate("CNOT", 0, 1)
    c.add_gate("H", 0); c.add_gate("H", 1); c.add_gate("M", 0); c.add_gate("M", 1)
    assert Program(c).simulate() == [MR(0, True, 1), MR(1, True, 1)]
    c = Circuit(1, 2)
    for g in ("H", "P", "P", "H", "M"):
        c.add_gate(g, 0)
    assert Program(c).simulate() == [MR(0, True, 1)]
    c = Circuit(1, 3); c.add_gate("X", 0); c.add_gate("M", 0)
    assert Program(c).simulate() == [MR(0, True, 1)]


@pytest.mark.parametrize("d", [2, 3, 5, 7])
def test_swap_basis_states(d):
    """tests/test_circuit.py:48-94 (prime dimensions)."""
    c = Circuit(2, d)
    c.add_gate("X", 0); c.add_gate("SWAP", 0, 1); c.add_gate("M", [0, 1])
    assert Program(c).simulate() == [MR(0, True, 0), MR(1, True, 1)]
    c = Circuit(2, d)
    c.add_gate("X", 0); c.add_gate("SWAP", 0, 1); c.add_gate("SWAP", 0, 1); c.add_gate("M", [0, 1])
    assert Program(c).simulate() == [MR(0, True, 1), MR(1, True, 0)]


@pytest.mark.parametrize("i,j", [(a, b) for a in range(3) for b in range(3)])
def test_swap_in_x_basis_qutrit(i, j):
    """tests/test_circuit.py:96-136: prepare X-basis states, swap, measure with MX."""
    c = Circuit(2, 3)
    for _ in range(i):
        c.add_gate("X", 0)
    for _ in range(j):
        c.add_gate("X", 1)
    c.add_gate("DFT", 0); c.add_gate("DFT", 1)
    c.add_gate("SWAP", 0, 1)
    c.add_gate("MX", 0); c.add_gate("MX", 1)
    assert Program(c).simulate() == [MR(0, True, j), MR(1, True, i)]


def test_deutsch_qutrit_and_stabilizer_extraction():
    """tests/test_circuit.py:169-263."""
    # Z-type stabilizer extraction on qutrits
    c = Circuit(4, 3)
    c.add_gate("X", 0); c.add_gate("X", 1); c.add_gate("X", 1)
    c.add_gate("CNOT", 0, 2); c.add_gate("CNOT", 1, 2)
    c.add_gate("M", 2)
    assert Program(c).simulate() == [MR(2, True, 0)]
    # balanced-function Deutsch on qutrits: f(x) = x
    c = Circuit(2, 3)
    c.add_gate("X", 1); c.add_gate("H", 0); c.add_gate("H", 1)
    c.add_gate("CNOT", 0, 1)
    c.add_gate("H_INV", 0); c.add_gate("M", 0)
    res = Program(c).simulate()
    assert res[0].deterministic and res[0].measurement_value != 0


def test_shipped_circuits_1000_shots():
    """BASELINE config 1: epr.chp and css_steane_final.chp, d=2, 1k shots."""
    st = Program(read_circuit("circuits/css_steane_final.chp")).simulate(shots=1000, force_tableau=True)
    vals = [st[q][0] for q in range(7, 13)]
    for q, want in zip(range(7, 13), [1, 1, 0, 1, 1, 0]):
        assert all(r == MR(q, True, want) for r in st[q][0])
    assert all(st[q] == [] for q in range(7))
    epr = Program(read_circuit("circuits/epr.chp")).simulate(shots=1000, seed=7)
    ones = sum(r.measurement_value for r in epr[1][0])
    assert all(not r.deterministic for r in epr[1][0])
    assert abs(ones - 500) < 5 * (250 ** 0.5)                 # chi-square vs 50/50 at ~5 sigma


def test_epr_correlations_and_seeding():
    c = Circuit(2, 5)
    c.add_gate("H", 0); c.add_gate("CNOT", 0, 1); c.add_gate("M", [0, 1])
    p = Program(c)
    t = p.simulate_records(5000, seed=3)
    assert np.array_equal(t.values[:, 0], t.values[:, 1])
    assert not t.deterministic[:, 0].any() and t.deterministic[:, 1].all()
    hist = np.bincount(t.values[:, 0], minlength=5) / 5000
    assert np.abs(hist - 0.2).max() < 0.03
    assert np.array_equal(p.simulate_records(5000, seed=3).values, t.values)
    random.seed(123); a = p.simulate_records(100).values
    random.seed(123); b = p.simulate_records(100).values
    assert np.array_equal(a, b)                                 # follows the user's random.seed like the reference


@pytest.mark.parametrize("channel", ["f", "p", "d"])
def test_noise_channel_statistics(channel):
    """tests/test_noise_and_io.py:101-171: outcome histogram within 0.01 of the ideal at 1e5 shots."""
    rnd = random.Random(hash(channel) & 0xFFFF)
    d = rnd.choice([3, 5, 7, 11, 13, 17])
    shots = 100000
    c = Circuit(dimension=d, num_qudits=1)
    if channel == "d":
        p = rnd.uniform(0.0, (d * d - 1) / (d * d))
        c.add_gate("N1", 0, prob=p, noise_channel="d")
        ideal = [(1 - p) + (d - 1) * p / (d * d - 1)] + [d * p / (d * d - 1)] * (d - 1)
    else:
        p = rnd.uniform(0.0, 1.0)
        if channel == "p":
            c.add_gate("H", 0)
        c.add_gate("N1", 0, prob=p, noise_channel=channel)
        if channel == "p":
            c.add_gate("H_INV", 0)
        ideal = [1 - p] + [p / (d - 1)] * (d - 1)
    c.add_gate("M", 0)
    t = Program(c).simulate_records(shots, seed=11)
    emp = np.bincount(t.values[:, 0], minlength=d) / shots
    assert np.abs(emp - np.array(ideal)).max() < 0.01


def test_reset_after_max_mixing():
    """tests/test_reset.py: N1 (max mixing) then RESET then X^k then M -> always k."""
    for d in (3, 5, 7, 11):
        k = d - 2
        c = Circuit(dimension=d, num_qudits=1)
        c.add_gate("N1", 0, prob=(d * d - 1) / (d * d), noise_channel="d")
        c.add_gate("RESET", 0)
        for _ in range(k):
            c.add_gate("X", 0)
        c.add_gate("M", 0)
        t = Program(c).simulate_records(200000, seed=d)
        assert (t.values[:, 1] == k).all()
        # the RESET record itself is the physical pre-reset outcome: uniform under max mixing (unlike B-5)
        hist = np.bincount(t.values[:, 0], minlength=d) / 200000
        assert np.abs(hist - 1 / d).max() < 0.01


def test_random_sequence_times_inverse_is_identity():
    """tests/test_noise_and_io.py:174-200."""
    inv = {"H": "H_INV", "H_INV": "H", "P": "P_INV", "P_INV": "P", "Z": "Z_INV", "Z_INV": "Z", "X": "X_INV",
           "X_INV": "X", "I": "I"}
    rnd = random.Random(4)
    for d in (2, 3, 5):
        gates = [rnd.choice(list(inv)) for _ in range(200)]
        c = Circuit(dimension=d, num_qudits=1)
        for g in gates + [inv[g] for g in reversed(gates)]:
            c.add_gate(g, 0)
        c.add_gate("M", 0)
        t = Program(c).simulate_records(1000, seed=1)
        assert (t.values == 0).all() and t.deterministic.all()


def test_stabilizer_tableau_and_record_tableau():
    """`stabilizer_tableau` after simulate and `record_tableau=True` snapshots vs the oracle's tableau."""
    from make_cases import random_circuit
    from oracle.tableau_oracle import run_shot
    from sdim_b200.ir import compile_circuits
    from sdim_b200.rng import measurement_draws, noise_draws
    circ = random_circuit(21, 6, 3, 80)
    prog = compile_circuits([circ])
    P = Program(circ)
    res = P.simulate(shots=3, seed=17)
    last = P.stabilizer_tableau
    md = measurement_draws(17, 3, [2], prog.n_meas)[0]
    nd = noise_draws(17, 3, [2], prog.noise_thresh24, prog.noise_channel)[0]
    recs, t = run_shot(6, 3, prog.ops, lambda k: int(md[k]), nd)
    for got, want in zip((last.x_block, last.z_block, last.phase_vector, last.destab_x_block, last.destab_z_block,
                          last.destab_phase_vector), t.arrays()):
        assert np.array_equal(got, want)
    flat = sorted((q, r, res[q][r][2].deterministic, res[q][r][2].measurement_value)
                  for q in range(6) for r in range(len(res[q])))
    cnt, want_flat = {}, []
    for q, det, m in recs:
        want_flat.append((q, cnt.get(q, 0), det, m)); cnt[q] = cnt.get(q, 0) + 1
    assert flat == sorted(want_flat)
    # record_tableau: the last measurement's snapshot equals the final tableau
    P2 = Program(circ)
    one = P2.simulate(shots=1, record_tableau=True, seed=17)
    assert all(r.stabilizer_tableau is not None for r in one)
    final = P2.stabilizer_tableau
    last_meas_q = int(prog.meas_qudit[-1])
    snap = [r for r in one if r.qudit_index == last_meas_q][-1].get_tableau()
    assert np.array_equal(snap.x_block, final.x_block) and np.array_equal(snap.phase_vector, final.phase_vector)


def test_initial_tableau_injection_and_apply_gate():
    from sdim_b200 import ExtendedTableau
    t0 = ExtendedTableau(2, 3)
    # start from |1,0>: stabilizer Z_0 with phase such that outcome 1 -> phase = -1*po mod order = 2
    t0.phase_vector[0] = 2
    c = Circuit(2, 3); c.add_gate("M", [0, 1])
    assert Program(c, tableau=t0).simulate() == [MR(0, True, 1), MR(1, True, 0)]
    p = Program(Circuit(2, 3))
    from sdim_b200 import CircuitInstruction
    gd = p.circuits[0].gate_data
    assert p.apply_gate(CircuitInstruction(gd, "X", 0)) is None
    assert p.apply_gate(CircuitInstruction(gd, "X", 0)) is None
    r = p.apply_gate(CircuitInstruction(gd, "M", 0))
    assert r == MR(0, True, 2)


def test_verbose_and_show_gate_run(capsys):
    c = Circuit(2, 2); c.add_gate("H", 0); c.add_gate("CNOT", 0, 1); c.add_gate("M", [0, 1])
    res = Program(c).simulate(shots=1, verbose=True, show_gate=True, show_measurement=True, seed=5)
    out = capsys.readouterr().out
    assert "Initial state" in out and "Final step 3" in out and "Destabilizer X Block" in out
    assert "Measured qudit (0)" in out
    assert res[0].measurement_value == res[1].measurement_value and res[1].deterministic


def test_extended_tableau_gate_methods_match_oracle():
    """The reference's per-gate interface on ExtendedTableau (tableau_prime.py:97-292), each call one device op."""
    from oracle.tableau_oracle import OracleTableau
    from sdim_b200 import ExtendedTableau
    for d in (2, 3, 5):
        t, o = ExtendedTableau(3, d), OracleTableau(3, d)
        t.hadamard(0); o.hadamard(0)
        t.cnot(0, 1); o.cnot(0, 1)
        t.phase(1); o.phase(1)
        t.hadamard_inv(2); o.hadamard(2, inverse=True)
        t.cnot_inv(2, 0); o.cnot(2, 0, inverse=True)
        t.phase_inv(0); o.phase(0, inverse=True)
        for got, want in zip((t.x_block, t.z_block, t.phase_vector, t.destab_x_block, t.destab_z_block,
                              t.destab_phase_vector), o.arrays()):
            assert np.array_equal(got, want)
        r = t.measure(1)
        det, m = o.measure(1, lambda: r.measurement_value)
        assert (r.qudit_index, r.deterministic) == (1, det) and r.measurement_value == m
        for got, want in zip((t.x_block, t.z_block, t.phase_vector, t.destab_x_block, t.destab_z_block,
                              t.destab_phase_vector), o.arrays()):
            assert np.array_equal(got, want)


def test_many_waves_pipeline_gives_the_same_records():
    """Large shot counts run in waves whose records leave the device through pinned double buffers on a copy stream
    (Program._run_local); forcing tiny waves must not change a single record, in resident and HBM-store modes."""
    from sdim_b200 import Program
    from make_cases import random_circuit
    circ = random_circuit(seed=9, n=11, d=3, depth=150)
    prog = Program(circ)
    want = prog.simulate_records(1000, seed=4).values
    want_det = prog.last_records.deterministic
    small = Program(circ)
    small.WAVE_RECORD_BYTES = 64 * small._compiled().n_meas          # 64 shots per wave: 16 waves, last one ragged
    got = small.simulate_records(1000, seed=4)
    assert np.array_equal(got.values, want) and np.array_equal(got.deterministic, want_det)
    got = small.simulate_records(130, seed=4, shot_offset=870)
    assert np.array_equal(got.values, want[870:])


@pytest.mark.parametrize("d,n", [(3, 140), (5, 100)])
def test_waves_are_sized_by_the_per_shot_device_state(d, n):
    """The two-kernel paths keep device state per shot of a wave (one generator-major slab per shot for d = 2, 3; one HBM
    tableau per shot for uint8 lanes): Program sizes its waves by it (WAVE_DEVICE_BYTES), and the records do not
    depend on the wave size — checked against the C oracle through the public API."""
    from oracle import c_oracle
    from sdim_b200 import Program
    from make_cases import random_circuit
    circ = random_circuit(seed=3 * n + d, n=n, d=d, depth=300, p_meas=0.0)
    prog = Program(circ)
    compiled = prog._compiled()
    engine = prog._get_engine(compiled)
    per = engine.device_bytes_per_shot(500)
    assert per >= (10_000 if d == 3 else engine.layout.shot_bytes)
    want = c_oracle.run_philox(compiled, 500, 0, 6)

    def packed(table):
        return table.values | (table.deterministic.astype(np.uint8) << 7)

    assert np.array_equal(packed(prog.simulate_records(500, seed=6)), want)
    small = Program(circ)
    small.WAVE_DEVICE_BYTES = 70 * (per + 5 * compiled.n_meas)       # about 70 shots per wave
    assert np.array_equal(packed(small.simulate_records(500, seed=6)), want)


def test_fold_gates_gives_the_same_records():
    """Program(fold_gates=True) uploads the peephole-folded stream (sdim_b200/peephole.py): same records, same
    deterministic flags, same final tableau under the same seed."""
    from sdim_b200 import Program
    from test_peephole import _redundant_circuit
    for d, n in ((2, 9), (3, 7), (5, 4)):
        circ = _redundant_circuit(n, d, 200, 40 + d)
        plain, folded = Program(circ), Program(circ, fold_gates=True)
        assert folded._compiled(fold=True).n_ops < plain._compiled().n_ops
        a = plain.simulate_records(300, seed=8)
        b = folded.simulate_records(300, seed=8)
        assert np.array_equal(a.values, b.values) and np.array_equal(a.deterministic, b.deterministic)
        ta, tb = plain.stabilizer_tableau, folded.stabilizer_tableau
        for key in ("x_block", "z_block", "phase_vector", "destab_x_block", "destab_z_block", "destab_phase_vector"):
            assert np.array_equal(getattr(ta, key), getattr(tb, key)), key


def test_record_tableau_snapshot_of_a_reset_precedes_its_correction():
    """The reference snapshots a RESET right after measure(), before the X^(-m) correction
    (sdim/program.py:323-324 vs :335-339): the snapshot must equal the oracle's tableau at that point."""
    from oracle.tableau_oracle import run_shot
    from sdim_b200.ir import compile_circuits
    from sdim_b200.rng import measurement_draws
    for d in (2, 3, 5):
        c = Circuit(3, d)
        c.add_gate("H", 0); c.add_gate("CNOT", 0, 1); c.add_gate("P", 1); c.add_gate("H", 2); c.add_gate("CZ", 2, 1)
        head = compile_circuits([c])
        c.add_gate("RESET", 1)
        prog = compile_circuits([c])
        for seed in range(6):
            res = Program(c).simulate(shots=1, record_tableau=True, seed=seed)
            md = measurement_draws(seed, d, [0], prog.n_meas)[0]
            _, t = run_shot(3, d, head.ops, lambda k: 0)
            det, value = t.measure(1, lambda: int(md[0]))
            assert (res[0].deterministic, res[0].measurement_value) == (det, value)
            snap = res[0].get_tableau()
            for got, want in zip((snap.x_block, snap.z_block, snap.phase_vector, snap.destab_x_block,
                                  snap.destab_z_block, snap.destab_phase_vector), t.arrays()):
                assert np.array_equal(got, want), (d, seed)
