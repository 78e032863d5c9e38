"""CPU: the host logic of sdim_b200.program on a stand-in engine (tests/fake_engine.py, numpy oracle inside).

The device work is replaced, everything above it is the product code: result grouping (sdim/program.py:321-365),
host-stepped modes, record_tableau snapshots, apply_gate, the lazily materialised last-shot tableau, gate folding.
The GPU suite runs the same scenarios on the real engine (tests/test_gpu_program.py)."""
import numpy as np
import pytest

import sdim_b200.engine as engine_mod
from fake_engine import FakeEngine
from make_cases import random_circuit
from oracle.tableau_oracle import run_shot
from sdim_b200 import Circuit, CircuitInstruction, ExtendedTableau, MeasurementResult, Program
from sdim_b200.ir import compile_circuits
from sdim_b200.rng import measurement_draws, noise_draws

KEYS = ("x_block", "z_block", "phase_vector", "destab_x_block", "destab_z_block", "destab_phase_vector")


@pytest.fixture(autouse=True)
def fake_engine(monkeypatch):
    monkeypatch.setattr(engine_mod, "TableauEngine", FakeEngine)
    import torch
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda device=None: (1 << 34, 1 << 34))


def oracle_shot(prog, seed, shot):
    md = measurement_draws(seed, prog.dimension, [shot], prog.n_meas)[0]
    nd = noise_draws(seed, prog.dimension, [shot], prog.noise_thresh24, prog.noise_channel)[0] if prog.n_noise else None
    return run_shot(prog.num_qudits, prog.dimension, prog.ops, lambda k: int(md[k]), nd)


def same_tableau(got: ExtendedTableau, oracle_tableau):
    for key, want in zip(KEYS, oracle_tableau.arrays()):
        assert np.array_equal(getattr(got, key), want), key


@pytest.mark.parametrize("d", [2, 3, 5, 257])
def test_simulate_shapes_grouping_and_last_tableau(d):
    circ = random_circuit(31 + d, 5, d, 70)
    prog = compile_circuits([circ])
    P = Program(circ)
    flat = P.simulate(shots=1, seed=9)                       # flat list ordered by (qudit, round)  program.py:357-363
    recs, t = oracle_shot(prog, 9, 0)
    by_qudit = {}
    for q, det, m in recs:
        by_qudit.setdefault(q, []).append(MeasurementResult(q, det, m))
    assert flat == [r for q in sorted(by_qudit) for r in by_qudit[q]]
    same_tableau(P.stabilizer_tableau, t)
    nested = P.simulate(shots=4, seed=9)                     # [qudit][round][shot]                 program.py:364-365
    assert len(nested) == 5 and all(len(group) == 4 for per_q in nested for group in per_q)
    for s in range(4):
        recs, t = oracle_shot(prog, 9, s)
        seen = {}
        for q, det, m in recs:
            r = seen.get(q, 0)
            seen[q] = r + 1
            assert nested[q][r][s] == MeasurementResult(q, det, m)
    same_tableau(P.stabilizer_tableau, t)                    # tableau of the LAST shot
    assert P.measurement_results is nested


@pytest.mark.parametrize("d", [3, 257])
def test_record_tableau_snapshots_including_reset(capsys, d):
    circ = random_circuit(77, 4, d, 60, p_meas=0.2)
    prog = compile_circuits([circ])
    assert 16 in prog.meas_opcode                            # the circuit has RESETs
    res = Program(circ).simulate(shots=2, record_tableau=True, seed=5)
    for s in range(2):
        md = measurement_draws(5, d, [s], prog.n_meas)[0]
        nd = noise_draws(5, d, [s], prog.noise_thresh24, prog.noise_channel)[0]
        seen, t = {}, None
        for i, (op, a, b, slot) in enumerate(prog.ops.tolist()):
            # RESET: run its measurement only — the snapshot sits between measure() and the correction
            _, t = run_shot(4, d, [[14 if op == 16 else op, a, b, slot]], lambda k: int(md[k]), nd, tableau=t)
            snap_want = [arr.copy() for arr in t.arrays()]
            if op in (14, 15, 16):
                r = seen.get(a, 0)
                seen[a] = r + 1
                got = res[a][r][s]
                snap = got.get_tableau()
                for key, want in zip(KEYS, snap_want):
                    assert np.array_equal(getattr(snap, key), want), (s, i, key)
                if op == 16:                                 # now the correction, to carry on
                    t.pauli(a, (-got.measurement_value) % d, 0)


def test_verbose_and_show_gate_step_every_op(capsys):
    c = Circuit(2, 2); c.add_gate("H", 0); c.add_gate("I", 1); c.add_gate("CNOT", 0, 1); c.add_gate("M", [0, 1])
    P = Program(c)
    res = P.simulate(shots=1, verbose=True, show_gate=True, show_measurement=True, seed=5)
    out = capsys.readouterr().out
    # the reference's numbering (sdim/program.py:311-346): `time` counts I gates; "Final step" at time = total ops - 1
    assert "Initial state" in out and "Time step 0 \t H 0" in out and "Time step 1 \t I 1" in out and "Time step 2 \t CNOT 0 1" in out
    assert "Time step 3 \t M 0" in out and "Final step 4 \t M 1" in out and "Measured qudit (0)" in out
    assert res[0].measurement_value == res[1].measurement_value and res[1].deterministic and not res[0].deterministic
    assert [r[1] for r in P._engine.runs] == [(i, i + 1) for i in range(4)]     # one launch per op, I not launched


def test_show_gate_numbering_restarts_in_every_appended_circuit(capsys):
    a = Circuit(2, 3); a.add_gate("H", 0); a.add_gate("CNOT", 0, 1)
    b = Circuit(2, 3); b.add_gate("I", 0); b.add_gate("M", 1)
    P = Program(a)
    P.append_circuit(b)
    P.simulate(shots=1, verbose=True, show_gate=True, seed=1)
    out = capsys.readouterr().out
    lines = [ln for ln in out.splitlines() if "step" in ln]
    # `time` restarts with each circuit; with 4 operations in total only time = 3 would be the "Final step" (reference :341-344)
    assert [ln.split("\t")[0].strip() for ln in lines] == ["Time step 0", "Time step 1", "Time step 0", "Time step 1"]
    assert out.count("Initial state") == 2                                      # once per circuit (reference :313-316)


def test_fold_gates_uploads_the_folded_stream_and_changes_nothing():
    from test_peephole import _redundant_circuit
    circ = _redundant_circuit(5, 3, 160, 3)
    plain, folded = Program(circ), Program(circ, fold_gates=True)
    a, b = plain.simulate_records(6, seed=2), folded.simulate_records(6, seed=2)
    assert np.array_equal(a.values, b.values) and np.array_equal(a.deterministic, b.deterministic)
    assert folded._engine.runs[0][2] < plain._engine.runs[0][2]                 # fewer ops reached the engine
    for key in KEYS:
        assert np.array_equal(getattr(plain.stabilizer_tableau, key), getattr(folded.stabilizer_tableau, key))
    assert folded.simulate(shots=3, seed=2)[0][0][1] == plain.simulate(shots=3, seed=2)[0][0][1]


def test_apply_gate_and_initial_tableau():
    t0 = ExtendedTableau(2, 3)
    t0.phase_vector[0] = 2                                   # |1, 0>
    c = Circuit(2, 3); c.add_gate("M", [0, 1])
    assert Program(c, tableau=t0).simulate(seed=1) == [MeasurementResult(0, True, 1), MeasurementResult(1, True, 0)]
    assert len(Program(c, tableau=t0).simulate(shots=3, seed=1)[0][0]) == 3
    p = Program(Circuit(2, 3))
    gd = p.circuits[0].gate_data
    assert p.apply_gate(CircuitInstruction(gd, "X", 0)) is None
    assert p.apply_gate(CircuitInstruction(gd, "H", 1)) is None
    # RESET through apply_gate only measures (tableau_gates.py:331-346): |1> stays |1>, no correction applied
    r = p.apply_gate(CircuitInstruction(gd, "RESET", 0))
    assert r == MeasurementResult(0, True, 1)
    assert p.apply_gate(CircuitInstruction(gd, "M", 0)) == MeasurementResult(0, True, 1)
    want = None
    for name, q in (("X", 0), ("H", 1)):
        _, want = run_shot(2, 3, [[gd.get_gate_id(name), q, -1, -1]], tableau=want)
    want.measure(0, lambda: 0)
    same_tableau(p.stabilizer_tableau, want)
    with pytest.raises(ValueError, match="Invalid gate value"):
        bad = CircuitInstruction(gd, "X", 0); bad.gate_id = 99
        p.apply_gate(bad)
    with pytest.raises(ValueError):
        Program(Circuit(3, 3), tableau=t0).simulate()        # tableau / circuit size mismatch


@pytest.mark.parametrize("d", [3, 1031])
def test_replayed_draws_and_record_table_columns(d):
    circ = random_circuit(5, 4, d, 50)
    prog = compile_circuits([circ])
    rm = np.random.default_rng(0).integers(0, d, size=(3, prog.n_meas)).astype(np.uint16 if d > 127 else np.uint8)
    rn = np.zeros((3, prog.n_noise, 2), dtype=rm.dtype)
    table = Program(circ).simulate_records(3, seed=1, replay_meas=rm, replay_noise=rn)
    for s in range(3):
        recs, _ = run_shot(4, d, prog.ops, lambda k: int(rm[s, k]), rn[s])
        assert [(m, det) for _, det, m in recs] == list(zip(table.values[s].tolist(), table.deterministic[s].tolist()))
    assert table.column(int(prog.meas_qudit[-1]), int(prog.meas_round[-1])) == prog.n_meas - 1
    with pytest.raises(ValueError):
        Program(circ).simulate_records(2, method="nope")


def test_frame_method_takes_one_reference_shot_and_frames_for_the_rest():
    """method="frame": shot 0 is a noiseless tableau shot, shots 1.. are Pauli frames (sdim/program.py:244-265);
    force_tableau / shots == 1 stay on the tableau path."""
    c = Circuit(3, 3)
    c.add_gate("H", 0); c.add_gate("CNOT", 0, [1, 2]); c.add_gate("N1", 1, prob=1.0, noise_channel="f")
    c.add_gate("M", [0, 1, 2])
    P = Program(c)
    res = P.simulate(shots=50, seed=3, method="frame")
    assert len(res) == 3 and len(res[0][0]) == 50
    ref = [res[q][0][0].measurement_value for q in range(3)]
    assert ref[0] == ref[1] == ref[2]                        # the reference shot ignores N1 (program.py:31)
    flips = sum(res[0][0][s].measurement_value != res[1][0][s].measurement_value for s in range(1, 50))
    assert flips == 49                                       # every frame carries the certain flip on qudit 1
    assert all(res[0][0][s].measurement_value == res[2][0][s].measurement_value for s in range(50))
    assert {res[0][0][s].measurement_value for s in range(50)} == {0, 1, 2}
    runs_before = len(P._engine.runs)
    P.simulate(shots=50, seed=3, method="frame", force_tableau=True)
    assert P._engine.runs[runs_before][0] == 50              # all 50 shots went through the tableau path
