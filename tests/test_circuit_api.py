"""CPU: host-side mirror of the reference interface (Circuit, GateData, .chp IO, IR, result helpers).

Cases follow the reference's own tests (tests/test_program.py:50-148, tests/test_noise_and_io.py:64-98) and
the behaviours listed in SURVEY section 8 (a14, a16-a18)."""
import os
import random

import numpy as np
import pytest

from sdim_b200 import (Circuit, CircuitInstruction, GateData, MeasurementResult, Program, read_circuit,
                       write_circuit, generate_random_clifford_circuit)
from sdim_b200.ir import compile_circuits, is_prime
from sdim_b200.results import MEASUREMENT_DTYPE


def test_gate_ids_and_aliases():
    gd = GateData(3)
    names = ["I", "X", "X_INV", "Z", "Z_INV", "H", "H_INV", "P", "P_INV", "CNOT", "CNOT_INV", "CZ", "CZ_INV",
             "SWAP", "M", "M_X", "RESET", "N1"]
    assert [gd.get_gate_id(n) for n in names] == list(range(18))      # sdim/gatedata.py:65-102
    assert gd.num_gates == 18
    alias = {"R": 5, "DFT": 5, "R_INV": 6, "DFT_INV": 6, "H_DAG": 6, "R_DAG": 6, "DFT_DAG": 6, "PHASE": 7, "S": 7,
             "PHASE_INV": 8, "S_INV": 8, "SUM": 9, "CX": 9, "C": 9, "SUM_INV": 10, "CX_INV": 10, "C_INV": 10,
             "MEASURE": 14, "COLLAPSE": 14, "MZ": 14, "MEASURE_X": 15, "MX": 15, "MR": 16, "MEASURE_RESET": 16,
             "MEASURE_R": 16, "NOISE1": 17}
    for a, gid in alias.items():
        assert gd.get_gate_id(a) == gid, a
    assert gd.get_gate_id("NOPE") is None
    assert gd.get_gate_name(13) == "SWAP"
    with pytest.raises(ValueError):
        gd.get_gate_name(99)
    assert gd.gateMap["N1"].defaults == {"channel": "d", "prob": 0.01}


def test_circuit_construction_and_errors():
    with pytest.raises(ValueError):
        Circuit(0, 2)
    with pytest.raises(ValueError):
        Circuit(2, 1)
    c = Circuit(4, 3)
    with pytest.raises(ValueError, match="not found"):
        c.add_gate("FOO", 0)
    c.add_gate("h", 0)
    assert c.operations[0].gate_name == "H" and c.operations[0].gate_id == 5 and c.operations[0].name == "H"
    c.add_gate("cx", 0, [1, 2, 3])                    # one control, k targets
    assert [(o.qudit_index, o.target_index) for o in c.operations[1:]] == [(0, 1), (0, 2), (0, 3)]
    assert c.operations[1].gate_name == "CX" and c.operations[1].name == "CNOT"
    c.add_gate("CZ", [0, 1, 2], 3)                    # k controls, one target
    c.add_gate("SWAP", [0, 1], [2, 3])                # zipped
    assert [(o.qudit_index, o.target_index) for o in c.operations[-2:]] == [(0, 2), (1, 3)]
    with pytest.raises(ValueError, match="Invalid combination"):
        c.add_gate("CNOT", [0, 1], [1, 2, 3])
    c.add_gate("MEASURE", [0, 1, 2, 3])
    assert [o.gate_id for o in c.operations[-4:]] == [14] * 4
    assert str(c.operations[0]) == "5 0 None"


def test_params_defaults_and_two_qudit_drop():
    c = Circuit(2, 3)
    c.add_gate("N1", 0, prob=0.25, noise_channel="f")
    assert c.operations[0].params == {"prob": 0.25, "noise_channel": "f", "channel": "d"}
    c.add_gate("N1", 1)
    assert c.operations[1].params == {"channel": "d", "prob": 0.01}
    c.add_gate("H", 0)
    assert c.operations[2].params == {}
    c.add_gate("CNOT", 0, 1, prob=0.5)               # reference drops kwargs of two-qudit gates (circuit.py:125)
    assert c.operations[3].params is None


def test_circuit_operators():
    a = Circuit(2, 3); a.add_gate("H", 0)
    b = Circuit(3, 3); b.add_gate("X", 2)
    s = a + b
    assert s.num_qudits == 3 and [o.gate_id for o in s.operations] == [5, 1]
    a += b
    assert a.num_qudits == 3 and len(a.operations) == 2
    r = a * 3
    assert r is a and len(a.operations) == 6          # `*` mutates and returns self (circuit.py:128-141)
    with pytest.raises(ValueError):
        _ = a + Circuit(2, 5)
    c = Circuit.from_operation_list([("H", [0]), ("CNOT", [0, 1]), CircuitInstruction(a.gate_data, "P", 1)], 2, 3)
    assert [o.gate_id for o in c.operations] == [5, 9, 7]
    with pytest.raises(ValueError):
        Circuit.from_operation_list([("H", [0, 1, 2])], 3, 3)
    with pytest.raises(ValueError):
        Circuit.from_operation_list([42], 3, 3)


def test_read_shipped_circuits():
    epr = read_circuit("circuits/epr.chp")
    assert (epr.num_qudits, epr.dimension) == (2, 2)
    assert [(o.gate_id, o.qudit_index, o.target_index) for o in epr.operations] == [(5, 0, None), (9, 0, 1), (14, 1, None)]
    st = read_circuit("circuits/css_steane_final.chp")
    assert (st.num_qudits, st.dimension, len(st.operations)) == (13, 2, 52)


def test_chp_round_trip_with_params(tmp_path):
    """reference tests/test_noise_and_io.py:64-98 (generic_read_write_test), seeded."""
    from make_cases import random_circuit
    for seed, noisy in ((1, False), (2, True)):
        c = random_circuit(seed, 23, 5, 3000, p_meas=0.0, p_noise=0.15 if noisy else 0.0, final_measure=False)
        path = write_circuit(c, "rt.chp", comment="To go where no test has ever gone.", directory=str(tmp_path))
        cc = read_circuit(path)
        assert cc.dimension == c.dimension and len(cc.operations) == len(c.operations)
        for op, op2 in zip(c.operations, cc.operations):
            assert (op.gate_name, op.qudit_index, op.target_index, op.gate_id, op.name) == \
                   (op2.gate_name, op2.qudit_index, op2.target_index, op2.gate_id, op2.name)
            if op.params is None:
                assert op2.params is None
            else:
                for k in op.params:
                    assert str(op.params[k]) == str(op2.params[k])
    text = open(path).read().splitlines()
    assert text[1] == "#" and text[2] == "d 5"


def test_read_circuit_errors(tmp_path):
    p = tmp_path / "bad.chp"
    p.write_text("hdr\n#\nd 3\nH 0 1 2\n")
    with pytest.raises(ValueError, match="Unexpected number of arguments"):
        read_circuit(str(p))
    p.write_text("hdr\n#\nd 3\nN1 0 prob=0.1=2\n")
    with pytest.raises(ValueError, match="correct format"):
        read_circuit(str(p))
    p.write_text("hdr\n#\nh 0\nc 0 1\nm 1\n")                  # no `d` line -> qubits
    assert read_circuit(str(p)).dimension == 2


def test_random_circuit_matches_reference_stream():
    """Same stdlib-random draw order as sdim/random_circuit.py:39-61; fixture generated from the reference."""
    c = generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1)
    assert len(c.operations) == 2064
    head = [(o.name, o.qudit_index, o.target_index) for o in c.operations[:4]]
    assert head == [("CNOT", 8, 32), ("P", 63, None), ("CNOT_INV", 60, 48), ("X", 12, None)]
    assert all(o.gate_id == 14 for o in c.operations[2000:])
    c2 = generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1)
    assert [str(o) for o in c.operations] == [str(o) for o in c2.operations]


def test_ir_compile_slots_and_validation():
    c = Circuit(3, 3)
    c.add_gate("H", 0); c.add_gate("I", 1); c.add_gate("CNOT", 0, 1)
    c.add_gate("N1", 2, prob=1.0, noise_channel="p")
    c.add_gate("M", [0, 1]); c.add_gate("RESET", 0); c.add_gate("M_X", 2); c.add_gate("N1", 0)
    p = compile_circuits([c])
    assert p.n_user_gates == 9 and p.n_ops == 8                 # I dropped from the stream, still counted
    assert p.ops.tolist() == [[5, 0, -1, -1], [9, 0, 1, -1], [17, 2, -1, 0], [14, 0, -1, 0], [14, 1, -1, 1],
                              [16, 0, -1, 2], [15, 2, -1, 3], [17, 0, -1, 1]]
    assert p.meas_qudit.tolist() == [0, 1, 0, 2] and p.meas_round.tolist() == [0, 0, 1, 0]
    assert p.rounds_per_qudit == [2, 1, 1]
    assert p.noise_channel.tolist() == [2, 0] and p.noise_thresh24.tolist() == [0, round(0.99 * 2 ** 24)]
    bad = Circuit(2, 3); bad.add_gate("CNOT", 0, 0)
    with pytest.raises(ValueError):
        compile_circuits([bad])
    bad = Circuit(2, 3); bad.operations.append(CircuitInstruction(bad.gate_data, "H", 5))
    with pytest.raises(ValueError):
        compile_circuits([bad])
    bad = Circuit(2, 3); bad.add_gate("N1", 0, noise_channel="q")
    with pytest.raises(ValueError):
        compile_circuits([bad])
    assert is_prime(2) and is_prime(127) and not is_prime(1) and not is_prime(9) and not is_prime(4)


def test_program_rejects_composite_dimension():
    with pytest.raises(ValueError, match="not prime"):
        Program(Circuit(2, 4))


def test_results_to_array_and_combine():
    """reference tests/test_program.py:50-110."""
    results = [[MeasurementResult(0, True, 1), MeasurementResult(0, True, 2)],
               [MeasurementResult(2, True, 3), MeasurementResult(2, True, 4)]]
    expected = np.array([[(0, 0, 0, True, 1), (0, 1, 0, True, 2)], [(2, 0, 0, True, 3), (2, 1, 0, True, 4)]],
                        dtype=MEASUREMENT_DTYPE)
    np.testing.assert_array_equal(Program._results_to_array(results), expected)
    nested = [[[MeasurementResult(0, True, 0)]], [[MeasurementResult(1, True, 0)]]]
    expected = np.array([[(0, 0, 0, True, 0)], [(1, 0, 0, True, 0)]], dtype=MEASUREMENT_DTYPE)
    np.testing.assert_array_equal(Program._results_to_array(nested), expected)
    with pytest.raises(ValueError):
        Program._results_to_array([])
    with pytest.raises(ValueError):
        Program._results_to_array([[3]])
    c = Circuit(2, 3); c.add_gate("M", [0, 1])
    prog = Program(c)
    prog.measurement_results = nested
    extra = np.array([[[(0, 0, 0, True, 0)]], [[(1, 0, 0, True, 0)]]], dtype=MEASUREMENT_DTYPE)
    prog._combine_results(extra)
    assert prog.measurement_results == [[[MeasurementResult(0, True, 0), MeasurementResult(0, True, 0)]],
                                        [[MeasurementResult(1, True, 0), MeasurementResult(1, True, 0)]]]


def test_build_ir_reference_format():
    """reference tests/test_program.py:112-148."""
    ir_dtype = np.dtype([("gate_id", np.int64), ("qudit_index", np.int64), ("target_index", np.int64)])
    c = Circuit(dimension=3, num_qudits=2)
    c.add_gate("H", 0); c.add_gate("CNOT", 0, 1); c.add_gate("M", [0, 1])
    p = Program(c)
    ir, noise = p._build_ir(p.circuits, 1)
    np.testing.assert_array_equal(ir, np.array([(5, 0, -1), (9, 0, 1), (14, 0, -1), (14, 1, -1)], dtype=ir_dtype))
    assert noise.shape == (1, 1, 2)
    c.add_gate("N1", 0, prob=1.0, noise_channel="f")
    c.add_gate("N1", 1, prob=1.0, noise_channel="p")
    c.add_gate("N1", 1, prob=1.0, noise_channel="d")
    ir, noise = p._build_ir(p.circuits, 5)
    assert ir[-3:].tolist() == [(17, 0, -1), (17, 1, -1), (17, 1, -1)] and noise.shape == (3, 5, 2)
    assert (noise[0, :, 0] > 0).all() and (noise[0, :, 1] == 0).all()
    assert (noise[1, :, 1] > 0).all() and (noise[1, :, 0] == 0).all()
    assert (noise[2].sum(axis=1) > 0).all()


def test_measurement_result_semantics():
    m = MeasurementResult(3, False, 2)
    assert str(m) == "Measured qudit (3) as (2) and was random" == repr(m)
    assert str(MeasurementResult(0, True, 1)) == "Measured qudit (0) as (1) and was deterministic"
    assert m == MeasurementResult(3, False, 2) and m != MeasurementResult(3, True, 2)
    with pytest.raises(ValueError):
        m.get_tableau()


def test_append_circuit():
    a = Circuit(2, 3); b = Circuit(4, 3)
    p = Program(a)
    p.append_circuit(b)
    assert a.num_qudits == 4 and len(p.circuits) == 2
    with pytest.raises(ValueError):
        p.append_circuit(Circuit(4, 5))


def test_workload_builders():
    from sdim_b200.workloads import noisy_random_clifford, qudit_repetition_code, rotated_surface_code
    sc = rotated_surface_code(7, 7)
    assert sc.num_qudits == 97 and sc.dimension == 2            # 49 data + 48 ancilla
    p = compile_circuits([sc])
    assert p.n_meas == 7 * 48 + 49 and p.n_noise == 7 * 49
    rc = qudit_repetition_code(25, 25, 3)
    p = compile_circuits([rc])
    assert rc.num_qudits == 49 and p.n_meas == 625 and p.n_noise == 625     # config 4: 625 records per shot
    hl = compile_circuits([noisy_random_clifford(256, 2000, 3)])
    assert hl.n_meas == 256 and hl.n_ops == hl.n_user_gates == 2000 + hl.n_noise + 256


def test_record_table_structured_view_and_npz_round_trip(tmp_path):
    """RecordTable -> the reference's structured record array (sdim/program.py:34-40, [qudit, round, shot]) and a
    lossless .npz round trip; no GPU involved (the table is built by hand)."""
    import numpy as np
    from sdim_b200.program import RecordTable
    from sdim_b200.results import MEASUREMENT_DTYPE
    rng = np.random.default_rng(0)
    meas_qudit = np.array([0, 2, 0, 1, 2, 2], dtype=np.int32)          # qudit 2 is measured three times
    meas_round = np.array([0, 0, 1, 0, 1, 2], dtype=np.int32)
    values = rng.integers(0, 5, size=(7, 6)).astype(np.uint8)
    det = rng.random((7, 6)) < 0.4
    table = RecordTable(values, det, meas_qudit, meas_round, seed=2**63 - 5, shot_offset=100)
    arr = table.to_structured(num_qudits=4)
    assert arr.dtype == MEASUREMENT_DTYPE and arr.shape == (4, 3, 7)
    for k in range(6):
        cell = arr[meas_qudit[k], meas_round[k]]
        assert np.array_equal(cell["measurement_value"], values[:, k])
        assert np.array_equal(cell["deterministic"], det[:, k])
        assert np.array_equal(cell["shot"], np.arange(100, 107))
        assert set(cell["qudit_index"]) == {meas_qudit[k]} and set(cell["meas_round"]) == {meas_round[k]}
    zero = np.zeros((), dtype=MEASUREMENT_DTYPE)
    assert (arr[3] == zero).all() and (arr[1, 1:] == zero).all()       # never measured / fewer rounds: zeros
    assert table.column(2, 2) == 5
    path = table.save(str(tmp_path / "records"))
    back = RecordTable.load(path)
    assert np.array_equal(back.values, values) and np.array_equal(back.deterministic, det)
    assert np.array_equal(back.meas_qudit, meas_qudit) and np.array_equal(back.meas_round, meas_round)
    assert back.seed == 2**63 - 5 and back.shot_offset == 100 and back.shots == 7


def test_program_fold_gates_compiles_a_shorter_equivalent_stream():
    from sdim_b200 import Circuit, Program
    c = Circuit(3, 3)
    for _ in range(4):
        c.add_gate("H", 0)
    c.add_gate("CNOT", 0, 1); c.add_gate("CNOT_INV", 0, 1); c.add_gate("P", 2); c.add_gate("M", [0, 1, 2])
    prog = Program(c, fold_gates=True)
    full, slim = prog._compiled(), prog._compiled(fold=True)
    assert full.n_ops == 10 and slim.n_ops == 4 and slim.n_user_gates == full.n_user_gates == 10
    assert slim.ops[:, 0].tolist() == [7, 14, 14, 14]


def test_undo_reset_correction_recovers_the_post_measurement_tableau():
    """record_tableau snapshots of a RESET are taken where the reference takes them: after measure(), before the
    X^(-m) correction (sdim/program.py:323-324 vs :335-339).  Checked on the numpy oracle's tableau."""
    import numpy as np
    from oracle.tableau_oracle import run_shot
    from make_cases import random_program
    from sdim_b200.program import undo_reset_correction
    keys = ("x", "z", "p", "dx", "dz", "dp")
    for d in (2, 3, 5, 7):
        prog = random_program(seed=90 + d, n=5, d=d, depth=70, p_meas=0.0)
        for q in range(5):
            for m in range(d):
                _, t = run_shot(5, d, prog.ops.tolist(), lambda k: 0)
                det, value = t.measure(q, lambda: m)
                before = {k: a.copy() for k, a in zip(keys, t.arrays())}
                t.pauli(q, (-value) % d, 0)
                after = {k: a.copy() for k, a in zip(keys, t.arrays())}
                got = undo_reset_correction(after, q, value, d)
                for k in keys:
                    assert np.array_equal(np.asarray(got[k]) % (2 * d if d == 2 else d) if k in ("p", "dp") else got[k],
                                          before[k]), (d, q, m, k)
