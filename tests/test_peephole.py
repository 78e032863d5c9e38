"""Peephole folding of the op stream (sdim_b200/peephole.py): exact against the oracles, CPU only."""
import numpy as np
import pytest

from make_cases import random_program
from oracle import c_oracle
from oracle.tableau_oracle import run_shots
from sdim_b200.circuit import Circuit
from sdim_b200.ir import compile_circuits
from sdim_b200.peephole import family_order, fold_ops, fold_program


def _final_and_records(prog, ops, shots, seed):
    rec, fin = c_oracle.run(prog.num_qudits, prog.dimension, ops, shots, 0, seed, thresh24=prog.noise_thresh24,
                            channel=prog.noise_channel, want_final=True)
    return rec, fin


@pytest.mark.parametrize("d", [2, 3, 5, 7])
def test_family_orders_are_exact_on_the_tableau(d):
    """G^order leaves a generic tableau (phases included) unchanged, for every single-qudit family."""
    base = random_program(seed=7 + d, n=6, d=d, depth=80, p_meas=0.0)
    for name, inv in (("X", "X_INV"), ("Z", "Z_INV"), ("H", "H_INV"), ("P", "P_INV")):
        c = Circuit(6, d)
        order = family_order(c.gate_data.get_gate_id(name), d)
        for _ in range(order):
            c.add_gate(name, 2)
        tail = compile_circuits([c]).ops
        _, want = _final_and_records(base, base.ops, 1, 3)
        _, got = _final_and_records(base, np.concatenate([base.ops, tail]), 1, 3)
        for key in want:
            assert np.array_equal(want[key], got[key]), (name, key)
        assert fold_ops(tail, d).shape[0] == 0
        c.add_gate(inv, 2)                                   # G^order G^-1 -> one G^-1
        folded = fold_ops(compile_circuits([c]).ops, d)
        assert folded.shape[0] == 1


def _redundant_circuit(n, d, depth, seed):
    """Random circuit salted with cancelling and foldable patterns, measurements, RESET and noise."""
    rng = np.random.default_rng(seed)
    c = Circuit(n, d)
    single = ["X", "X_INV", "Z", "Z_INV", "H", "H_INV", "P", "P_INV"]
    pairs = [("CNOT", "CNOT_INV"), ("CZ", "CZ_INV"), ("SWAP", "SWAP")]
    for _ in range(depth):
        r = rng.random()
        q = int(rng.integers(n))
        if r < 0.45:
            g = single[int(rng.integers(len(single)))]
            for _ in range(int(rng.integers(1, 6))):         # runs of one family, mixed directions
                c.add_gate(g if rng.random() < 0.7 else single[single.index(g) ^ 1], q)
        elif r < 0.75 and n > 1:
            t = int((q + 1 + rng.integers(n - 1)) % n)
            a, b = pairs[int(rng.integers(len(pairs)))]
            c.add_gate(a, q, t)
            if rng.random() < 0.5:
                c.add_gate(single[int(rng.integers(len(single)))], int((t + 1) % n) if n > 2 else q)
            if rng.random() < 0.6:
                if a != "CNOT" and rng.random() < 0.5:
                    c.add_gate(b, t, q)                      # CZ / SWAP cancel in either order
                else:
                    c.add_gate(b, q, t)
        elif r < 0.83:
            c.add_gate("N1", q, prob=0.3, noise_channel="dfp"[int(rng.integers(3))])
        elif r < 0.93:
            c.add_gate(["M", "M_X", "RESET"][int(rng.integers(3))], q)
        else:
            c.add_gate("I", q)
    c.add_gate("M", list(range(n)))
    return c


@pytest.mark.parametrize("d,n", [(2, 5), (2, 12), (3, 4), (3, 9), (5, 6), (7, 3)])
def test_folded_stream_gives_identical_records_and_final_tableau(d, n):
    folded_any = 0
    for seed in range(6):
        prog = compile_circuits([_redundant_circuit(n, d, 220, 100 * d + seed)])
        slim = fold_program(prog)
        assert slim.n_ops < prog.n_ops and slim.n_user_gates == prog.n_user_gates
        assert slim.n_meas == prog.n_meas and slim.n_noise == prog.n_noise
        # event slots are untouched and still in chronological order
        for code in ((14, 15, 16), (17,)):
            keep = np.isin(slim.ops[:, 0], code)
            assert np.array_equal(slim.ops[keep][:, 3], prog.ops[np.isin(prog.ops[:, 0], code)][:, 3])
        rec_a, fin_a = _final_and_records(prog, prog.ops, 24, 11 + seed)
        rec_b, fin_b = _final_and_records(prog, slim.ops, 24, 11 + seed)
        assert np.array_equal(rec_a, rec_b)
        for key in fin_a:
            assert np.array_equal(fin_a[key], fin_b[key]), key
        folded_any += prog.n_ops - slim.n_ops
    assert folded_any > 100


def test_folding_matches_the_numpy_oracle_too():
    """Same check through the numpy restatement (reference orientation, int64) under replayed draws."""
    from sdim_b200.rng import measurement_draws, noise_draws
    prog = compile_circuits([_redundant_circuit(5, 3, 150, 5)])
    slim = fold_program(prog)
    ids = np.arange(8)
    md = measurement_draws(3, 3, ids, prog.n_meas)
    nd = noise_draws(3, 3, ids, prog.noise_thresh24, prog.noise_channel)
    a, ta = run_shots(5, 3, prog.ops, 8, meas_draws=md, noise_ab=nd)
    b, tb = run_shots(5, 3, slim.ops, 8, meas_draws=md, noise_ab=nd)
    assert np.array_equal(a, b)


def test_nothing_folds_across_a_measurement_or_an_intervening_gate():
    c = Circuit(3, 3)
    c.add_gate("H", 0); c.add_gate("M", 1); c.add_gate("H_INV", 0)          # measurement: full barrier
    c.add_gate("P", 2); c.add_gate("CNOT", 2, 1); c.add_gate("P_INV", 2)     # CNOT touches qudit 2
    c.add_gate("CNOT", 0, 1); c.add_gate("CNOT_INV", 1, 0)                   # CNOT is directional
    c.add_gate("X", 1); c.add_gate("N1", 1, prob=0.1, noise_channel="d"); c.add_gate("X_INV", 1)
    prog = compile_circuits([c])
    assert fold_ops(prog.ops, 3).shape[0] == prog.n_ops
    c2 = Circuit(3, 3)
    c2.add_gate("H", 0); c2.add_gate("X", 1); c2.add_gate("H_INV", 0)        # X on another qudit does not block
    c2.add_gate("CZ", 0, 2); c2.add_gate("CZ_INV", 2, 0)
    out = fold_ops(compile_circuits([c2]).ops, 3)
    assert out.tolist() == [[1, 1, -1, -1]]
