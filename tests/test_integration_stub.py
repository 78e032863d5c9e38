"""The reference-side binding of INTEGRATION.md, executed for real.

The ctypes stub is cut out of INTEGRATION.md verbatim and imported next to the UNMODIFIED reference package
(/root/reference in the build container, baseline/_ref — installed by baseline/install_reference.sh — on the GPU box).
Circuits and Programs are built with the REFERENCE's own API (sdim.Circuit.add_gate, sdim.Program,
sdim/circuit.py:76-126, sdim/program.py:194-204); the stub lowers them to the C ABI's op stream.

CPU: the op rows / noise tables / record slots equal what sdim_b200.ir.compile_circuits produces for the mirrored
sdim_b200 circuit (so both front ends feed the library identically).
GPU: the stub's sdimb_simulate_host call returns, regrouped as measurement_results[qudit][round][shot], exactly the C
oracle's records for the same Philox seed."""
import importlib.util
import os
import random
import re
import sys

import numpy as np
import pytest

from oracle import ref_harness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not rh.reference_available(),
                                     reason="no copy of the reference (run baseline/install_reference.sh)")


def _load_stub(tmp_path):
    sdim = rh.load_reference()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# sdim/_b200\.py.*?)```", text, re.S).group(1)
    path = os.path.join(str(tmp_path), "_b200_stub.py")
    with open(path, "w") as fh:
        fh.write(block)
    spec = importlib.util.spec_from_file_location("_b200_stub", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return sdim, mod


def _random_pair(sdim, seed, n, d, depth):
    """The same random circuit built twice: with the reference's Circuit API and with sdim_b200's."""
    from sdim_b200 import Circuit
    rng = random.Random(seed)
    ref, ours = sdim.Circuit(n, d), Circuit(n, d)
    one = ["I", "X", "X_INV", "Z", "Z_INV", "H", "H_INV", "P", "P_INV", "M", "M_X", "RESET"]
    two = ["CNOT", "CNOT_INV", "CZ", "CZ_INV", "SWAP"]
    for _ in range(depth):
        u = rng.random()
        if u < 0.12:
            q, pr, chn = rng.randrange(n), rng.choice([0.0, 0.05, 0.5, 1.0]), rng.choice("dfp")
            for c in (ref, ours):
                c.add_gate("N1", q, prob=pr, noise_channel=chn)
        elif u < 0.45 and n >= 2:
            a, b = rng.sample(range(n), 2)
            name = rng.choice(two)
            for c in (ref, ours):
                c.add_gate(name, a, b)
        else:
            name, q = rng.choice(one), rng.randrange(n)
            for c in (ref, ours):
                c.add_gate(name, q)
    for c in (ref, ours):
        c.add_gate("M", list(range(n)))
    return ref, ours


@needs_reference
@pytest.mark.parametrize("d,n,depth", [(2, 5, 80), (3, 9, 200), (5, 4, 120), (7, 13, 300)])
def test_stub_lowers_a_reference_program_like_our_compiler(tmp_path, d, n, depth):
    from sdim_b200.ir import compile_circuits
    sdim, stub = _load_stub(tmp_path)
    ref_circ, our_circ = _random_pair(sdim, 40 + d, n, d, depth)
    program = sdim.Program(ref_circ)                       # the reference's own Program (ExtendedTableau inside)
    assert type(program).__module__ == "sdim.program" and program.stabilizer_tableau.num_qudits == n
    ops, thr, ch, slots = stub.build_op_stream(program)
    want = compile_circuits([our_circ])
    assert np.array_equal(ops, want.ops)
    assert np.array_equal(thr, want.noise_thresh24) and np.array_equal(ch, want.noise_channel)
    assert slots == want.meas_qudit.tolist()
    # appended circuits run back to back (program.py:311-312)
    tail_ref, tail_ours = _random_pair(sdim, 90 + d, n, d, 20)
    program.append_circuit(tail_ref)
    ops2, _, _, slots2 = stub.build_op_stream(program)
    want2 = compile_circuits([our_circ, tail_ours])
    assert np.array_equal(ops2, want2.ops) and slots2 == want2.meas_qudit.tolist()


@needs_reference
def test_stub_on_the_shipped_chp_files(tmp_path):
    """read_circuit of the reference -> stub == read_circuit of sdim_b200 -> compile_circuits."""
    from sdim_b200 import read_circuit
    from sdim_b200.ir import compile_circuits
    sdim, stub = _load_stub(tmp_path)
    for name in ("epr.chp", "css_steane_final.chp"):
        ref_circ = sdim.read_circuit(os.path.join(ROOT, "circuits", name))
        ops, _, _, slots = stub.build_op_stream(sdim.Program(ref_circ))
        want = compile_circuits([read_circuit(os.path.join(ROOT, "circuits", name))])
        assert np.array_equal(ops, want.ops) and slots == want.meas_qudit.tolist(), name


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("d,n,depth,shots", [(2, 13, 300, 64), (3, 40, 900, 50), (5, 9, 200, 33)])
def test_stub_end_to_end_on_the_gpu(tmp_path, monkeypatch, d, n, depth, shots):
    """A reference Program through the INTEGRATION.md stub and libsdimb.so: measurement_results[qudit][round][shot]
    of reference MeasurementResult objects, equal to the C oracle's records for the same seed."""
    from oracle import c_oracle
    from sdim_b200.build import LIB_PATH
    from sdim_b200.ir import compile_circuits
    monkeypatch.setenv("SDIMB_LIB", LIB_PATH)
    sdim, stub = _load_stub(tmp_path)
    ref_circ, our_circ = _random_pair(sdim, 7 + d, n, d, depth)
    program = sdim.Program(ref_circ)
    out = stub.simulate_tableau_b200(program, shots, 2026)
    prog = compile_circuits([our_circ])
    want = c_oracle.run_philox(prog, shots, 0, 2026)
    assert program.measurement_results is out and len(out) == n
    seen = [0] * n
    for k, q in enumerate(prog.meas_qudit.tolist()):
        column = out[q][seen[q]]
        seen[q] += 1
        assert len(column) == shots and type(column[0]).__module__ == "sdim.tableau.dataclasses"
        assert [int(r.measurement_value) | (0x80 if r.deterministic else 0) for r in column] == want[:, k].tolist()
        assert all(r.qudit_index == q for r in column)
    with pytest.raises(ValueError):                         # error behaviour: bad dimension -> ValueError
        bad = sdim.Program(sdim.Circuit(2, 3))
        bad.stabilizer_tableau.dimension = 4
        stub.simulate_tableau_b200(bad, 1, 0)
