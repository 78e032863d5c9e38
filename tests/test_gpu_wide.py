"""GPU parity of the uint16-lane path (csrc/wide.cuh): primes 127 < d < 2^15, through the C ABI and the public API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import circuit_from_ops

KEYS = ("x", "z", "p", "dx", "dz", "dp")


def _engine(n, d, ops):
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    prog = compile_circuits([circuit_from_ops(n, d, ops)])
    return prog, TableauEngine(prog)


def test_wide_goldens_replay(golden_wide_primes):
    """Outputs of the UNMODIFIED reference at d in {131, 251, 257, 1031, 32749}: records and all six final arrays,
    bit-exact, under replayed draws (tests/golden/wide_primes.json)."""
    import torch
    for case in golden_wide_primes:
        n, d, ops = case["n"], case["d"], case["ops"]
        prog, eng = _engine(n, d, ops)
        assert eng.plan(None)[0] == "lanes16-global" and eng.layout.elem_bytes == 2 and eng.layout.rec_bytes == 2
        want = np.array([(m & 0x7FFF) | (0x8000 if det else 0) for _, det, m in case["records"]], dtype=np.uint16)
        shots = 3
        rm = torch.from_numpy(np.tile((want & 0x7FFF)[None, :], (shots, 1)).view(np.int16))
        noise = np.array(case["noise_ab"], dtype=np.uint16).reshape(-1, 2)
        rn = torch.from_numpy(np.tile(noise[None], (shots, 1, 1)).view(np.int16)) if prog.n_noise else None
        rec = eng.run(shots, 0, 99, rm, rn, keep_tableau=True)
        got = rec.cpu().numpy().view(np.uint16)
        for s in range(shots):
            assert np.array_equal(got[s], want), f"records differ: seed {case['seed']} n={n} d={d} shot {s}"
        for s in (0, shots - 1):
            arrs = eng.export(eng.tableau, s)
            for key in KEYS:
                assert np.array_equal(arrs[key], np.array(case["final"][key])), (case["seed"], key)


@pytest.mark.parametrize("d,n,depth", [(131, 7, 300), (251, 33, 900), (257, 64, 1500), (1031, 100, 2000), (32749, 19, 600),
                                        (131, 300, 3000)])
def test_wide_philox_mode_matches_c_oracle(d, n, depth):
    """Free-running mode, every opcode incl. M_X, RESET, SWAP and all three noise channels, ragged n, rows wider than
    one pass of the CTA (n = 300: 640 lanes over 256 threads): records and the last shot's final tableau."""
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine
    prog = random_program(seed=1000 * d + n, n=n, d=d, depth=depth)
    shots, seed = 40, 2026 + d
    eng = TableauEngine(prog)
    got = eng.run(shots, 5, seed, keep_tableau=True).cpu().numpy().view(np.uint16)
    want, fin = c_oracle.run(n, d, prog.ops, shots, 5, seed, thresh24=prog.noise_thresh24, channel=prog.noise_channel,
                             want_final=True)
    assert want.dtype == np.uint16 and np.array_equal(got, want)
    arrs = eng.export(eng.tableau, shots - 1)
    for key in KEYS:
        assert np.array_equal(arrs[key], fin[key]), key
    assert (got & 0x7FFF).max() > 127
    if prog.n_noise:
        assert len({got[s].tobytes() for s in range(shots)}) > 1


def test_wide_public_api_and_host_entry():
    """Program.simulate / simulate_records / the host-buffer C entry at d = 257: values above 255 reach
    MeasurementResult, sharding by shot_offset is invariant, stepped modes and apply_gate work on the uint16 store."""
    import random
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200 import Circuit, Program
    from sdim_b200.engine import simulate_host
    d, n = 257, 12
    c = Circuit(n, d)
    for q in range(n):
        c.add_gate("H", q)
    for q in range(n - 1):
        c.add_gate("CNOT", q, q + 1)
    c.add_gate("N1", 3, prob=0.5, noise_channel="d")
    c.add_gate("P", 2)
    c.add_gate("RESET", 0)
    for q in range(n):
        c.add_gate("M", q)
    prog = Program(c)
    table = prog.simulate_records(64, seed=11)
    compiled = prog._compiled()
    want = c_oracle.run_philox(compiled, 64, 0, 11)
    assert table.values.dtype == np.uint16
    assert np.array_equal(table.values, want & 0x7FFF) and np.array_equal(table.deterministic, (want & 0x8000) != 0)
    assert table.values.max() > 255
    # shot ranges
    part = prog.simulate_records(20, seed=11, shot_offset=30)
    assert np.array_equal(part.values, table.values[30:50])
    # host-buffer entry
    rec, _ms = simulate_host(compiled, 64, 0, 11)
    assert rec.dtype == np.uint16 and np.array_equal(rec, want)
    # replay of the oracle's own outcomes
    rm = want & 0x7FFF
    again = prog.simulate_records(64, seed=12345, replay_meas=rm, replay_noise=None)
    assert np.array_equal(again.values[:, -n:][again.deterministic[:, -n:] == False],
                          rm[:, -n:][again.deterministic[:, -n:] == False])
    # public results + final tableau of the last shot against the oracle
    random.seed(5)
    res = Program(c).simulate(shots=1, seed=11)
    assert [r.measurement_value for r in res] == [int(v) for v in (want[0] & 0x7FFF)]
    _, fin = c_oracle.run(n, d, compiled.ops, 1, 0, 11, thresh24=compiled.noise_thresh24, channel=compiled.noise_channel,
                          want_final=True)
    p1 = Program(c)
    p1.simulate(shots=1, seed=11)
    t = p1.stabilizer_tableau
    assert np.array_equal(t.x_block, fin["x"]) and np.array_equal(t.destab_z_block, fin["dz"])
    assert np.array_equal(t.phase_vector, fin["p"]) and np.array_equal(t.destab_phase_vector, fin["dp"])
    # record_tableau (host-stepped) and the frame method's refusal
    snap = Program(c).simulate(shots=1, seed=11, record_tableau=True)
    assert [r.measurement_value for r in snap] == [int(v) for v in (want[0] & 0x7FFF)]
    with pytest.raises(ValueError):
        Program(c).simulate_records(4, seed=1, method="frame")


def test_wide_rejects_what_it_cannot_hold():
    from sdim_b200 import _native as N
    import ctypes as C
    L = N.SdimbLayout()
    assert N.lib().sdimb_layout(4, 32771, C.byref(L)) == N.EDIM          # prime, but above 2^15
    assert N.lib().sdimb_layout(4, 32767, C.byref(L)) == N.EDIM          # composite
    assert N.lib().sdimb_layout(20000, 131, C.byref(L)) == N.ETOOBIG
    assert N.lib().sdimb_layout(4, 32749, C.byref(L)) == N.OK and L.elem_bytes == 2 and L.order == 32749
    k, need = N.plan(8, 131, 0)
    assert (k, need) == (4, True)
    with pytest.raises(ValueError):
        N.plan(8, 131, N.FORCE_PLANES)
