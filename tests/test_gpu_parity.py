"""GPU parity: the CUDA interpreter (through the C ABI) against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import circuit_from_ops, pack_records


def _engine(n, d, ops):
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    prog = compile_circuits([circuit_from_ops(n, d, ops)])
    return prog, TableauEngine(prog)


@pytest.mark.parametrize("mode", ["resident", "global", "planes", "planes-warp", "cluster", "planes-global"])
def test_golden_random_circuits_replay(golden_random, mode):
    """Every reference golden case: records AND all six final arrays, bit-exact, under replayed draws."""
    import torch
    checked = 0
    for case in golden_random:
        n, d, ops = case["n"], case["d"], case["ops"]
        if mode.startswith("planes") and d > 3:
            continue
        prog, eng = _engine(n, d, ops)
        assert prog.n_ops == sum(1 for o in ops if o[0] != 0)
        want = np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in case["records"]], dtype=np.uint8)
        shots = 3   # identical replay in every shot -> identical records
        rm = torch.from_numpy(np.tile((want & 0x7F)[None, :], (shots, 1)))
        noise = np.array(case["noise_ab"], dtype=np.uint8).reshape(-1, 2)
        rn = torch.from_numpy(np.tile(noise[None], (shots, 1, 1))) if prog.n_noise else None
        rec = eng.run(shots, 0, 1234, rm, rn, keep_tableau=True, mode=mode)
        got = rec.cpu().numpy()
        for s in range(shots):
            assert np.array_equal(got[s], want), f"records differ: seed {case['seed']} n={n} d={d} shot {s}"
        for s in (0, shots - 1):
            arrs = eng.export(eng.tableau, s)
            for key in ("x", "z", "p", "dx", "dz", "dp"):
                assert np.array_equal(arrs[key], np.array(case["final"][key])), \
                    f"final {key} differs: seed {case['seed']} n={n} d={d}"
        checked += 1
    assert checked == (len(golden_random) if not mode.startswith("planes") else sum(c["d"] <= 3 for c in golden_random))
    assert checked >= 40


@pytest.mark.parametrize("mode", [None, "resident", "global", "planes", "planes-warp", "planes-global", "cluster"])
def test_golden_config_sizes_replay(golden_config_sizes, mode):
    """Outputs of the UNMODIFIED reference at the BASELINE.json config sizes (config 2 n = 64, surface code n = 97,
    repetition code n = 49, headline n = 256): records and all six final arrays in every kernel that can hold the
    shape (tests/golden/config_sizes.npz, oracle/make_golden.py --configs)."""
    import torch
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    checked = 0
    for case in golden_config_sizes:
        n, d, ops = case["n"], case["d"], case["ops"]
        prog = compile_circuits([circuit_from_ops(n, d, ops)])
        eng = TableauEngine(prog)
        try:
            eng.plan(mode)
        except ValueError:
            continue                      # this shape does not fit that kernel (n = 256 in shared memory)
        want = np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in case["records"]], dtype=np.uint8)
        shots = 5
        rm = torch.from_numpy(np.tile((want & 0x7F)[None, :], (shots, 1)))
        rn = torch.from_numpy(np.tile(case["noise_ab"][None], (shots, 1, 1))) if prog.n_noise else None
        got = eng.run(shots, 0, 99, rm, rn, keep_tableau=True, mode=mode).cpu().numpy()
        for s in range(shots):
            assert np.array_equal(got[s], want), f"records differ: {case['name']} shot {s} mode {mode}"
        arrs = eng.export(eng.tableau, shots - 1)
        for key in ("x", "z", "p", "dx", "dz", "dp"):
            assert np.array_equal(arrs[key], case["final"][key]), f"final {key} differs: {case['name']} mode {mode}"
        checked += 1
    assert checked >= (6 if mode in (None, "global", "planes-global", "cluster") else 5)


def test_headline_free_running_shot0_equals_reference_golden(golden_config_sizes):
    """The headline golden's N1 events are shot 0 of Philox seed 2026.  With only the measurement outcomes replayed
    and the noise left to the device's own Philox draws, global shot 0 of the bench workload must reproduce the
    reference's records: pins the device-side noise draws to the reference at the headline size."""
    import torch
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import noisy_random_clifford
    case = next(c for c in golden_config_sizes if c["name"].startswith("headline"))
    prog = compile_circuits([noisy_random_clifford(256, 2000, 3, seed=1, prob=1e-3, channel="d")])
    assert np.array_equal(prog.ops, case["ops"])
    want = np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in case["records"]], dtype=np.uint8)
    rm = torch.from_numpy((want & 0x7F)[None, :].copy())
    got = TableauEngine(prog).run(1, 0, 2026, rm, None).cpu().numpy()      # noise: Philox seed 2026, global shot 0
    assert np.array_equal(got[0], want)


# ---------------------------------------------------------------------------------------------------
# Free-running (Philox) mode against the C oracle: same counters on both sides -> bit-exact records
# ---------------------------------------------------------------------------------------------------
def _run_gpu(prog, shots, seed, mode=None, shot_offset=0, keep=False):
    from sdim_b200.engine import TableauEngine
    eng = TableauEngine(prog)
    rec = eng.run(shots, shot_offset, seed, mode=mode, keep_tableau=keep)
    return eng, rec.cpu().numpy()


@pytest.mark.parametrize("d,n,depth", [(2, 5, 120), (2, 40, 900), (2, 97, 2500), (3, 1, 30), (3, 17, 500), (3, 64, 1500), (3, 100, 2500),
                                        (5, 33, 800), (7, 16, 400), (11, 9, 300), (13, 21, 500), (127, 6, 200)])
@pytest.mark.parametrize("mode", ["resident", "global", "planes", "planes-warp", "planes-global"])
def test_philox_mode_matches_c_oracle(d, n, depth, mode):
    """Every opcode incl. M_X, RESET, SWAP and all three noise channels; ragged n (not a multiple of 16)."""
    from make_cases import random_program
    from oracle import c_oracle
    if mode.startswith("planes") and d > 3:
        pytest.skip("bit planes exist for d = 2, 3")
    prog = random_program(seed=1000 * d + n, n=n, d=d, depth=depth)
    shots, seed = 96, 2026 + d
    eng, got = _run_gpu(prog, shots, seed, mode, keep=True)
    want, fin = c_oracle.run(n, d, prog.ops, shots, 0, seed, thresh24=prog.noise_thresh24,
                             channel=prog.noise_channel, want_final=True)
    assert np.array_equal(got, want)
    arrs = eng.export(eng.tableau, shots - 1)       # final tableau of the last shot, all six arrays
    for key in ("x", "z", "p", "dx", "dz", "dp"):
        assert np.array_equal(arrs[key], fin[key]), key
    if prog.n_noise:
        assert len({got[s].tobytes() for s in range(shots)}) > 1    # noise/draws really differ between shots


@pytest.mark.parametrize("img", ["scratch", "shared"])
@pytest.mark.parametrize("uni", [True, False])
def test_tile_interpreter_both_measurement_forms(golden_random, golden_config_sizes, monkeypatch, uni, img):
    """The tile interpreter (several shots per warp, n <= 128) with its images in caller scratch (the default, 20 warps
    per SM) and in shared memory (SDIMB_TILE_SMEM), and with its measurement walks shared by the warp (fresh
    batches: identical X / Z blocks in every tile) and per tile (SDIMB_TILE_NO_UNI, the form non-fresh runs take):
    reference goldens incl. final tableaus, the config-size goldens that fit, and free-running shots that differ in
    noise and outcomes against the C oracle — with shot counts that leave tiles of the last warp empty."""
    import torch
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    if not uni:
        monkeypatch.setenv("SDIMB_TILE_NO_UNI", "1")
    if img == "shared":
        monkeypatch.setenv("SDIMB_TILE_SMEM", "1")
    else:
        monkeypatch.setenv("SDIMB_TILE_GLB", "1")          # also for the few shots of this test
    checked = 0
    cases = [c for c in golden_random if c["d"] <= 3] + [c for c in golden_config_sizes if c["n"] <= 128]
    for case in cases:
        n, d = case["n"], case["d"]
        ops = case["ops"].tolist() if hasattr(case["ops"], "tolist") else case["ops"]
        prog = compile_circuits([circuit_from_ops(n, d, ops)])
        eng = TableauEngine(prog)
        assert eng.plan("planes")[0] == "planes-tile"
        want = np.array([(int(m) & 0x7F) | (0x80 if det else 0) for _, det, m in case["records"]], dtype=np.uint8)
        shots = 11
        rm = torch.from_numpy(np.tile((want & 0x7F)[None, :], (shots, 1)))
        noise = np.array(case["noise_ab"], dtype=np.uint8).reshape(-1, 2)
        rn = torch.from_numpy(np.tile(noise[None], (shots, 1, 1))) if prog.n_noise else None
        got = eng.run(shots, 0, 77, rm, rn, keep_tableau=True, mode="planes").cpu().numpy()
        assert all(np.array_equal(got[s], want) for s in range(shots)), (n, d)
        arrs = eng.export(eng.tableau, shots - 1)
        for key in ("x", "z", "p", "dx", "dz", "dp"):
            assert np.array_equal(arrs[key], np.array(case["final"][key])), (n, d, key)
        checked += 1
    assert checked >= 45
    for d, n, depth, shots in ((2, 97, 2500, 37), (3, 49, 2000, 1001), (3, 64, 1500, 13), (2, 128, 1200, 9), (3, 5, 200, 3), (2, 33, 700, 50)):
        prog = random_program(seed=31 * d + n, n=n, d=d, depth=depth)
        eng = TableauEngine(prog)
        got = eng.run(shots, 3, 99, mode="planes", keep_tableau=True).cpu().numpy()
        want, fin = c_oracle.run(n, d, prog.ops, shots, 3, 99, thresh24=prog.noise_thresh24, channel=prog.noise_channel,
                                 want_final=True)
        assert np.array_equal(got, want), (d, n)
        arrs = eng.export(eng.tableau, shots - 1)
        for key in ("x", "z", "p", "dx", "dz", "dp"):
            assert np.array_equal(arrs[key], fin[key]), (d, n, key)


def test_noise_cache_refill_many_events():
    """More N1 events than one shared-memory noise chunk (1024), interleaved with measurements."""
    from make_cases import random_program
    from oracle import c_oracle
    prog = random_program(seed=77, n=8, d=3, depth=9000, p_meas=0.05, p_noise=0.4, noise_prob=0.2)
    assert prog.n_noise > 3000
    _, got = _run_gpu(prog, 40, 5)
    want = c_oracle.run_philox(prog, 40, 0, 5)
    assert np.array_equal(got, want)


def test_headline_size_matches_c_oracle():
    """BASELINE headline shape (d=3, n=256, 2000 gates + N1 after each + 256 M) at a shot count the C oracle
    finishes in seconds; global-memory mode (one tableau = 262,656 B does not fit in shared memory)."""
    from oracle import c_oracle
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import noisy_random_clifford
    prog = compile_circuits([noisy_random_clifford(256, 2000, 3, prob=0.02)])
    shots, seed = 128, 2026
    eng, got = _run_gpu(prog, shots, seed, keep=True)
    want, fin = c_oracle.run(256, 3, prog.ops, shots, 0, seed, thresh24=prog.noise_thresh24,
                             channel=prog.noise_channel, want_final=True)
    assert np.array_equal(got, want)
    arrs = eng.export(eng.tableau, shots - 1)
    for key in ("x", "z", "p", "dx", "dz", "dp"):
        assert np.array_equal(arrs[key], fin[key]), key
    assert len({got[s].tobytes() for s in range(shots)}) > 8


def test_config2_random_clifford_n64_replay_first_shots():
    """BASELINE config 2: generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1)."""
    from oracle import c_oracle
    from sdim_b200 import generate_random_clifford_circuit
    from sdim_b200.ir import compile_circuits
    prog = compile_circuits([generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1)])
    assert prog.n_ops == 2064
    _, got = _run_gpu(prog, 64, 1)
    assert np.array_equal(got, c_oracle.run_philox(prog, 64, 0, 1))


def test_sharding_invariance_and_shot_offset():
    """Records depend on the GLOBAL shot id only: one call over [0,200) == calls over [0,77) + [77,200)."""
    from make_cases import random_program
    prog = random_program(seed=5, n=12, d=3, depth=300)
    _, whole = _run_gpu(prog, 200, 9)
    _, a = _run_gpu(prog, 77, 9, shot_offset=0)
    _, b = _run_gpu(prog, 123, 9, shot_offset=77)
    assert np.array_equal(whole, np.concatenate([a, b]))
    _, far = _run_gpu(prog, 4, 9, shot_offset=2 ** 33)       # 64-bit shot ids reach the counter's high word
    from oracle import c_oracle
    assert np.array_equal(far, c_oracle.run_philox(prog, 4, 2 ** 33, 9))


def test_host_buffer_entry_matches_device_path():
    from make_cases import random_program
    from sdim_b200.engine import simulate_host
    prog = random_program(seed=8, n=20, d=5, depth=400)
    _, dev = _run_gpu(prog, 300, 31)
    host, ms = simulate_host(prog, 300, 0, 31)
    assert np.array_equal(dev, host) and ms > 0
    host_g, _ = simulate_host(prog, 300, 0, 31, mode="global")
    assert np.array_equal(dev, host_g)


def test_edge_cases_empty_and_tiny():
    from sdim_b200.circuit import Circuit
    from sdim_b200.ir import compile_circuits
    from sdim_b200.engine import TableauEngine
    # no ops at all: fresh tableau comes back as |0...0>
    prog = compile_circuits([Circuit(3, 5)])
    eng = TableauEngine(prog)
    rec = eng.run(4, keep_tableau=True)
    assert rec.shape == (4, 0)
    arrs = eng.export(eng.tableau, 3)
    assert np.array_equal(arrs["z"], np.eye(3, dtype=np.int64)) and np.array_equal(arrs["dx"], np.eye(3, dtype=np.int64))
    assert not arrs["x"].any() and not arrs["dz"].any() and not arrs["p"].any() and not arrs["dp"].any()
    # zero shots
    c = Circuit(1, 2); c.add_gate("H", 0); c.add_gate("M", 0)
    prog = compile_circuits([c])
    assert TableauEngine(prog).run(0).shape == (0, 1)
    # single qubit H, M: random 0/1
    rec = TableauEngine(prog).run(4000, 0, 3).cpu().numpy()
    assert set(np.unique(rec)) == {0, 1} and abs((rec == 1).mean() - 0.5) < 0.05
    with pytest.raises(ValueError):
        TableauEngine(prog).run(4, mode="bogus")


def test_large_tableau_forced_resident_is_rejected():
    from sdim_b200.circuit import Circuit
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    c = Circuit(400, 3); c.add_gate("H", 0); c.add_gate("M", 0)
    with pytest.raises(ValueError, match="does not fit"):
        TableauEngine(compile_circuits([c])).run(2, mode="resident")
    rec = TableauEngine(compile_circuits([c])).run(2, mode="global").cpu().numpy()
    assert rec.shape == (2, 1) and not (rec & 0x80).any()


def test_many_shots_multiple_waves_are_race_free():
    """4096 headline-shaped shots (several shots per CTA, every SM busy) twice, bit-exact vs the C oracle:
    guards the shared-memory scratch reuse inside the measurement code against rare races."""
    from oracle import c_oracle
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import noisy_random_clifford
    prog = compile_circuits([noisy_random_clifford(256, 2000, 3)])
    want = c_oracle.run_philox(prog, 4096, 0, 2026)
    for _ in range(2):
        _, got = _run_gpu(prog, 4096, 2026)
        assert np.array_equal(got, want)
    from make_cases import random_program
    prog = random_program(seed=3, n=40, d=3, depth=1200)
    want = c_oracle.run_philox(prog, 20000, 0, 1)
    _, got = _run_gpu(prog, 20000, 1, "resident")
    assert np.array_equal(got, want)


@pytest.mark.parametrize("d,n,auto", [(3, 256, "planes-global"), (2, 300, "planes-resident"), (3, 440, "planes-global"),
                                      (3, 160, "planes-resident")])
def test_planes_kernel_large_resident(d, n, auto):
    """Bit-plane interpreter at the headline size and near the shared-memory limit, vs the C oracle, with the image
    in shared memory ("planes") and in scratch memory ("planes-global"); `auto` picks the plane interpreter for d in
    {2, 3}: resident while four or more CTAs fit per SM, on the global image below that."""
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import noisy_random_clifford
    prog = compile_circuits([noisy_random_clifford(n, 2500, d, prob=0.02)])
    eng = TableauEngine(prog)
    assert eng.plan(None) == (auto, False) and eng.plan("planes") == ("planes-resident", False)
    shots, seed = 600, 11
    want, fin = c_oracle.run(n, d, prog.ops, shots, 0, seed, thresh24=prog.noise_thresh24,
                             channel=prog.noise_channel, want_final=True)
    for mode in ("planes", "planes-global", None):
        got = eng.run(shots, 0, seed, keep_tableau=True, mode=mode).cpu().numpy()
        assert np.array_equal(got, want), mode
        arrs = eng.export(eng.tableau, shots - 1)
        for key in ("x", "z", "p", "dx", "dz", "dp"):
            assert np.array_equal(arrs[key], fin[key]), (mode, key)
    lanes = TableauEngine(prog).run(64, 0, seed, mode="lanes").cpu().numpy()
    assert np.array_equal(lanes, want[:64])


def test_plane_interpreter_without_scratch_stays_resident():
    """A C-ABI caller that passes no scratch (the slabs of the global image) still gets the resident plane interpreter
    where it fits: same records."""
    import ctypes as C
    import torch
    from sdim_b200 import _native as N
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import noisy_random_clifford
    prog = compile_circuits([noisy_random_clifford(256, 300, 3, prob=0.05)])
    eng = TableauEngine(prog)
    want = eng.run(32, 0, 5).cpu().numpy()
    rec = torch.zeros((32, prog.n_meas), dtype=torch.uint8, device="cuda")
    a = N.SdimbRunArgs()
    a.struct_size = C.sizeof(N.SdimbRunArgs)
    a.flags = N.FRESH
    a.n, a.d, a.shots, a.shot_offset = 256, 3, 32, 0
    a.ops, a.n_ops = eng.ops.data_ptr(), prog.n_ops
    a.records, a.n_meas, a.rec_stride = rec.data_ptr(), prog.n_meas, prog.n_meas
    a.noise_thresh24, a.noise_channel, a.n_noise = eng.noise_thresh.data_ptr(), eng.noise_channel.data_ptr(), prog.n_noise
    a.seed = 5
    a.stream = torch.cuda.current_stream().cuda_stream
    N.check(N.lib().sdimb_run(C.byref(a)))
    torch.cuda.synchronize()
    assert np.array_equal(rec.cpu().numpy(), want)


@pytest.mark.parametrize("d,n,depth", [(3, 500, 4000), (2, 700, 6000)])
def test_planes_on_a_global_image_beyond_the_shared_memory_limit(d, n, depth):
    """d = 2, 3 tableaus whose bit planes do not fit in shared memory run the same plane interpreter on an image in
    scratch memory (auto mode), bit-exact vs the C oracle incl. the final tableau, and identical to the uint8 lanes."""
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine
    prog = random_program(seed=11 * n + d, n=n, d=d, depth=depth, p_meas=0.03)
    eng = TableauEngine(prog)
    assert eng.plan(None) == ("planes-global", False) and eng.plan("lanes")[0] == "lanes-global"
    shots, seed = 24, 31
    got = eng.run(shots, 0, seed, keep_tableau=True, mode="planes-global").cpu().numpy()
    want, fin = c_oracle.run(n, d, prog.ops, shots, 0, seed, thresh24=prog.noise_thresh24,
                             channel=prog.noise_channel, want_final=True)
    assert np.array_equal(got, want)
    arrs = eng.export(eng.tableau, shots - 1)
    for key in ("x", "z", "p", "dx", "dz", "dp"):
        assert np.array_equal(arrs[key], fin[key]), key
    assert np.array_equal(eng.run(shots, 0, seed, mode="lanes").cpu().numpy(), want)
    assert np.array_equal(eng.run(shots, 0, seed).cpu().numpy(), want)      # auto: planes, or clusters for few shots


@pytest.mark.parametrize("min_np", ["0", "100000"])
def test_both_global_images_on_the_reference_goldens(golden_random, golden_config_sizes, monkeypatch, min_np):
    """The global-image plane interpreter exists twice: plain rows (the headline shape) and the interleaved image with
    merged measurement passes that sdimb_run picks for np >= 384.  SDIMB_PG_IL_MIN_NP = 0 / huge forces every
    planes-global run through one of them: reference goldens (all d = 2, 3 cases incl. the config sizes), records and
    all six final arrays."""
    import torch
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    monkeypatch.setenv("SDIMB_PG_IL_MIN_NP", min_np)
    cases = [(c["n"], c["d"], c["ops"], np.array(c["noise_ab"], dtype=np.uint8).reshape(-1, 2),
              np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in c["records"]], dtype=np.uint8),
              {k: np.array(v) for k, v in c["final"].items()}) for c in golden_random if c["d"] <= 3]
    cases += [(c["n"], c["d"], c["ops"], c["noise_ab"],
               np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in c["records"]], dtype=np.uint8), c["final"])
              for c in golden_config_sizes]
    assert len(cases) >= 40
    for n, d, ops, noise, want, final in cases:
        prog = compile_circuits([circuit_from_ops(n, d, ops)])
        eng = TableauEngine(prog)
        shots = 3
        rm = torch.from_numpy(np.tile((want & 0x7F)[None, :], (shots, 1)))
        rn = torch.from_numpy(np.tile(noise[None], (shots, 1, 1))) if prog.n_noise else None
        got = eng.run(shots, 0, 5, rm, rn, keep_tableau=True, mode="planes-global").cpu().numpy()
        assert all(np.array_equal(got[s], want) for s in range(shots)), (n, d, min_np)
        arrs = eng.export(eng.tableau, shots - 1)
        for key in ("x", "z", "p", "dx", "dz", "dp"):
            assert np.array_equal(arrs[key], final[key]), (n, d, key, min_np)


@pytest.mark.parametrize("d,n,depth", [(2, 33, 900), (3, 97, 2500), (3, 256, 4000), (2, 700, 6000), (3, 450, 5000)])
def test_interleaved_image_matches_c_oracle_free_running(monkeypatch, d, n, depth):
    """Interleaved image forced at every size: all opcodes incl. M_X / RESET / SWAP and the three noise channels,
    ragged n, Philox draws; records of all shots and the final tableau vs the C oracle."""
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine
    monkeypatch.setenv("SDIMB_PG_IL_MIN_NP", "0")
    prog = random_program(seed=700 * d + n, n=n, d=d, depth=depth)
    eng = TableauEngine(prog)
    shots, seed = 40, 77
    got = eng.run(shots, 0, seed, keep_tableau=True, mode="planes-global").cpu().numpy()
    want, fin = c_oracle.run(n, d, prog.ops, shots, 0, seed, thresh24=prog.noise_thresh24,
                             channel=prog.noise_channel, want_final=True)
    assert np.array_equal(got, want)
    arrs = eng.export(eng.tableau, shots - 1)
    for key in ("x", "z", "p", "dx", "dz", "dp"):
        assert np.array_equal(arrs[key], fin[key]), key


def _launches():
    from sdim_b200 import _native as N
    return int(N.lib().sdimb_launch_count())


@pytest.mark.parametrize("min_run", ["2", "1000000"])
def test_tail_run_kernel_on_the_reference_goldens(golden_random, golden_config_sizes, monkeypatch, min_run):
    """The run of M ops that ends a stream executes in run_tail_kernel (one warp per shot on a generator-major copy
    of the image, planes_gm.cuh) when it is at least SDIMB_GM_MIN_RUN long and the tableau is not kept.  2 sends the
    tail of every reference golden there (d = 2, 3, incl. the config sizes), a huge value none: the records must equal
    the reference's either way, and the launch count tells which path ran."""
    import torch
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    monkeypatch.setenv("SDIMB_GM_MIN_RUN", min_run)
    cases = [(c["n"], c["d"], c["ops"], np.array(c["noise_ab"], dtype=np.uint8).reshape(-1, 2),
              np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in c["records"]], dtype=np.uint8))
             for c in golden_random if c["d"] <= 3]
    cases += [(c["n"], c["d"], c["ops"], c["noise_ab"],
               np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in c["records"]], dtype=np.uint8))
              for c in golden_config_sizes]
    two_kernels = 0
    for n, d, ops, noise, want in cases:
        prog = compile_circuits([circuit_from_ops(n, d, ops)])
        eng = TableauEngine(prog)
        shots = 37
        rm = torch.from_numpy(np.tile((want & 0x7F)[None, :], (shots, 1)))
        rn = torch.from_numpy(np.tile(noise[None], (shots, 1, 1))) if prog.n_noise else None
        before = _launches()
        got = eng.run(shots, 0, 5, rm, rn, mode="planes-global").cpu().numpy()
        two_kernels += int(_launches() - before == 2)
        assert (_launches() - before == 2) == (eng.tail_run_len > 0)
        assert all(np.array_equal(got[s], want) for s in range(shots)), (n, d, min_run)
    assert (two_kernels >= 20) if min_run == "2" else (two_kernels == 0)


def test_host_entry_writes_pinned_output_directly():
    """sdimb_simulate_host with `records` in pinned host memory (the device writes it directly, no staging copy), in a
    caller-provided pageable array, and in a fresh one: the same records; bad `out` arrays are rejected."""
    import torch
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import simulate_host
    prog = random_program(seed=12, n=70, d=3, depth=500)
    want = c_oracle.run_philox(prog, 700, 5, 3)
    fresh, _ = simulate_host(prog, 700, 5, 3)
    pinned = torch.zeros((700, prog.n_meas), dtype=torch.uint8, pin_memory=True).numpy()
    got, _ = simulate_host(prog, 700, 5, 3, out=pinned)
    assert got is pinned
    pageable = np.zeros((700, prog.n_meas), dtype=np.uint8)
    simulate_host(prog, 700, 5, 3, out=pageable)
    assert np.array_equal(fresh, want) and np.array_equal(pinned, want) and np.array_equal(pageable, want)
    with pytest.raises(ValueError):
        simulate_host(prog, 700, 5, 3, out=np.zeros((699, prog.n_meas), dtype=np.uint8))
    with pytest.raises(ValueError):
        simulate_host(prog, 700, 5, 3, out=np.zeros((700, prog.n_meas), dtype=np.int32))


@pytest.mark.parametrize("d,n", [(3, 130), (2, 200), (3, 256), (5, 256), (7, 100)])
def test_two_kernel_paths_edge_cases(d, n):
    """Degenerate streams through the two-kernel paths (gate streams + generator-major tail for d = 2, 3; lane
    interpreter + byte tail for d >= 5): nothing but the final measurement (empty front), only Pauli gates and noise in
    front of it, a single shot, one shot more than a full wave of CTAs — records vs the C oracle."""
    from oracle import c_oracle
    from sdim_b200.circuit import Circuit
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    mode = None if d <= 3 else "global"
    only_m = Circuit(n, d)
    only_m.add_gate("M", list(range(n)))
    paulis = Circuit(n, d)
    for q in range(0, n, 3):
        paulis.add_gate("X", q)
        paulis.add_gate("Z_INV", (q + 1) % n)
        paulis.add_gate("N1", q, prob=0.7, noise_channel="d")
    paulis.add_gate("M", list(range(n)))
    mixed = Circuit(n, d)
    mixed.add_gate("H", list(range(n)))                # every final measurement is random
    mixed.add_gate("CNOT", 0, 1)
    mixed.add_gate("M", list(range(n)))
    for circ in (only_m, paulis, mixed):
        prog = compile_circuits([circ])
        eng = TableauEngine(prog)
        for shots in (1, 445):
            before = _launches()
            got = eng.run(shots, 2, 33, mode=mode).cpu().numpy()
            assert _launches() - before == 2
            want, _ = c_oracle.run(n, d, prog.ops, shots, 2, 33, thresh24=prog.noise_thresh24, channel=prog.noise_channel)
            assert np.array_equal(got, want)


@pytest.mark.parametrize("form", ["tail8", "tail8-tma", "one-kernel", "keep"])
def test_lane_kernels_on_the_reference_goldens(golden_lanes_sizes, monkeypatch, form):
    """Outputs of the UNMODIFIED reference at d = 5, 7, 11 and n = 97 ... 256 (tests/golden/lanes_sizes.npz) replayed into
    the uint8-lane kernels: the two-kernel path (lane interpreter + run_tail8_kernel, with ordinary loads and with TMA
    staging), the interpreter alone, and — tableau kept — all six final arrays, in every mode that holds the shape."""
    import torch
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    if form == "tail8-tma":
        monkeypatch.setenv("SDIMB_TAIL8_TMA", "1")
    if form == "one-kernel":
        monkeypatch.setenv("SDIMB_NO_TAIL8", "1")
    for case in golden_lanes_sizes:
        n, d, ops = case["n"], case["d"], case["ops"]
        prog = compile_circuits([circuit_from_ops(n, d, ops)])
        eng = TableauEngine(prog)
        want = np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in case["records"]], dtype=np.uint8)
        shots = 6
        rm = torch.from_numpy(np.tile((want & 0x7F)[None, :], (shots, 1)))
        rn = torch.from_numpy(np.tile(case["noise_ab"][None], (shots, 1, 1)))
        if form == "keep":
            for mode in ("global", "resident", "global-cta"):
                try:
                    eng.plan(mode)
                except ValueError:
                    continue
                got = eng.run(shots, 0, 99, rm, rn, keep_tableau=True, mode=mode).cpu().numpy()
                assert all(np.array_equal(got[s], want) for s in range(shots)), (case["name"], mode)
                arrs = eng.export(eng.tableau, shots - 1)
                for key in ("x", "z", "p", "dx", "dz", "dp"):
                    assert np.array_equal(arrs[key], case["final"][key]), (case["name"], mode, key)
            continue
        before = _launches()
        got = eng.run(shots, 0, 99, rm, rn).cpu().numpy()              # auto: HBM store (n >= 96), tail run in the second kernel
        assert _launches() - before == (1 if form == "one-kernel" else 2), case["name"]
        assert all(np.array_equal(got[s], want) for s in range(shots)), (case["name"], form)


@pytest.mark.parametrize("off", [False, True, "tma"])
@pytest.mark.parametrize("d,n,depth", [(5, 256, 2500), (7, 97, 1500), (3, 130, 1200), (2, 200, 1500), (13, 40, 900),
                                       (127, 64, 700), (5, 500, 900), (11, 5, 300), (5, 512, 600)])
def test_lanes_tail_run_kernel_matches_c_oracle(monkeypatch, d, n, depth, off):
    """uint8 lanes on the HBM store ("global"): the run of M ops that ends the stream executes in run_tail8_kernel — one
    warp per shot on a generator-major byte copy (lanes_gm.cuh) — for every prime d <= 127, ragged n up to 512, all
    opcodes incl. mid-circuit M / M_X / RESET and noise in front of it, Philox draws and replayed outcomes; the records of
    every shot vs the C oracle, with the second kernel and without it (SDIMB_NO_TAIL8), and from the host-buffer entry."""
    import torch
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine, simulate_host
    if off == "tma":               # the row lists staged by TMA bulk copies (cp.async.bulk + mbarrier) instead of loads
        monkeypatch.setenv("SDIMB_TAIL8_TMA", "1")
        off = False
    elif off:
        monkeypatch.setenv("SDIMB_NO_TAIL8", "1")
    prog = random_program(seed=500 * d + n, n=n, d=d, depth=depth)
    eng = TableauEngine(prog)
    assert eng.tail_run_len_raw >= n
    shots, seed = (600 if n <= 130 else 150), 19
    before = _launches()
    got = eng.run(shots, 7, seed, mode="global").cpu().numpy()
    assert _launches() - before == (1 if off else 2)
    want, _ = c_oracle.run(n, d, prog.ops, shots, 7, seed, thresh24=prog.noise_thresh24, channel=prog.noise_channel)
    assert np.array_equal(got, want)
    host, _ = simulate_host(prog, 40, 7, seed, mode="global")
    assert np.array_equal(host, want[:40])
    # replayed outcomes and noise
    k = 30
    rng = np.random.default_rng(n + d)
    rm = rng.integers(0, d, size=(k, prog.n_meas), dtype=np.uint8)
    rn = rng.integers(0, d, size=(k, prog.n_noise, 2), dtype=np.uint8)
    got = eng.run(k, 0, seed, torch.from_numpy(rm), torch.from_numpy(rn), mode="global").cpu().numpy()
    want, _ = c_oracle.run(n, d, prog.ops, k, 0, seed, replay_meas=rm, replay_noise=rn)
    assert np.array_equal(got, want)
    # the tableau is kept: one kernel, same records, final tableau vs the oracle
    before = _launches()
    got = eng.run(8, 0, seed, keep_tableau=True, mode="global").cpu().numpy()
    assert _launches() - before == 1
    want, fin = c_oracle.run(n, d, prog.ops, 8, 0, seed, thresh24=prog.noise_thresh24, channel=prog.noise_channel, want_final=True)
    assert np.array_equal(got, want)
    arrs = eng.export(eng.tableau, 7)
    assert all(np.array_equal(arrs[key], fin[key]) for key in ("x", "z", "p", "dx", "dz", "dp"))


@pytest.mark.parametrize("form", ["smem-8", "smem-4", "smem-3", "global", "off"])
@pytest.mark.parametrize("d,n,depth,il", [(3, 256, 2500, "100000"), (2, 300, 3000, "100000"), (3, 97, 1500, "0"),
                                          (2, 33, 900, "100000"), (3, 500, 1200, "0"), (2, 512, 1500, "100000"),
                                          (3, 40, 800, "100000"), (3, 5, 300, "100000")])
def test_gate_stream_kernel_matches_c_oracle(monkeypatch, d, n, depth, il, form):
    """Gates and the three noise channels in front of the final all-qudit measurement: the front part runs as
    pre-decoded per-warp streams (gate_stream_kernel, planes_stream.cuh) — image in shared memory with 8 / 4 / 3 warps
    per shot, or on the per-shot global image — and must give the records of the C oracle for every shot, like the
    interpreted form ("off").  Also: N1 events replayed instead of drawn, and a run that continues from a uint8 store."""
    import torch
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine, simulate_host
    monkeypatch.setenv("SDIMB_GM_MIN_RUN", "2")
    monkeypatch.setenv("SDIMB_PG_IL_MIN_NP", il)
    if form == "off":
        monkeypatch.setenv("SDIMB_NO_GATE_STREAM", "1")
    elif form == "global":
        monkeypatch.setenv("SDIMB_GS_GLOBAL", "1")
    else:
        monkeypatch.setenv("SDIMB_GS_WARPS", form.split("-")[1])
    prog = random_program(seed=77 * d + n, n=n, d=d, depth=depth, p_meas=0.0, p_noise=0.25)
    eng = TableauEngine(prog)
    assert eng.tail_run_len == n and eng.gate_stream is not None      # "off": compiled, but the library ignores it
    shots, seed = (3000 if n <= 97 else 400), 41
    before = _launches()
    got = eng.run(shots, 11, seed, mode="planes-global").cpu().numpy()
    assert _launches() - before == 2
    want, _ = c_oracle.run(n, d, prog.ops, shots, 11, seed, thresh24=prog.noise_thresh24, channel=prog.noise_channel)
    assert np.array_equal(got, want)
    if n >= 130:                                   # the host-buffer entry compiles its own streams (auto: planes-global)
        host, _ = simulate_host(prog, 64, 11, seed)
        assert np.array_equal(host, want[:64])
    # replayed N1 exponents: every event fires
    k = 50
    rng = np.random.default_rng(n)
    rn = rng.integers(0, d, size=(k, prog.n_noise, 2), dtype=np.uint8)
    got = eng.run(k, 0, seed, replay_noise=torch.from_numpy(rn), mode="planes-global").cpu().numpy()
    want, _ = c_oracle.run(n, d, prog.ops, k, 0, seed, replay_noise=rn)
    assert np.array_equal(got, want)
    # continue from a store: |0...0> written by sdimb_init, packed into bit planes by the kernel
    tab = eng.alloc_tableau(k)
    eng.init_tableau(tab)
    got = eng.run(k, 0, seed, mode="planes-global", tableau=tab, fresh=False).cpu().numpy()
    want, _ = c_oracle.run(n, d, prog.ops, k, 0, seed, thresh24=prog.noise_thresh24, channel=prog.noise_channel)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("il", ["0", "100000"])
@pytest.mark.parametrize("d,n,depth", [(2, 33, 900), (3, 97, 2500), (3, 256, 3000), (2, 300, 3000), (3, 500, 3000),
                                       (2, 512, 2500), (3, 17, 600), (3, 1, 40), (2, 2, 60)])
def test_tail_run_kernel_matches_c_oracle_free_running(monkeypatch, d, n, depth, il):
    """All opcodes incl. mid-circuit M / M_X / RESET and the three noise channels in front of the final all-qudit
    measurement, ragged n up to the 512-qudit limit of the generator-major path, both global images, Philox draws,
    more shots than one wave of warps: the records of every shot vs the C oracle; the same records without the
    second kernel (tableau kept) and from the host-buffer entry."""
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine, simulate_host
    monkeypatch.setenv("SDIMB_GM_MIN_RUN", "2")
    monkeypatch.setenv("SDIMB_PG_IL_MIN_NP", il)
    prog = random_program(seed=900 * d + n, n=n, d=d, depth=depth)
    eng = TableauEngine(prog)
    assert eng.tail_run_len >= n or n < 2          # a single M is not a run
    shots, seed = (5000 if n <= 97 else 300), 78
    before = _launches()
    got = eng.run(shots, 3, seed, mode="planes-global").cpu().numpy()
    assert _launches() - before == (2 if eng.tail_run_len else 1)
    want, _ = c_oracle.run(n, d, prog.ops, shots, 3, seed, thresh24=prog.noise_thresh24, channel=prog.noise_channel)
    assert np.array_equal(got, want)
    assert np.array_equal(eng.run(64, 3, seed, keep_tableau=True, mode="planes-global").cpu().numpy(), want[:64])
    rec, _ = simulate_host(prog, 64, 3, seed, mode="planes-global")
    assert np.array_equal(rec, want[:64])


def test_planes_continue_from_store_and_stepped():
    """!FRESH path of the plane kernel (pack from / unpack to the uint8 store): op-by-op stepping on a persistent
    store gives the same records and final tableau as one fused launch."""
    from make_cases import random_program
    from sdim_b200.engine import TableauEngine
    import torch
    for d in (2, 3):
        prog = random_program(seed=50 + d, n=37, d=d, depth=300)
        eng = TableauEngine(prog)
        fused = eng.run(5, 0, 3, keep_tableau=True).cpu().numpy()
        fused_tab = eng.tableau.clone()
        store = eng.alloc_tableau(5)
        eng.init_tableau(store)
        rec = torch.zeros((5, prog.n_meas), dtype=torch.uint8, device="cuda")
        for i in range(prog.n_ops):
            eng.run(5, 0, 3, keep_tableau=True, tableau=store, fresh=False, op_range=(i, i + 1), records=rec)
        assert np.array_equal(rec.cpu().numpy(), fused)
        assert torch.equal(store, fused_tab)


def test_config3_surface_code_d7_matches_c_oracle():
    """BASELINE config 3 (synthesised): rotated surface code distance 7, d = 2, 97 qubits, depolarising
    N1 on data qubits, RESET on ancillas; bit-exact records vs the C oracle, and a noiseless run is all zeros."""
    from oracle import c_oracle
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import rotated_surface_code
    prog = compile_circuits([rotated_surface_code(7, 7, prob=0.01)])
    assert prog.num_qudits == 97 and prog.n_meas == 7 * 48 + 49
    shots = 3000
    for mode in (None, "lanes"):
        _, got = _run_gpu(prog, shots, 2026, mode)
        assert np.array_equal(got, c_oracle.run_philox(prog, shots, 0, 2026))
    clean = compile_circuits([rotated_surface_code(7, 7, prob=0.0)])
    _, rec = _run_gpu(clean, 500, 5)
    vals = rec & 0x7F
    # first-round X-ancilla outcomes are random, every later syndrome repeats it; Z syndromes and data are 0/consistent:
    first, later = vals[:, :48], vals[:, 48:7 * 48].reshape(500, 6, 48)
    assert (later == first[:, None, :]).all()


def test_config4_qutrit_repetition_code_matches_c_oracle():
    """BASELINE config 4 (generalised examples/repetition_code.ipynb): qutrit repetition code distance 25, 25 rounds,
    flip noise + RESET, 625 records per shot."""
    from oracle import c_oracle
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import qudit_repetition_code
    prog = compile_circuits([qudit_repetition_code(25, 25, 3, prob=0.01)])
    assert prog.num_qudits == 49 and prog.n_meas == 625
    shots = 4000
    _, got = _run_gpu(prog, shots, 7)
    assert np.array_equal(got, c_oracle.run_philox(prog, shots, 0, 7))
    clean = compile_circuits([qudit_repetition_code(25, 25, 3, prob=0.0)])
    _, rec = _run_gpu(clean, 300, 1)
    assert (rec == 0x80).all()           # noiseless: every syndrome and data outcome is a deterministic 0
    # with flip noise the final data readout differs from 0 somewhere and syndromes fire
    assert ((got & 0x7F) != 0).any()


def test_config5_large_single_tableau():
    """BASELINE config 5 shape at a size the oracle finishes quickly (n = 1024, d = 5 and 7, one shot): the
    uint8-lane interpreter on the HBM store, rows spanning several words per thread."""
    from oracle import c_oracle
    from sdim_b200 import generate_random_clifford_circuit
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    for d in (5, 7):
        prog = compile_circuits([generate_random_clifford_circuit(1024, 4096, d, measurement_rounds=1, seed=1)])
        eng = TableauEngine(prog)
        assert eng.plan(None)[0] == "lanes-global"
        got = eng.run(2, 0, 3, keep_tableau=True).cpu().numpy()
        want, fin = c_oracle.run(1024, d, prog.ops, 2, 0, 3, want_final=True)
        assert np.array_equal(got, want)
        arrs = eng.export(eng.tableau, 1)
        for key in ("x", "z", "p", "dx", "dz", "dp"):
            assert np.array_equal(arrs[key], fin[key]), key


@pytest.mark.parametrize("d", [5, 7])
def test_config5_at_full_size_n4096(d):
    """BASELINE config 5 AT ITS STATED SIZE: generate_random_clifford_circuit(4096, 8192, d, measurement_rounds=1,
    seed=1), one 64 MiB tableau, d = 5 and 7 — the cluster interpreter (what auto picks) and one CTA per shot, records
    and all six final arrays against the C oracle (0.15 s per shot on one host core)."""
    from oracle import c_oracle
    from sdim_b200 import generate_random_clifford_circuit
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    n = 4096
    prog = compile_circuits([generate_random_clifford_circuit(n, 2 * n, d, measurement_rounds=1, seed=1)])
    assert prog.n_ops == 3 * n and prog.n_meas == n
    want, fin = c_oracle.run(n, d, prog.ops, 1, 0, 3, want_final=True)
    assert ((want & 0x80) != 0).any() and ((want & 0x80) == 0).any()
    eng = TableauEngine(prog)
    assert eng.plan(None)[0] == "lanes-global" and eng.cluster_size(1, None) >= 8
    for mode in (None, "global-cta"):
        got = eng.run(1, 0, 3, keep_tableau=True, mode=mode).cpu().numpy()
        assert np.array_equal(got, want), mode
        arrs = eng.export(eng.tableau, 0)
        for key in ("x", "z", "p", "dx", "dz", "dp"):
            assert np.array_equal(arrs[key], fin[key]), (mode, key)


# ---------------------------------------------------------------------------------------------------
# Cluster interpreter (one shot per thread-block cluster, sdim_b200/csrc/clusters.cuh)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("csize", [1, 2, 8, 16])
@pytest.mark.parametrize("d,n,depth", [(2, 97, 2500), (3, 100, 2500), (5, 33, 800), (7, 70, 1500), (13, 300, 3000)])
def test_cluster_interpreter_matches_c_oracle(d, n, depth, csize, monkeypatch):
    """Every opcode, all noise channels, ragged n, for several cluster sizes (lane words and rows are dealt over the
    CTAs differently for each): records of all shots and the final tableau, bit-exact vs the C oracle."""
    from make_cases import random_program
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine
    monkeypatch.setenv("SDIMB_CLUSTER_SIZE", str(csize))
    prog = random_program(seed=3000 * d + n, n=n, d=d, depth=depth)
    eng = TableauEngine(prog)
    shots, seed = 40, 99 + d
    assert eng.plan("cluster")[0] == "lanes-global"
    assert eng.cluster_size(shots, "cluster") == csize
    assert eng.cluster_size(shots, "global-cta") == 0
    got = eng.run(shots, 0, seed, mode="cluster", keep_tableau=True).cpu().numpy()
    want, fin = c_oracle.run(n, d, prog.ops, shots, 0, seed, thresh24=prog.noise_thresh24,
                             channel=prog.noise_channel, want_final=True)
    assert np.array_equal(got, want)
    arrs = eng.export(eng.tableau, shots - 1)
    for key in ("x", "z", "p", "dx", "dz", "dp"):
        assert np.array_equal(arrs[key], fin[key]), key


@pytest.mark.parametrize("d,n", [(5, 150), (2, 97), (7, 700)])
def test_cluster_runs_of_measurements(d, n):
    """Runs of M ops with no gate in between are where the cluster interpreter drops its barriers (deterministic
    measurements) and prefetches rows: repeated rounds of all-qudit measurement (second round fully deterministic),
    duplicates inside a run, RESETs and noise splitting runs, factor lists of several generators; bit-exact vs the C
    oracle incl. the final tableau."""
    import random
    from oracle import c_oracle
    from sdim_b200.circuit import Circuit
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    from make_cases import ONE, TWO
    rng = random.Random(17 * n + d)
    c = Circuit(n, d)
    for _ in range(6 * n):
        if rng.random() < 0.45:
            a, b = rng.sample(range(n), 2)
            c.add_gate(rng.choice(TWO), a, b)
        else:
            c.add_gate(rng.choice(ONE), rng.randrange(n))
    order = list(range(n))
    c.add_gate("M", order)                                   # random and deterministic outcomes interleaved
    rng.shuffle(order)
    c.add_gate("M", order)                                   # all deterministic now: long runs
    c.add_gate("M", [order[0], order[0], order[1]])          # the same qudit twice in one run
    for q in order[:10]:
        c.add_gate("RESET", q)
    c.add_gate("N1", order[3], prob=0.5, noise_channel="d")
    c.add_gate("H", order[5])
    c.add_gate("M", order[:40])
    prog = compile_circuits([c])
    shots, seed = 5, 77
    want, fin = c_oracle.run(n, d, prog.ops, shots, 0, seed, thresh24=prog.noise_thresh24,
                             channel=prog.noise_channel, want_final=True)
    eng = TableauEngine(prog)
    got = eng.run(shots, 0, seed, mode="cluster", keep_tableau=True).cpu().numpy()
    assert np.array_equal(got, want)
    arrs = eng.export(eng.tableau, shots - 1)
    for key in ("x", "z", "p", "dx", "dz", "dp"):
        assert np.array_equal(arrs[key], fin[key]), key


def test_cluster_continue_from_store_and_stepped(monkeypatch):
    """!FRESH path and the unlayered stream of the cluster interpreter: stepping through the op stream in chunks on
    a persistent store gives the records and the store of one fused launch (and of the one-CTA kernel)."""
    from make_cases import random_program
    from sdim_b200.engine import TableauEngine
    import torch
    monkeypatch.setenv("SDIMB_CLUSTER_SIZE", "4")
    prog = random_program(seed=61, n=45, d=5, depth=400)
    eng = TableauEngine(prog)
    fused = eng.run(6, 0, 3, keep_tableau=True, mode="cluster").cpu().numpy()
    fused_tab = eng.tableau.clone()
    cta = eng.run(6, 0, 3, keep_tableau=True, mode="global-cta").cpu().numpy()
    assert np.array_equal(fused, cta) and torch.equal(eng.tableau, fused_tab)
    store = eng.alloc_tableau(6)
    eng.init_tableau(store)
    rec = torch.zeros((6, prog.n_meas), dtype=torch.uint8, device="cuda")
    for lo in range(0, prog.n_ops, 7):
        eng.run(6, 0, 3, keep_tableau=True, tableau=store, fresh=False, op_range=(lo, min(lo + 7, prog.n_ops)),
                records=rec, mode="cluster")
    assert np.array_equal(rec.cpu().numpy(), fused)
    assert torch.equal(store, fused_tab)


def test_few_shots_of_a_large_qubit_tableau_take_the_cluster_path():
    """d = 2, 3 beyond the shared-memory limit: many shots run bit planes on a global image, a few shots run one uint8
    tableau per thread-block cluster (auto decides from the shot count); same records either way."""
    from oracle import c_oracle
    from sdim_b200 import generate_random_clifford_circuit
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    prog = compile_circuits([generate_random_clifford_circuit(1024, 3000, 3, measurement_rounds=1, seed=2)])
    eng = TableauEngine(prog)
    assert eng.plan(None)[0] == "planes-global"
    assert eng._auto_mode(None, 2) == "lanes" and eng._auto_mode(None, 5000) is None
    few = eng.run(2, 0, 9).cpu().numpy()
    assert np.array_equal(few, eng.run(2, 0, 9, mode="planes-global").cpu().numpy())
    assert np.array_equal(few, c_oracle.run_philox(prog, 2, 0, 9))


def test_cluster_is_the_default_for_a_large_single_tableau():
    """Config 5 shape (n = 1024 here): few shots of a wide tableau go to the largest cluster size of which one per
    shot is resident at once; many shots keep one CTA per shot; both agree with each other (the oracle comparison is
    test_config5_large_single_tableau)."""
    from sdim_b200 import generate_random_clifford_circuit
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    prog = compile_circuits([generate_random_clifford_circuit(1024, 4096, 5, measurement_rounds=1, seed=1)])
    eng = TableauEngine(prog)
    sizes = [eng.cluster_size(s) for s in (1, 4, 12, 30, 60, 1000)]
    assert sizes[0] == 16 and sizes[-1] == 0                 # one shot: the largest cluster; many shots: one CTA each
    assert all(a >= b for a, b in zip(sizes, sizes[1:])) and set(sizes) <= {16, 8, 4, 2, 0}
    a = eng.run(3, 0, 3).cpu().numpy()
    b = eng.run(3, 0, 3, mode="global-cta").cpu().numpy()
    assert np.array_equal(a, b)


def test_config2_full_size_ten_thousand_shots():
    """BASELINE config 2 at its full size: random Clifford d = 3, n = 64, depth 2k + all-qudit measurement, 10^4 shots.
    All 10^4 x 64 records bit-exact vs the C oracle (same Philox counters), plus the size-independent properties:
    determinism flags do not depend on the shot, random outcomes are uniform (chi-square per record)."""
    from oracle import c_oracle
    from sdim_b200 import generate_random_clifford_circuit
    from sdim_b200.ir import compile_circuits
    prog = compile_circuits([generate_random_clifford_circuit(64, 2000, 3, measurement_rounds=1, seed=1)])
    shots = 10000
    _, got = _run_gpu(prog, shots, 2026)
    assert np.array_equal(got, c_oracle.run_philox(prog, shots, 0, 2026))
    det = (got & 0x80) != 0
    assert (det == det[0]).all()
    vals = got & 0x7F
    for k in np.nonzero(~det[0])[0]:
        counts = np.bincount(vals[:, k], minlength=3)
        chi2 = ((counts - shots / 3) ** 2 / (shots / 3)).sum()
        assert chi2 < 30.0, (k, counts)          # 2 dof; 30 is far beyond any plausible fluctuation over 64 records


def test_config3_and_config4_at_a_million_shots():
    """BASELINE configs 3 and 4 at 10^6 shots per GPU (config 3's full count; one tenth of config 4's 10^7, which
    Program runs as waves of this size): the first 2 000 shots are bit-exact vs the C oracle, and over all shots the
    size-independent properties hold — determinism flags do not depend on the shot, noiseless runs reproduce."""
    import torch
    from oracle import c_oracle
    from sdim_b200.engine import TableauEngine
    from sdim_b200.ir import compile_circuits
    from sdim_b200.workloads import qudit_repetition_code, rotated_surface_code
    shots = 1_000_000
    for circ in (rotated_surface_code(7, 7, prob=1e-3), qudit_repetition_code(25, 25, 3, prob=1e-2)):
        prog = compile_circuits([circ])
        rec = TableauEngine(prog).run(shots, 0, 2026)
        head = rec[:2000].cpu().numpy()
        assert np.array_equal(head, c_oracle.run_philox(prog, 2000, 0, 2026))
        det = (rec & 0x80) != 0
        assert bool((det == det[0]).all())
        vals = rec & 0x7F
        assert int(vals.max()) < prog.dimension
        if prog.dimension == 2:
            # surface code: detection events (a syndrome differing from the previous round's) are rare but present
            synd = vals[:, : 7 * 48].reshape(shots, 7, 48)
            rate = (synd[:, 1:] != synd[:, :-1]).float().mean().item()
            assert 1e-4 < rate < 0.05, rate
        else:
            # repetition code: flip errors accumulate, so a growing but bounded fraction of syndromes is non-zero
            rate = (vals[:, :600] != 0).float().mean().item()
            assert 1e-3 < rate < 0.6, rate
        del rec, det, vals
        torch.cuda.empty_cache()
