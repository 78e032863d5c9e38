"""bench.py contract checks that need no GPU: the reference arm's JSON line (bounded sample), the byte accounting
of the roofline against SURVEY 8d, and the workload definition."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--cpu-shots", "24"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "shot_gates_per_sec" and line["unit"] == "shot*gates/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["gpu_launches"] == 0 and line["dtype"] == "u8" and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "n=256" in line["config"]["workload"] and line["config"]["gates_per_shot"] == 4875


def test_headline_workload_matches_survey_8d():
    sys.path.insert(0, ROOT)
    import bench
    circ, prog = bench.build_workload()
    assert (prog.num_qudits, prog.dimension) == (256, 3)
    assert prog.n_user_gates == len(circ.operations) == 4875            # 2000 gates + 2619 N1 + 256 M
    assert prog.n_meas == 256 and prog.n_noise == 2619
    assert np.all(prog.noise_prob == 1e-3) and np.all(prog.noise_channel == 0)


def test_algorithmic_bytes_follow_the_survey_table():
    """SURVEY 8d per-op bytes at N = 2n lanes: H 6N, P 5N, CNOT 6N, CZ 8N, SWAP 8N, Paulis 3N (odd d); dense random
    measurement 8n^2 + 2N + 1, dense deterministic 2n^2 + 2n + 1; d = 2 counts bits (1/8, phases 2/8)."""
    sys.path.insert(0, ROOT)
    import bench
    from sdim_b200.circuit import Circuit
    from sdim_b200.ir import compile_circuits

    def one(d, n, build):
        c = Circuit(n, d)
        build(c)
        return compile_circuits([c])

    n, N = 10, 20
    for name, want in (("H", 6 * N), ("H_INV", 6 * N), ("P", 5 * N), ("P_INV", 5 * N), ("X", 3 * N), ("Z_INV", 3 * N)):
        assert bench.algorithmic_bytes_per_shot(one(3, n, lambda c: c.add_gate(name, 1)), [], None) == want
    for name, want in (("CNOT", 6 * N), ("CZ", 8 * N), ("CZ_INV", 8 * N), ("SWAP", 8 * N)):
        assert bench.algorithmic_bytes_per_shot(one(3, n, lambda c: c.add_gate(name, 1, 2)), [], None) == want
    prog = one(3, n, lambda c: c.add_gate("M", 0))
    assert bench.algorithmic_bytes_per_shot(prog, [False], None) == 8 * n * n + 2 * N + 1
    assert bench.algorithmic_bytes_per_shot(prog, [True], None) == 2 * n * n + 2 * n + 1
    # the sparse rule never counts more than the dense one
    assert bench.algorithmic_bytes_per_shot(prog, [False], [n - 1]) <= 8 * n * n + 2 * N + 1 + 8 * n
    assert bench.algorithmic_bytes_per_shot(one(2, 16, lambda c: c.add_gate("H", 1)), [], None) == 4 * 32 / 8 + 2 * 32 * 2 / 8
