"""A third, independent line of validation: outcome DISTRIBUTIONS of the tableau path against a dense
state vector built from the reference's gate matrices (oracle/statevector_oracle.py, restating sdim/unitary.py:26-125).
Restates the reference's tests/test_tomography.py:13-77 (which needs Cirq), and sharpens it: on CPU the distribution
of the tableau oracle is computed EXACTLY by replaying every combination of measurement draws, so the comparison is
an equality, not a total-variation bound on 800 samples."""
import itertools
import random

import numpy as np
import pytest

from make_cases import ONE, TWO
from sdim_b200.circuit import Circuit
from sdim_b200.ir import compile_circuits


def _unitary_circuit(seed, n, d, depth):
    rng = random.Random(seed)
    c = Circuit(n, d)
    for _ in range(depth):
        if rng.random() < 0.4:
            a, b = rng.sample(range(n), 2)
            c.add_gate(rng.choice(TWO), a, b)
        else:
            c.add_gate(rng.choice(ONE), rng.randrange(n))
    c.add_gate("M", list(range(n)))
    return c


def _exact_tableau_distribution(prog):
    """Every combination of draws for the n terminal measurements, replayed through the C oracle: each combination
    has probability d^-n; deterministic measurements ignore their draw."""
    from oracle import c_oracle
    n, d = prog.num_qudits, prog.dimension
    combos = np.array(list(itertools.product(range(d), repeat=n)), dtype=np.uint8)
    rec, _ = c_oracle.run(n, d, prog.ops, len(combos), 0, 0, replay_meas=combos)
    dist = np.zeros([d] * n)
    for row in rec & 0x7F:
        dist[tuple(int(v) for v in row)] += 1.0 / len(combos)
    return dist


@pytest.mark.parametrize("d", [2, 3, 5, 7])
@pytest.mark.parametrize("depth", [5, 15, 50, 200])
def test_tableau_oracle_distribution_equals_state_vector(d, depth):
    from oracle.statevector_oracle import outcome_distribution
    n = 3
    for i in range(12):
        prog = compile_circuits([_unitary_circuit(7000 * d + 13 * depth + i, n, d, depth)])
        want = outcome_distribution(n, d, prog.ops)
        assert abs(want.sum() - 1.0) < 1e-9
        got = _exact_tableau_distribution(prog)
        assert np.allclose(got, want, atol=1e-9), f"d={d} depth={depth} circuit {i}"


def test_state_vector_matrices_are_the_reference_conventions():
    """Known answers that pin the conventions: X raises, Z|j> = w^j|j>, H X H^-1 = Z, P for d = 2 is S, SUM adds the
    control into the target, CZ is diag(w^(ij)), SWAP is what the reference's prime-d SWAP composite does to states."""
    from oracle import statevector_oracle as sv
    for d in (2, 3, 5):
        w = np.exp(2j * np.pi / d)
        x, z, h = sv.x_matrix(d), sv.z_matrix(d), sv.h_matrix(d)
        e0 = np.eye(d)[0]
        assert np.allclose(x @ e0, np.eye(d)[1])
        assert np.allclose(np.diag(z), w ** np.arange(d))
        assert np.allclose(h @ x @ h.conj().T, z)
        cz = sv.cz_matrix(d)
        assert np.allclose(cz, np.diag([w ** (i * j) for i in range(d) for j in range(d)]))
        cn = sv.cnot_matrix(d)
        for i in range(d):
            for j in range(d):
                assert cn[d * i + (i + j) % d, d * i + j] == 1
    assert np.allclose(sv.p_matrix(2), np.diag([1, 1j]))
    assert np.allclose(np.diag(sv.p_matrix(3)), [1, 1, np.exp(2j * np.pi / 3)])


@pytest.mark.gpu
@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("depth", [5, 30, 100, 1000])
def test_random_circuits_tomography_on_gpu(d, depth):
    """tests/test_tomography.py:13-77 through the public API: 800 shots of Program.simulate on a random 3-qudit
    Clifford circuit, total variation distance to |amplitude|^2 below 20 % (the reference's bound), plus the exact
    check: the support of the sampled distribution lies inside the support of the state vector."""
    from oracle.statevector_oracle import outcome_distribution
    from sdim_b200 import Program, generate_random_clifford_circuit
    n, shots = 3, 800
    for i in range(6):
        circuit = generate_random_clifford_circuit(n, depth, d, measurement_rounds=1, seed=100 * depth + i)
        want = outcome_distribution(n, d, compile_circuits([circuit]).ops)
        results = Program(circuit).simulate(shots=shots, force_tableau=True)
        counts = np.zeros([d] * n)
        for s in range(shots):
            counts[tuple(results[q][0][s].measurement_value for q in range(n))] += 1
        probs = counts / shots
        assert np.isclose(want.sum(), 1.0, atol=1e-6)
        assert np.abs(probs - want).sum() / 2 < 0.20
        assert not np.any((want < 1e-12) & (counts > 0))
