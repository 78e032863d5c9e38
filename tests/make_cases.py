"""Seeded random programs over every opcode, built through the public Circuit API (test infrastructure)."""
import random

from sdim_b200.circuit import Circuit
from sdim_b200.ir import compile_circuits

ONE = ["I", "X", "X_INV", "Z", "Z_INV", "H", "H_INV", "P", "P_INV"]
TWO = ["CNOT", "CNOT_INV", "CZ", "CZ_INV", "SWAP"]


def random_circuit(seed, n, d, depth, p_meas=0.08, p_noise=0.1, final_measure=True, noise_prob=None, p_burst=0.0):
    rng = random.Random(seed)
    c = Circuit(n, d)
    for _ in range(depth):
        u = rng.random()
        if p_burst and rng.random() < p_burst:        # a run of plain M ops on a random subset, mid-circuit
            c.add_gate("M", rng.sample(range(n), rng.randint(2, n)))
        elif u < p_meas:
            c.add_gate(rng.choice(["M", "M", "M_X", "RESET"]), rng.randrange(n))
        elif u < p_meas + p_noise:
            prob = noise_prob if noise_prob is not None else rng.choice([0.05, 0.3, 0.9, 1.0])
            c.add_gate("N1", rng.randrange(n), prob=prob, noise_channel=rng.choice(["d", "f", "p"]))
        elif n >= 2 and rng.random() < 0.4:
            a, b = rng.sample(range(n), 2)
            c.add_gate(rng.choice(TWO), a, b)
        else:
            c.add_gate(rng.choice(ONE), rng.randrange(n))
    if final_measure:
        c.add_gate("M", list(range(n)))
    return c


def random_program(seed, n, d, depth, **kw):
    return compile_circuits([random_circuit(seed, n, d, depth, **kw)])
