"""Random Clifford circuit generator (reference: sdim/random_circuit.py:9-98).

Workload generator for the benchmark configs.  It draws from the stdlib
`random` stream in the same order as the reference — one `choice` over the
gate list per gate, then `sample(range(n), 2)` for a two-qudit gate or
`randint(0, n-1)` for a single-qudit gate — so a given seed yields the same
circuit in both packages.
"""
from __future__ import annotations

import random
from typing import Optional, Sequence

from .circuit import Circuit
from .circuit_io import write_circuit

DEFAULT_GATE_SET = ("H", "P", "CNOT", "X", "Z", "H_INV", "P_INV", "CNOT_INV",
                    "X_INV", "Z_INV", "CZ", "CZ_INV")
_PAIR_GATES = frozenset(("CNOT", "CNOT_INV", "CZ", "CZ_INV"))


def generate_random_clifford_circuit(num_qudits: int, num_gates: int, dimension: int,
                                     measurement_rounds: int = 0, seed: Optional[int] = None,
                                     gate_set: Optional[Sequence[str]] = None) -> Circuit:
    """`num_gates` uniformly chosen Clifford gates, then `measurement_rounds` x `M` on every qudit."""
    if seed is not None:
        random.seed(seed)
    names = list(gate_set) if gate_set else list(DEFAULT_GATE_SET)
    circuit = Circuit(num_qudits, dimension)
    for _ in range(num_gates):
        name = random.choice(names)
        if name in _PAIR_GATES:
            a, b = random.sample(range(num_qudits), 2)
            circuit.add_gate(name, a, b)
        else:
            circuit.add_gate(name, random.randint(0, num_qudits - 1))
    for _ in range(measurement_rounds):
        for q in range(num_qudits):
            circuit.add_gate("M", q)
    return circuit


def generate_and_write_random_circuit(num_qudits: int, num_gates: int, dimension: int,
                                      measurement_rounds: int = 0,
                                      output_file: str = "random_circuit.chp",
                                      seed: Optional[int] = None) -> Circuit:
    circuit = generate_random_clifford_circuit(num_qudits, num_gates, dimension,
                                               measurement_rounds, seed)
    write_circuit(circuit, output_file)
    return circuit
