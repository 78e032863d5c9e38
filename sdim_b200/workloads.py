"""Synthetic workloads of BASELINE.json's configs (constructions: SURVEY section 8d).

Host-side circuit builders only; they use the public Circuit API so the same
construction can be fed to the reference.
"""
from __future__ import annotations

from .circuit import Circuit
from .random_circuit import generate_random_clifford_circuit


def noisy_random_clifford(num_qudits: int, num_gates: int, dimension: int, seed: int = 1,
                          prob: float = 1e-3, channel: str = "d", measurement_rounds: int = 1) -> Circuit:
    """Headline workload: `generate_random_clifford_circuit(n, gates, d, seed)` with an
    `N1 prob noise_channel` inserted after every gate on the qudit(s) it touched, then
    `measurement_rounds` x M on every qudit."""
    base = generate_random_clifford_circuit(num_qudits, num_gates, dimension, measurement_rounds=0, seed=seed)
    c = Circuit(num_qudits, dimension)
    for op in base.operations:
        if op.target_index is None:
            c.add_gate(op.gate_name, op.qudit_index)
            c.add_gate("N1", op.qudit_index, prob=prob, noise_channel=channel)
        else:
            c.add_gate(op.gate_name, op.qudit_index, op.target_index)
            c.add_gate("N1", [op.qudit_index, op.target_index], prob=prob, noise_channel=channel)
    for _ in range(measurement_rounds):
        c.add_gate("M", list(range(num_qudits)))
    return c


def qudit_repetition_code(distance: int = 25, rounds: int = 25, dimension: int = 3,
                          prob: float = 1e-2, channel: str = "f") -> Circuit:
    """Config 4: qutrit repetition-code memory experiment generalising examples/repetition_code.ipynb.

    Data qudits 0..distance-1, ancillas distance..2*distance-2.  Per round and ancilla i:
    CNOT(i, anc), CNOT(i+1, anc) x (d-1) (accumulates z_i - z_{i+1}), N1 on each data qudit,
    RESET (measure + reset) on each ancilla; final M on all data qudits.
    """
    n = 2 * distance - 1
    c = Circuit(n, dimension)
    for _ in range(rounds):
        for i in range(distance - 1):
            anc = distance + i
            c.add_gate("CNOT", i, anc)
            for _ in range(dimension - 1):
                c.add_gate("CNOT", i + 1, anc)
        for q in range(distance):
            c.add_gate("N1", q, prob=prob, noise_channel=channel)
        for i in range(distance - 1):
            c.add_gate("RESET", distance + i)
    c.add_gate("M", list(range(distance)))
    return c


def rotated_surface_code(distance: int = 7, rounds: int = 7, prob: float = 1e-3) -> Circuit:
    """Config 3: d = 2 rotated surface code memory experiment (examples/surface_code.ipynb is empty,
    so the circuit is synthesised): distance^2 data qubits, distance^2 - 1 ancillas; per round H on
    X-ancillas, four CNOT layers, H, single-qubit depolarising N1 on every data qubit, RESET on every
    ancilla; final M on the data qubits."""
    dd = distance
    data = {(r, c): r * dd + c for r in range(dd) for c in range(dd)}
    ancillas = []          # (index, kind, [data neighbours in NW, NE, SW, SE order or None])
    nxt = dd * dd
    for r in range(dd + 1):
        for c in range(dd + 1):
            kind = "X" if (r + c) % 2 == 0 else "Z"
            nbrs = [data.get((r - 1, c - 1)), data.get((r - 1, c)), data.get((r, c - 1)), data.get((r, c))]
            present = [q for q in nbrs if q is not None]
            if len(present) == 4:
                pass
            elif len(present) == 2:
                # boundary stabilisers: X-type on top/bottom edges, Z-type on left/right edges
                on_top_bottom = r in (0, dd)
                if (kind == "X") != on_top_bottom:
                    continue
            else:
                continue
            ancillas.append((nxt, kind, nbrs))
            nxt += 1
    circuit = Circuit(nxt, 2)
    x_anc = [a for a, k, _ in ancillas if k == "X"]
    for _ in range(rounds):
        if x_anc:
            circuit.add_gate("H", x_anc)
        for layer in range(4):
            for a, kind, nbrs in ancillas:
                # Z-type ancillas visit neighbours in N-order, X-type in Z-order (standard hook-error-safe schedule)
                order = (0, 1, 2, 3) if kind == "X" else (0, 2, 1, 3)
                q = nbrs[order[layer]]
                if q is None:
                    continue
                if kind == "X":
                    circuit.add_gate("CNOT", a, q)
                else:
                    circuit.add_gate("CNOT", q, a)
        if x_anc:
            circuit.add_gate("H", x_anc)
        circuit.add_gate("N1", list(range(dd * dd)), prob=prob, noise_channel="d")
        circuit.add_gate("RESET", [a for a, _, _ in ancillas])
    circuit.add_gate("M", list(range(dd * dd)))
    return circuit
