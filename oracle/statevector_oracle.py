"""STATE-VECTOR ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

A third, independent line of validation for small systems (n <= ~6 qudits): the circuit is applied to a dense state
vector with the reference's own gate MATRICES, restated from sdim/unitary.py:26-125 (generalised X, Z, Hadamard,
phase, SUM) and the CZ composition of sdim/unitary.py:354-376, the way the reference's tests/test_tomography.py does
through Cirq (sdim/circuit_io.py:140-222; Cirq is not installed here, and nothing of it is needed to multiply
matrices).  It knows nothing about tableaus: it shares no arithmetic with oracle/tableau_oracle.py, oracle/oracle.c
or the CUDA kernels, so agreement of outcome DISTRIBUTIONS is evidence about the stabilizer update rules themselves
(phase conventions of P for even / odd d, the sign of H vs H^-1, the direction of SUM).

Supported: every unitary opcode of the op stream (I .. SWAP) followed by terminal Z-basis measurements.
"""
from __future__ import annotations

import numpy as np

OP_NAMES = {0: "I", 1: "X", 2: "X_INV", 3: "Z", 4: "Z_INV", 5: "H", 6: "H_INV", 7: "P", 8: "P_INV",
            9: "CNOT", 10: "CNOT_INV", 11: "CZ", 12: "CZ_INV", 13: "SWAP"}


def x_matrix(d):
    """sdim/unitary.py:26-37: X|j> = |j+1>."""
    m = np.zeros((d, d), dtype=np.complex128)
    for i in range(d):
        m[i, (i - 1) % d] = 1
    return m


def z_matrix(d):
    """sdim/unitary.py:39-50: Z|j> = w^j |j>."""
    return np.diag(np.exp(2j * np.pi * np.arange(d) / d))


def h_matrix(d):
    """sdim/unitary.py:52-70 (prime d): H[m, n] = w^(mn) / sqrt(d)."""
    idx = np.arange(d)
    return np.exp(2j * np.pi * np.outer(idx, idx) / d) / np.sqrt(d)


def p_matrix(d):
    """sdim/unitary.py:88-110 (prime d): diag w^(j(j-1)/2) for odd d, w^(j^2/2) for d = 2."""
    j = np.arange(d)
    expo = j * (j - 1) / 2 if d % 2 == 1 else j ** 2 / 2
    return np.diag(np.exp(2j * np.pi * expo / d))


def cnot_matrix(d):
    """sdim/unitary.py:112-125: |i, j> -> |i, i+j>."""
    m = np.zeros((d * d, d * d), dtype=np.complex128)
    for i in range(d):
        for j in range(d):
            m[d * i + (i + j) % d, d * i + j] = 1
    return m


def cz_matrix(d):
    """sdim/unitary.py:354-365: (I (x) H) CNOT (I (x) H^dagger)."""
    h = h_matrix(d)
    return np.kron(np.eye(d), h) @ cnot_matrix(d) @ np.kron(np.eye(d), h.conj().T)


def swap_matrix(d):
    m = np.zeros((d * d, d * d), dtype=np.complex128)
    for i in range(d):
        for j in range(d):
            m[d * j + i, d * i + j] = 1
    return m


def _gate(name, d):
    base = {"X": x_matrix, "Z": z_matrix, "H": h_matrix, "P": p_matrix, "CNOT": cnot_matrix, "CZ": cz_matrix,
            "SWAP": swap_matrix}
    if name.endswith("_INV"):
        return base[name[:-4]](d).conj().T
    return base[name](d)


def final_state(n, d, ops):
    """State vector (shape [d]*n, qudit 0 = most significant axis, as Cirq's LineQid order) after the unitary ops."""
    psi = np.zeros([d] * n, dtype=np.complex128)
    psi[(0,) * n] = 1.0
    for op, a, b, _ in ops:
        op = int(op) & 0xFF
        if op == 0 or op >= 14:
            if op in (15, 16, 17):
                raise NotImplementedError("state-vector oracle: unitary ops + terminal M only")
            continue
        g = _gate(OP_NAMES[op], d)
        if op < 9:
            psi = np.moveaxis(np.tensordot(g, psi, axes=([1], [a])), 0, a)
        else:
            g4 = g.reshape(d, d, d, d)                      # [a', b', a, b]
            psi = np.moveaxis(np.tensordot(g4, psi, axes=([2, 3], [a, b])), [0, 1], [a, b])
    return psi


def outcome_distribution(n, d, ops):
    """P(m_0, ..., m_{n-1}) of measuring every qudit in the Z basis: float64 array of shape [d]*n."""
    return np.abs(final_state(n, d, ops)) ** 2
