"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

numpy restatement of the reference's prime-dimension stabilizer-tableau path
(events555/sdim).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module;
the product (`sdim_b200/`) never does and fails loudly without its CUDA library.

Parity status: PINNED.  `tests/golden/*.json` were produced by running the
unmodified reference (imported from /root/reference with `cirq` / `diophantine`
stubbed, see oracle/ref_harness.py and oracle/make_golden.py) and
tests/test_oracle_golden.py checks this file against every one of them —
measurement records and all six final arrays.

State uses the reference's own orientation (sdim/tableau/dataclasses.py:14,24-39,
sdim/tableau/tableau_prime.py:24-26,76-86): `x[q, g]` is the X exponent of
generator (column) g on qudit (row) q.  Unlike the reference, which lets values
drift and reduces every 64 gates (sdim/program.py:317-318), entries are kept
reduced mod d / mod order at all times; after the reference's `modulo()` the two
agree.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

OP_I, OP_X, OP_X_INV, OP_Z, OP_Z_INV = 0, 1, 2, 3, 4
OP_H, OP_H_INV, OP_P, OP_P_INV = 5, 6, 7, 8
OP_CNOT, OP_CNOT_INV, OP_CZ, OP_CZ_INV, OP_SWAP = 9, 10, 11, 12, 13
OP_M, OP_M_X, OP_RESET, OP_N1 = 14, 15, 16, 17


class OracleTableau:
    """Six int64 arrays of one shot (sdim/tableau/tableau_prime.py:8-95)."""

    def __init__(self, n: int, d: int):
        self.n, self.d = n, d
        # order / phase_order: sdim/tableau/dataclasses.py:88-106
        self.po = 2 if d % 2 == 0 else 1
        self.order = d * self.po
        # |0...0>: stabilizers Z_q, destabilizers X_q (dataclasses.py:34-39, tableau_prime.py:81-86)
        self.x = np.zeros((n, n), dtype=np.int64)
        self.z = np.eye(n, dtype=np.int64)
        self.p = np.zeros(n, dtype=np.int64)
        self.dx = np.eye(n, dtype=np.int64)
        self.dz = np.zeros((n, n), dtype=np.int64)
        self.dp = np.zeros(n, dtype=np.int64)

    def arrays(self):
        return self.x, self.z, self.p, self.dx, self.dz, self.dp

    # ---- unitary gates -------------------------------------------------------------------
    def _halves(self):
        return ((self.x, self.z, self.p), (self.dx, self.dz, self.dp))

    def hadamard(self, a: int, inverse: bool = False):
        """tableau_optimized.py:5-58: H (x,z)<-(-z,x); H^-1 (x,z)<-(z,-x); phase += po*new_x*new_z."""
        d, po, o = self.d, self.po, self.order
        for X, Z, P in self._halves():
            xa, za = X[a].copy(), Z[a].copy()
            P -= po * xa * za
            P %= o
            if inverse:
                X[a], Z[a] = za, (-xa) % d
            else:
                X[a], Z[a] = (-za) % d, xa

    def phase(self, a: int, inverse: bool = False):
        """tableau_optimized.py:62-96: even d phase +-= x^2, odd d phase +-= x(x-1)//2; z +-= x."""
        d, o = self.d, self.order
        s = -1 if inverse else 1
        for X, Z, P in self._halves():
            xa = X[a]
            inc = xa * xa if d % 2 == 0 else (xa * (xa - 1)) // 2
            P += s * inc
            P %= o
            Z[a] = (Z[a] + s * xa) % d

    def cnot(self, a: int, b: int, inverse: bool = False):
        """tableau_optimized.py:99-118: x[t] +-= x[c]; z[c] +-= (d-1) z[t]; no phase term."""
        d = self.d
        s = -1 if inverse else 1
        for X, Z, _ in self._halves():
            X[b] = (X[b] + s * X[a]) % d
            Z[a] = (Z[a] - s * Z[b]) % d

    def pauli(self, q: int, a: int, b: int):
        """Conjugation by X^a Z^b on qudit q: phase += po*(b*x - a*z).

        Net effect of the reference's composite Paulis (tableau_gates.py:27-137:
        X = H P^-1 H H P H for odd d, H P P H for d = 2, ...), SURVEY Appendix A-2.
        """
        po, o = self.po, self.order
        for X, Z, P in self._halves():
            P += po * (b * X[q] - a * Z[q])
            P %= o

    def cz(self, a: int, b: int, inverse: bool = False):
        """tableau_gates.py:229-261: CZ = H^-1(t) CNOT(c,t) H(t), folded (Appendix A-2)."""
        d, po, o = self.d, self.po, self.order
        s = -1 if inverse else 1
        for X, Z, P in self._halves():
            xa, xb = X[a].copy(), X[b].copy()
            P += s * po * xa * xb
            P %= o
            Z[a] = (Z[a] + s * xb) % d
            Z[b] = (Z[b] + s * xa) % d

    def swap(self, a: int, b: int):
        """tableau_gates.py:298-329 (prime branch): nine primitives whose net effect is a row swap."""
        for X, Z, _ in self._halves():
            X[[a, b]] = X[[b, a]]
            Z[[a, b]] = Z[[b, a]]

    # ---- measurement -----------------------------------------------------------------------
    def measure(self, q: int, draw: Callable[[], int]) -> Tuple[bool, int]:
        """tableau_prime.py:262-363.  Returns (deterministic, value); `draw()` supplies the
        outcome of a random measurement (the reference calls random.choice(range(d)), :332)."""
        n, d, po, o = self.n, self.d, self.po, self.order
        x, z, p, dx, dz, dp = self.arrays()
        nz = np.nonzero(x[q])[0]
        if nz.size == 0:
            return True, self._det_measure(q)
        piv = int(nz[0])                                   # first anticommuting stabilizer (:273-283)
        v = int(x[q, piv])
        if v != 1:                                         # exponentiate (:365-380)
            e = pow(v, -1, d)
            p[piv] = (p[piv] * e + int(x[:, piv] @ z[:, piv]) * (e * (e - 1) // 2) * po) % o
            x[:, piv] = (x[:, piv] * e) % d
            z[:, piv] = (z[:, piv] * e) % d
        xs, zs, ps = x[:, piv].copy(), z[:, piv].copy(), int(p[piv])
        sd = int(xs @ zs)
        # _random_measurement (:294-334): every generator with an X component on q absorbs
        # f copies of the pivot; iterations only read the pivot column, so they are independent.
        for X, Z, P, skip in ((dx, dz, dp, -1), (x, z, p, piv)):
            f = (-X[q]) % d
            if skip >= 0:
                f[skip] = 0
            cp = (xs @ Z) * f + sd * ((f * (f - 1)) // 2) * po
            X += np.outer(xs, f)
            X %= d
            Z += np.outer(zs, f)
            Z %= d
            P += f * ps + po * cp
            P %= o
        dx[:, piv], dz[:, piv], dp[piv] = xs, zs, ps       # destabilizer <- old pivot (:323-326)
        x[:, piv] = 0
        z[:, piv] = 0
        z[q, piv] = 1                                      # stabilizer <- Z_q (:328-330)
        m = int(draw())
        p[piv] = (-m * po) % o                             # (:331-333)
        return False, m

    def _det_measure(self, q: int) -> int:
        """_det_measurement (:336-363): ordered accumulation of stabilizers picked by destab X[q]."""
        n, d, po, o = self.n, self.d, self.po, self.order
        ax = np.zeros(n, dtype=np.int64)
        az = np.zeros(n, dtype=np.int64)
        ap = 0
        for i in range(n):
            f = int(self.dx[q, i])
            if f == 0:
                continue
            xi, zi = self.x[:, i], self.z[:, i]
            cp = int(az @ (f * xi)) + int(xi @ zi) * (f * (f - 1) // 2) * po
            ax = (ax + f * xi) % d
            az = (az + f * zi) % d
            ap = (ap + f * int(self.p[i]) + po * cp) % o
        return ((-ap) // po) % d                           # precedence as written at :362


def run_shot(n: int, d: int, ops: Sequence[Sequence[int]],
             meas_draw: Optional[Callable[[int], int]] = None,
             noise_ab: Optional[np.ndarray] = None,
             tableau: Optional[OracleTableau] = None):
    """One pass of Program._simulate_tableau's inner loop (sdim/program.py:311-351).

    ops: rows (opcode, a, b, slot); slot = chronological measurement index for
    M / M_X / RESET and noise-event index for N1.
    meas_draw(k) -> outcome used if measurement k turns out random.
    noise_ab[j] = (a, b): the Pauli X^a Z^b that N1 event j applies in this shot
    (the reference tableau path ignores N1, program.py:31; parity for noise goes
    through explicit Pauli substitution).
    Returns (records, tableau) with records = [(qudit, deterministic, value)] in
    chronological order.
    """
    t = tableau if tableau is not None else OracleTableau(n, d)
    records: List[Tuple[int, bool, int]] = []
    for op, a, b, slot in ops:
        op, a, b, slot = int(op), int(a), int(b), int(slot)
        if op == OP_I:
            pass
        elif op == OP_X:
            t.pauli(a, 1, 0)
        elif op == OP_X_INV:
            t.pauli(a, d - 1, 0)
        elif op == OP_Z:
            t.pauli(a, 0, 1)
        elif op == OP_Z_INV:
            t.pauli(a, 0, d - 1)
        elif op == OP_H:
            t.hadamard(a)
        elif op == OP_H_INV:
            t.hadamard(a, inverse=True)
        elif op == OP_P:
            t.phase(a)
        elif op == OP_P_INV:
            t.phase(a, inverse=True)
        elif op == OP_CNOT:
            t.cnot(a, b)
        elif op == OP_CNOT_INV:
            t.cnot(a, b, inverse=True)
        elif op == OP_CZ:
            t.cz(a, b)
        elif op == OP_CZ_INV:
            t.cz(a, b, inverse=True)
        elif op == OP_SWAP:
            t.swap(a, b)
        elif op in (OP_M, OP_M_X, OP_RESET):
            if op == OP_M_X:                                # tableau_gates.py:292-296, no rotation back
                t.hadamard(a, inverse=True)
            det, m = t.measure(a, (lambda k=slot: meas_draw(k)) if meas_draw else (lambda: 0))
            records.append((a, det, m))
            if op == OP_RESET:                              # program.py:335-339: X applied (-m) mod d times
                t.pauli(a, (-m) % d, 0)
        elif op == OP_N1:
            if noise_ab is not None:
                na, nb = int(noise_ab[slot][0]), int(noise_ab[slot][1])
                if na or nb:
                    t.pauli(a, na, nb)
        else:
            raise ValueError("Invalid gate value")          # program.py:381-382
    return records, t


def run_shots(n: int, d: int, ops, shots: int, meas_draws: Optional[np.ndarray] = None,
              noise_ab: Optional[np.ndarray] = None, keep_tableau: bool = False):
    """`shots` independent passes.  meas_draws[shot, k], noise_ab[shot, j, 2] are replayed draws.

    Returns records uint8[shots, n_meas] with bit 7 = deterministic flag and the
    low bits = value (the packed format of the CUDA path), plus the last tableau.
    """
    ops = np.asarray(ops, dtype=np.int64).reshape(-1, 4)
    n_meas = int(np.isin(ops[:, 0], (OP_M, OP_M_X, OP_RESET)).sum())
    wide = d > 127                       # uint16 records, bit 15 = deterministic (include/sdimb.h)
    out = np.zeros((shots, n_meas), dtype=np.uint16 if wide else np.uint8)
    last = None
    for s in range(shots):
        md = (lambda k, s=s: int(meas_draws[s, k])) if meas_draws is not None else None
        na = noise_ab[s] if noise_ab is not None else None
        recs, last = run_shot(n, d, ops, md, na)
        for k, (_, det, m) in enumerate(recs):
            out[s, k] = ((m & 0x7FFF) | (0x8000 if det else 0)) if wide else ((m & 0x7F) | (0x80 if det else 0))
    return out, (last if keep_tableau else None)
