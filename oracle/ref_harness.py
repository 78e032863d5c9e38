"""TEST INFRASTRUCTURE ONLY: drives the UNMODIFIED reference from /root/reference.

Used by oracle/make_golden.py (to produce tests/golden/*.json) and by the
opt-in tests that cross-check the oracle live when /root/reference is mounted.
Nothing here runs on the GPU box and nothing in the product imports it.

Import recipe: `cirq` and `diophantine` are not installed and are
not touched by the prime-dimension path, so empty stand-in modules satisfy the
reference's top-level imports.
"""
from __future__ import annotations

import os
import random as _stdlib_random
import sys
import types
from typing import List, Optional, Sequence, Tuple

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference() -> str:
    """$SDIM_REFERENCE_ROOT, else the read-only checkout of this container, else the offline install that travels to
    the GPU box (baseline/_ref, made by baseline/install_reference.sh; git-ignored)."""
    cands = [os.environ.get("SDIM_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "sdim")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference()

_NAMES = ["I", "X", "X_INV", "Z", "Z_INV", "H", "H_INV", "P", "P_INV", "CNOT", "CNOT_INV",
          "CZ", "CZ_INV", "SWAP", "M", "M_X", "RESET", "N1"]
_TWO = {9, 10, 11, 12, 13}
_MEAS = {14, 15, 16}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sdim"))


def load_reference():
    """Import the reference package `sdim` with the two absent third-party modules stubbed."""
    if "sdim" in sys.modules and getattr(sys.modules["sdim"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["sdim"]
    if "cirq" not in sys.modules:
        cirq = types.ModuleType("cirq")
        cirq.Gate = type("Gate", (), {"__init__": lambda self, *a, **k: None})
        sys.modules["cirq"] = cirq
    if "diophantine" not in sys.modules:
        sys.modules["diophantine"] = types.ModuleType("diophantine")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sdim  # noqa: E402  (the reference)
    return sdim


class _ChoiceFeed:
    """Stands in for the `random` module inside sdim.tableau.tableau_prime (its only RNG use is
    random.choice(range(d)) at tableau_prime.py:332)."""

    def __init__(self, rng: _stdlib_random.Random):
        self._rng = rng

    def choice(self, seq):
        return self._rng.choice(list(seq))


def build_reference_circuit(sdim, n: int, d: int, ops: Sequence[Sequence[int]],
                            noise_ab: Optional[np.ndarray] = None):
    """ops rows (opcode, a, b, slot) -> reference Circuit.  N1 events are replaced by the explicit
    Pauli X^a Z^b they stand for in this shot (the reference tableau path ignores N1)."""
    c = sdim.Circuit(n, d)
    for op, a, b, slot in ops:
        op, a, b, slot = int(op), int(a), int(b), int(slot)
        if op == 17:
            if noise_ab is not None:
                for _ in range(int(noise_ab[slot][0])):
                    c.add_gate("X", a)
                for _ in range(int(noise_ab[slot][1])):
                    c.add_gate("Z", a)
            continue
        if op in _TWO:
            c.add_gate(_NAMES[op], a, b)
        else:
            c.add_gate(_NAMES[op], a)
    return c


def _chronological(ops, results_by_qudit) -> List[Tuple[int, bool, int]]:
    count = {}
    out = []
    for op, a, _b, _slot in ops:
        if int(op) in _MEAS:
            a = int(a)
            r = count.get(a, 0)
            count[a] = r + 1
            res = results_by_qudit[a][r][0]
            out.append((a, bool(res.deterministic), int(res.measurement_value)))
    return out


def _final_arrays(t):
    return {k: np.array(getattr(t, name)).astype(np.int64) for k, name in (
        ("x", "x_block"), ("z", "z_block"), ("p", "phase_vector"),
        ("dx", "destab_x_block"), ("dz", "destab_z_block"), ("dp", "destab_phase_vector"))}


def ref_run(n: int, d: int, ops, noise_ab=None, draw_seed: int = 0):
    """Run one shot through the reference's own Program.simulate.

    Returns (records, arrays): chronological [(qudit, deterministic, value)] and the six final
    arrays after the reference's closing modulo() (program.py:351).
    """
    sdim = load_reference()
    import sdim.tableau.tableau_prime as tp
    circ = build_reference_circuit(sdim, n, d, ops, noise_ab)
    prog = sdim.Program(circ)
    saved = tp.random
    tp.random = _ChoiceFeed(_stdlib_random.Random(draw_seed))
    try:
        prog.simulate(shots=1)
    finally:
        tp.random = saved
    return _chronological(ops, prog.measurement_results), _final_arrays(prog.stabilizer_tableau)


def ref_run_eager_modulo(n: int, d: int, ops, noise_ab=None, draw_seed: int = 0):
    """Same shot, but stepping the reference's gate functions one at a time with modulo() before
    every gate.  If this disagrees with ref_run the reference overflowed int64 inside its
    64-gate lazy-reduction window and the case must be dropped."""
    sdim = load_reference()
    import sdim.tableau.tableau_prime as tp
    from sdim.tableau.tableau_gates import apply_X
    circ = build_reference_circuit(sdim, n, d, ops, noise_ab)
    prog = sdim.Program(circ)
    saved = tp.random
    tp.random = _ChoiceFeed(_stdlib_random.Random(draw_seed))
    recs = []
    try:
        t = prog.stabilizer_tableau
        for gate in circ.operations:
            t.modulo()
            res = prog.apply_gate(gate)
            if res is not None:
                recs.append((int(res.qudit_index), bool(res.deterministic), int(res.measurement_value)))
                if gate.gate_id == 16:
                    for _ in range((-int(res.measurement_value)) % d):
                        t.modulo()
                        apply_X(t, gate.qudit_index, None)
        t.modulo()
    finally:
        tp.random = saved
    return recs, _final_arrays(t)
