"""Peephole folding of the op stream (the reference runs every gate as written,
sdim/program.py:311-312).

Exact on the tableau, phases included — records under the same draws and final tableaus are unchanged
(`tests/test_peephole.py` checks both against the oracle):

* a run of one single-qudit family on one qudit, with nothing else touching that qudit in between, is reduced
  modulo the family's order: X, Z have order d; H has order 4 (2 for d = 2: `hadamard_optimized` applied four
  times is the identity with zero net phase, sdim/tableau/tableau_optimized.py:5-58); P has order d (4 for d = 2,
  `phase_optimized`, :62-96).  The exponent left is emitted as the shorter of G^e and (G^-1)^(order - e);
* a two-qudit gate directly followed (on both of its qudits) by its inverse on the same pair cancels:
  CNOT / CNOT_INV with the same control and target, CZ / CZ_INV and SWAP / SWAP in either order; for d = 2 a gate
  is its own inverse (`cnot_optimized`, :99-118).

Measurements, RESET and N1 are never merged and keep their event slots, so Philox draws and replay arrays address
the same events; a measurement or RESET is a barrier for every qudit (like in `sdimb_schedule`), an N1 for its own.
Gates on disjoint qudits commute exactly (the argument `sdimb_schedule` relies on), which is what makes "directly
followed on that qudit" the right notion of adjacency.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import numpy as np

from .gatedata import (OP_CNOT, OP_CNOT_INV, OP_CZ, OP_CZ_INV, OP_H, OP_H_INV, OP_I, OP_N1, OP_P, OP_P_INV, OP_SWAP,
                       OP_X, OP_X_INV, OP_Z, OP_Z_INV, TWO_QUDIT_OPS)

# single-qudit families: opcode -> (forward opcode, inverse opcode, sign)
_FAMILY = {OP_X: (OP_X, OP_X_INV, 1), OP_X_INV: (OP_X, OP_X_INV, -1),
           OP_Z: (OP_Z, OP_Z_INV, 1), OP_Z_INV: (OP_Z, OP_Z_INV, -1),
           OP_H: (OP_H, OP_H_INV, 1), OP_H_INV: (OP_H, OP_H_INV, -1),
           OP_P: (OP_P, OP_P_INV, 1), OP_P_INV: (OP_P, OP_P_INV, -1)}
_TWO_INVERSE = {OP_CNOT: OP_CNOT_INV, OP_CNOT_INV: OP_CNOT, OP_CZ: OP_CZ_INV, OP_CZ_INV: OP_CZ, OP_SWAP: OP_SWAP}
_TWO_FAMILY = {OP_CNOT: 0, OP_CNOT_INV: 0, OP_CZ: 1, OP_CZ_INV: 1, OP_SWAP: 2}


def family_order(forward_op: int, d: int) -> int:
    if forward_op in (OP_X, OP_Z):
        return d
    if forward_op == OP_H:
        return 2 if d == 2 else 4
    return 4 if d == 2 else d            # P


@dataclasses.dataclass
class _Node:
    kind: str                  # "run" | "two" | "fixed"
    op: int                    # run: forward opcode; two / fixed: opcode
    a: int
    b: int = -1
    slot: int = -1
    exponent: int = 0          # run only, in [1, order)
    alive: bool = True


def fold_ops(ops: np.ndarray, d: int) -> np.ndarray:
    """int32[n_ops, 4] -> folded int32[m, 4], m <= n_ops."""
    nodes: List[_Node] = []
    stacks: Dict[int, List[int]] = {}

    def top(q: int) -> Optional[int]:
        st = stacks.get(q)
        return st[-1] if st else None

    def push(q: int, idx: int) -> None:
        stacks.setdefault(q, []).append(idx)

    for op, a, b, slot in np.asarray(ops, dtype=np.int64).reshape(-1, 4).tolist():
        if op == OP_I:
            continue
        fam = _FAMILY.get(op)
        if fam is not None:
            forward, _inverse, sign = fam
            order = family_order(forward, d)
            t = top(a)
            if t is not None and nodes[t].kind == "run" and nodes[t].op == forward:
                e = (nodes[t].exponent + sign) % order
                if e == 0:
                    nodes[t].alive = False
                    stacks[a].pop()
                else:
                    nodes[t].exponent = e
            elif order > 1:
                nodes.append(_Node("run", forward, a, exponent=sign % order))
                push(a, len(nodes) - 1)
            continue
        if op in TWO_QUDIT_OPS:
            ta, tb = top(a), top(b)
            if ta is not None and ta == tb and nodes[ta].kind == "two":
                prev = nodes[ta]
                same_pair = (prev.a, prev.b) == (a, b)
                either_order = same_pair or (prev.a, prev.b) == (b, a)
                cancels = op == _TWO_INVERSE[prev.op] or (d == 2 and _TWO_FAMILY[op] == _TWO_FAMILY[prev.op])
                pair_ok = same_pair if _TWO_FAMILY[op] == 0 else either_order      # CNOT is directional
                if cancels and pair_ok:
                    prev.alive = False
                    stacks[a].pop()
                    stacks[b].pop()
                    continue
            nodes.append(_Node("two", op, a, b))
            push(a, len(nodes) - 1)
            push(b, len(nodes) - 1)
            continue
        nodes.append(_Node("fixed", op, a, -1, slot))       # M, M_X, RESET, N1: never merged, slot kept
        if op == OP_N1:
            push(a, len(nodes) - 1)                         # a Pauli on its own qudit
        else:
            stacks.clear()                                  # a measurement rewrites every generator: full barrier,
                                                            # as in sdimb_schedule

    rows = []
    for nd in nodes:
        if not nd.alive:
            continue
        if nd.kind == "run":
            forward, inverse, _ = _FAMILY[nd.op]
            order = family_order(forward, d)
            e = nd.exponent
            if e <= order - e:
                rows.extend([(forward, nd.a, -1, -1)] * e)
            else:
                rows.extend([(inverse, nd.a, -1, -1)] * (order - e))
        else:
            rows.append((nd.op, nd.a, nd.b, nd.slot))
    return np.array(rows, dtype=np.int32).reshape(-1, 4)


def fold_program(prog):
    """CompiledProgram with its op stream folded; everything else (slots, noise tables, the user gate count the
    shot*gates metric uses) is unchanged."""
    return dataclasses.replace(prog, ops=fold_ops(prog.ops, prog.dimension))
