"""Gate table for the prime-dimension tableau path.

Mirrors the id assignment and alias set of the reference gate registry
(reference: sdim/gatedata.py:42-116): ids 0..17 in registration order and the
same alias spellings, so `.chp` files and `Circuit.add_gate` calls written for
the reference resolve to the same integer ids here.  The table is built once
(the reference rebuilds it per Circuit and scans it linearly per instruction,
gatedata.py:112-116); lookups here are dict hits.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

# (canonical name, arity, aliases)
_GATE_TABLE = (
    ("I", 1, ()),
    ("X", 1, ()),
    ("X_INV", 1, ()),
    ("Z", 1, ()),
    ("Z_INV", 1, ()),
    ("H", 1, ("R", "DFT")),
    ("H_INV", 1, ("R_INV", "DFT_INV", "H_DAG", "R_DAG", "DFT_DAG")),
    ("P", 1, ("PHASE", "S")),
    ("P_INV", 1, ("PHASE_INV", "S_INV")),
    ("CNOT", 2, ("SUM", "CX", "C")),
    ("CNOT_INV", 2, ("SUM_INV", "CX_INV", "C_INV")),
    ("CZ", 2, ()),
    ("CZ_INV", 2, ()),
    ("SWAP", 2, ()),
    ("M", 1, ("MEASURE", "COLLAPSE", "MZ")),
    ("M_X", 1, ("MEASURE_X", "MX")),
    ("RESET", 1, ("MR", "MEASURE_RESET", "MEASURE_R")),
    ("N1", 1, ("NOISE1",)),
)

# Integer opcodes shared by the host IR compiler, the CUDA interpreter and the oracle.
OP_I, OP_X, OP_X_INV, OP_Z, OP_Z_INV = 0, 1, 2, 3, 4
OP_H, OP_H_INV, OP_P, OP_P_INV = 5, 6, 7, 8
OP_CNOT, OP_CNOT_INV, OP_CZ, OP_CZ_INV, OP_SWAP = 9, 10, 11, 12, 13
OP_M, OP_M_X, OP_RESET, OP_N1 = 14, 15, 16, 17
NUM_OPS = 18

TWO_QUDIT_OPS = frozenset((OP_CNOT, OP_CNOT_INV, OP_CZ, OP_CZ_INV, OP_SWAP))
MEASURE_OPS = frozenset((OP_M, OP_M_X, OP_RESET))


@dataclass
class Gate:
    """One entry of the gate table (reference: sdim/gatedata.py:6-21)."""

    name: str
    arg_count: int
    gate_id: int
    defaults: Optional[dict] = None

    def __str__(self) -> str:
        return f"{self.name} {self.gate_id}"


@dataclass
class GateData:
    """Name/alias -> gate lookup for one qudit dimension.

    Same public attributes as the reference (`gateMap`, `aliasMap`,
    `num_gates`, `dimension`; sdim/gatedata.py:23-62).  The default parameter
    dict of N1 is `{"channel": "d", "prob": 0.01}` exactly as in the reference
    (gatedata.py:102); the simulator reads `noise_channel` first and falls
    back to `channel`.
    """

    gateMap: Dict[str, Gate] = field(default_factory=dict)
    aliasMap: Dict[str, str] = field(default_factory=dict)
    num_gates: int = 0
    dimension: int = 2

    def __init__(self, dimension: int = 2):
        self.gateMap = {}
        self.aliasMap = {}
        self.num_gates = 0
        self.dimension = dimension
        self._by_id: Dict[int, str] = {}
        for name, arity, aliases in _GATE_TABLE:
            self.add_gate(name, arity)
            self.add_gate_alias(name, aliases)
        self.gateMap["N1"].defaults = {"channel": "d", "prob": 0.01}

    def __str__(self) -> str:
        return "\n".join(str(g) for g in self.gateMap.values())

    def add_gate(self, name: str, arg_count: int) -> None:
        gate = Gate(name, arg_count, self.num_gates)
        self.gateMap[name] = gate
        self._by_id[gate.gate_id] = name
        self.num_gates += 1

    def add_gate_alias(self, name: str, list_alias) -> None:
        for alias in list_alias:
            self.aliasMap[alias] = name

    def get_gate_id(self, gate_name: str) -> Optional[int]:
        gate = self.gateMap.get(gate_name)
        if gate is None:
            primary = self.aliasMap.get(gate_name)
            if primary is None:
                return None
            gate = self.gateMap[primary]
        return gate.gate_id

    def get_gate_name(self, gate_id: int) -> str:
        try:
            return self._by_id[gate_id]
        except KeyError:
            raise ValueError(f"Gate ID {gate_id} not found") from None


_SHARED: Dict[int, GateData] = {}


def shared_gate_data(dimension: int) -> GateData:
    """One GateData per dimension, reused by every Circuit of that dimension."""
    gd = _SHARED.get(dimension)
    if gd is None:
        gd = _SHARED[dimension] = GateData(dimension)
    return gd
