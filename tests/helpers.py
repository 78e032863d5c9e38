"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import numpy as np

from sdim_b200.circuit import Circuit
from sdim_b200.gatedata import shared_gate_data, TWO_QUDIT_OPS

_NAMES = {g.gate_id: name for name, g in shared_gate_data(2).gateMap.items()}


def circuit_from_ops(n, d, ops, noise_params=None):
    """ops rows (opcode, a, b, slot) -> sdim_b200 Circuit (N1 gets prob/channel from noise_params[slot])."""
    c = Circuit(n, d)
    for op, a, b, slot in ops:
        name = _NAMES[int(op)]
        if int(op) in TWO_QUDIT_OPS:
            c.add_gate(name, int(a), int(b))
        elif int(op) == 17:
            prob, ch = (noise_params[int(slot)] if noise_params else (0.5, "d"))
            c.add_gate(name, int(a), prob=prob, noise_channel=ch)
        else:
            c.add_gate(name, int(a))
    return c


def pack_records(recs):
    """[(q, det, m)] -> uint8 row in the device record format."""
    return np.array([(m & 0x7F) | (0x80 if det else 0) for _, det, m in recs], dtype=np.uint8)
