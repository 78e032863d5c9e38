"""Host IR compiler: Circuit.operations -> flat int32 op stream for the CUDA interpreter.

Plays the role of the reference's `Program._build_ir` (sdim/program.py:456-526),
which lowers a circuit to `(gate_id, qudit, target)` triples plus pre-sampled
noise.  Here the op rows carry a fourth field, the *event slot*:

    ops[i] = (opcode, a, b, slot)        int32[n_ops, 4]

  opcode  gate id 0..17, same numbering as the reference gate table
  a, b    qudit / target (b = -1 for single-qudit ops)
  slot    M / M_X / RESET: chronological measurement index k (column of the record
          matrix and Philox slot);  N1: noise-event index j;  otherwise -1

`I` gates are dropped from the stream (they still count as user gates for the
shot·gates metric), like program.py:479-480.  Noise is not pre-sampled: each N1
row points at `(thresh24[j], channel[j])` and the device draws per shot.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List

import numpy as np

from .circuit import Circuit
from .gatedata import (MEASURE_OPS, NUM_OPS, OP_I, OP_N1, TWO_QUDIT_OPS)
from .rng import CHANNEL_CODES, prob_to_thresh24

MAX_DIMENSION = 32749      # largest prime below 2^15: uint8 lanes up to 127, uint16 lanes above (csrc/wide.cuh)


def is_prime(d: int) -> bool:
    if d < 2:
        return False
    if d % 2 == 0:
        return d == 2
    f = 3
    while f * f <= d:
        if d % f == 0:
            return False
        f += 2
    return True


@dataclass
class CompiledProgram:
    num_qudits: int
    dimension: int
    ops: np.ndarray                       # int32 [n_ops, 4]
    n_user_gates: int                     # len(operations) as written, the metric's "gates"
    meas_qudit: np.ndarray                # int32 [n_meas]
    meas_round: np.ndarray                # int32 [n_meas]   per-qudit running count (program.py:326-327)
    meas_opcode: np.ndarray               # int32 [n_meas]
    noise_qudit: np.ndarray               # int32 [n_noise]
    noise_thresh24: np.ndarray            # uint32 [n_noise]
    noise_channel: np.ndarray             # uint8 [n_noise]
    noise_prob: np.ndarray                # float64 [n_noise]
    rounds_per_qudit: List[int] = field(default_factory=list)

    @property
    def n_ops(self) -> int:
        return int(self.ops.shape[0])

    @property
    def n_meas(self) -> int:
        return int(self.meas_qudit.shape[0])

    @property
    def n_noise(self) -> int:
        return int(self.noise_qudit.shape[0])


def _noise_params(params) -> tuple:
    params = params or {}
    # The reference reads params['noise_channel'] (program.py:486) while the gate default is
    # stored under 'channel' (gatedata.py:102) and would raise KeyError; accept both (B-4).
    channel = params.get("noise_channel", params.get("channel", "d"))
    if channel not in CHANNEL_CODES:
        raise ValueError(f"Unknown noise channel {channel!r}; expected 'd', 'f' or 'p'")
    prob = float(params.get("prob", 0.01))
    if not 0.0 <= prob <= 1.0:
        raise ValueError(f"Noise probability {prob} outside [0, 1]")
    return CHANNEL_CODES[channel], prob


def compile_circuits(circuits: Iterable[Circuit]) -> CompiledProgram:
    """Lower the circuits of a Program (run back to back, program.py:311-312) to one op stream."""
    circuits = list(circuits)
    n = max(c.num_qudits for c in circuits)
    d = circuits[0].dimension
    if any(c.dimension != d for c in circuits):
        raise ValueError("Circuits must have the same dimension")
    rows = []
    meas_q, meas_r, meas_op = [], [], []
    noise_q, noise_t, noise_c, noise_p = [], [], [], []
    rounds = [0] * n
    n_user = 0
    for circuit in circuits:
        for ins in circuit.operations:
            n_user += 1
            op = ins.gate_id
            if op is None or not 0 <= op < NUM_OPS:
                raise ValueError("Invalid gate value")          # program.py:381-382
            a = int(ins.qudit_index)
            b = -1 if ins.target_index is None else int(ins.target_index)
            if not 0 <= a < n or (b != -1 and not 0 <= b < n):
                raise ValueError(f"Qudit index out of range for gate {ins.name} ({a}, {b}); circuit has {n} qudits")
            if op in TWO_QUDIT_OPS:
                if b == -1:
                    raise ValueError(f"Gate {ins.name} needs a target qudit")
                if a == b:
                    raise ValueError(f"Gate {ins.name} needs two distinct qudits, got {a} twice")
            if op == OP_I:
                continue
            slot = -1
            if op in MEASURE_OPS:
                slot = len(meas_q)
                meas_q.append(a)
                meas_r.append(rounds[a])
                meas_op.append(op)
                rounds[a] += 1
            elif op == OP_N1:
                code, prob = _noise_params(ins.params)
                slot = len(noise_q)
                noise_q.append(a)
                noise_t.append(prob_to_thresh24(prob))
                noise_c.append(code)
                noise_p.append(prob)
            rows.append((op, a, b if op in TWO_QUDIT_OPS else -1, slot))
    ops = np.array(rows, dtype=np.int32).reshape(-1, 4)
    return CompiledProgram(
        num_qudits=n, dimension=d, ops=ops, n_user_gates=n_user,
        meas_qudit=np.array(meas_q, dtype=np.int32), meas_round=np.array(meas_r, dtype=np.int32),
        meas_opcode=np.array(meas_op, dtype=np.int32),
        noise_qudit=np.array(noise_q, dtype=np.int32),
        noise_thresh24=np.array(noise_t, dtype=np.uint32),
        noise_channel=np.array(noise_c, dtype=np.uint8),
        noise_prob=np.array(noise_p, dtype=np.float64),
        rounds_per_qudit=rounds,
    )
