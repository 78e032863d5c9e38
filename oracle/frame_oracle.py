"""CPU ORACLE for the Pauli-frame sampler — TEST INFRASTRUCTURE ONLY.  Not product code.

numpy restatement of the reference's `simulate_frame` (sdim/program.py:45-165) with every random draw passed in.
Parity status: PINNED — tests/golden/frame_cases.json holds outputs of the unmodified reference function (its
`np.random.randint` calls fed from recorded draws, oracle/make_golden_frames.py); tests/test_oracle_golden.py checks
this file against them with `reset_records="reference"`.  `reset_records="physical"` is the corrected RESET record
 that the CUDA kernel implements.
"""
from __future__ import annotations

import numpy as np


def simulate_frames(n, d, ops, reference, z0, zm, noise_ab=None, reset_records="physical"):
    """ops rows (opcode, a, b, slot); reference uint8[n_meas] packed records of the reference shot;
    z0 [shots, n] initial z frames (program.py:64); zm [shots, n_meas] z redraws (program.py:144,156);
    noise_ab [shots, n_noise, 2].  Returns packed records uint8[shots, n_meas]."""
    z0 = np.asarray(z0, dtype=np.int64)
    shots = z0.shape[0]
    x = np.zeros((n, shots), dtype=np.int64)                      # program.py:63
    z = z0.T.copy()
    reference = np.asarray(reference, dtype=np.int64)
    out = np.zeros((shots, len(reference)), dtype=np.uint8)
    for op, a, b, slot in np.asarray(ops, dtype=np.int64).reshape(-1, 4):
        op &= 0xFF
        if op == 5:                                               # program.py:90-93
            x[a], z[a] = (-z[a]) % d, x[a].copy()
        elif op == 6:                                             # :94-97
            x[a], z[a] = z[a].copy(), (-x[a]) % d
        elif op == 7:
            z[a] = (z[a] + x[a]) % d
        elif op == 8:
            z[a] = (z[a] - x[a]) % d
        elif op == 9:                                             # :102-104
            x[b] = (x[b] + x[a]) % d
            z[a] = (z[a] - z[b]) % d
        elif op == 10:
            x[b] = (x[b] - x[a]) % d
            z[a] = (z[a] + z[b]) % d
        elif op == 11:                                            # :108-110
            z[b] = (z[b] + x[a]) % d
            z[a] = (z[a] + x[b]) % d
        elif op == 12:
            z[b] = (z[b] - x[a]) % d
            z[a] = (z[a] - x[b]) % d
        elif op == 13:
            x[[a, b]] = x[[b, a]]
            z[[a, b]] = z[[b, a]]
        elif op in (14, 15):                                      # :121-144
            if op == 15:
                x[a], z[a] = z[a].copy(), (-x[a]) % d
            ref = int(reference[slot])
            out[:, slot] = ((ref & 0x7F) + x[a]) % d | (ref & 0x80)
            z[a] = zm[:, slot]
        elif op == 16:                                            # :146-157
            ref = int(reference[slot])
            if reset_records == "reference":
                out[:, slot] = ref
            else:
                out[:, slot] = ((ref & 0x7F) + x[a]) % d | (ref & 0x80)
            x[a] = 0
            z[a] = zm[:, slot]
        elif op == 17 and noise_ab is not None:                   # :159-162
            x[a] = (x[a] + noise_ab[:, slot, 0]) % d
            z[a] = (z[a] + noise_ab[:, slot, 1]) % d
    return out
