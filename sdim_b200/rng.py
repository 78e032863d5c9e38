"""Host mirror of the device random streams (Philox4x32-10, counter based).

The CUDA interpreter draws every random number from Philox keyed by
`(seed, global shot id, event slot, stream)`, so the record of a shot does not
depend on which GPU or which launch simulated it.  This module reproduces those
draws bit-for-bit on the host: tests replay them into the CPU oracle to check
the free-running mode exactly, and `Program` uses it to explain a result.

Streams:
  STREAM_MEAS   slot = chronological measurement index k.  Outcome of a random
                measurement (reference draws random.choice(range(d)),
                sdim/tableau/tableau_prime.py:332):  m = (w0 * d) >> 32.
  STREAM_NOISE  slot = N1 event index j.  Distribution of the reference's
                `_build_ir` (sdim/program.py:486-507): the event fires iff
                (w0 >> 8) >= thresh24 with thresh24 = round((1 - prob) * 2^24);
                then channel 'd': r = 1 + ((w1 * (d*d - 1)) >> 32), a = r % d,
                b = r // d; 'f': a = 1 + ((w1 * (d - 1)) >> 32), b = 0; 'p': the
                same on b.
"""
from __future__ import annotations

import numpy as np

STREAM_MEAS = 0
STREAM_NOISE = 1
STREAM_FRAME_Z0 = 2      # Pauli-frame sampler: initial z frame, slot = qudit            (sdim/program.py:64)
STREAM_FRAME_ZM = 3      # Pauli-frame sampler: z redraw after measurement k, slot = k   (sdim/program.py:144,156)

CHANNEL_D, CHANNEL_F, CHANNEL_P = 0, 1, 2
CHANNEL_CODES = {"d": CHANNEL_D, "f": CHANNEL_F, "p": CHANNEL_P}

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32(c0, c1, c2, c3, k0: int, k1: int):
    """Ten-round Philox4x32 on broadcastable uint32 counter arrays; returns four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)))
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> _S32) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK
        n2 = (p0 >> _S32) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _words(seed: int, shots, slots, stream: int):
    shots = np.asarray(shots, dtype=np.uint64)
    slots = np.asarray(slots, dtype=np.uint64)
    return philox4x32(shots & _MASK, shots >> _S32, slots, np.uint64(stream),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def prob_to_thresh24(prob: float) -> int:
    """24-bit no-fire threshold for an N1 event of probability `prob`."""
    t = int(round((1.0 - float(prob)) * (1 << 24)))
    return min(max(t, 0), 1 << 24)


def measurement_draws(seed: int, d: int, shot_ids, n_meas: int) -> np.ndarray:
    """uint8[len(shot_ids), n_meas]: the value measurement k takes in each shot if it is random."""
    shot_ids = np.asarray(shot_ids, dtype=np.uint64).reshape(-1, 1)
    slots = np.arange(n_meas, dtype=np.uint64).reshape(1, -1)
    w0 = _words(seed, shot_ids, slots, STREAM_MEAS)[0].astype(np.uint64)
    return ((w0 * np.uint64(d)) >> _S32).astype(np.uint16 if d > 127 else np.uint8)


def noise_draws(seed: int, d: int, shot_ids, thresh24, channel) -> np.ndarray:
    """uint8[len(shot_ids), n_noise, 2]: Pauli exponents (a, b) of every N1 event in each shot."""
    thresh24 = np.asarray(thresh24, dtype=np.uint64).reshape(1, -1)
    channel = np.asarray(channel, dtype=np.uint8).reshape(1, -1)
    shot_ids = np.asarray(shot_ids, dtype=np.uint64).reshape(-1, 1)
    slots = np.arange(thresh24.shape[1], dtype=np.uint64).reshape(1, -1)
    w = _words(seed, shot_ids, slots, STREAM_NOISE)
    w0, w1 = w[0].astype(np.uint64), w[1].astype(np.uint64)
    fire = (w0 >> np.uint64(8)) >= thresh24
    r = np.uint64(1) + ((w1 * np.uint64(d * d - 1)) >> _S32)
    e = np.uint64(1) + ((w1 * np.uint64(d - 1)) >> _S32)
    a = np.where(channel == CHANNEL_D, r % np.uint64(d), np.where(channel == CHANNEL_F, e, 0))
    b = np.where(channel == CHANNEL_D, r // np.uint64(d), np.where(channel == CHANNEL_P, e, 0))
    out = np.stack((np.where(fire, a, 0), np.where(fire, b, 0)), axis=-1)
    return out.astype(np.uint16 if d > 127 else np.uint8)


def frame_z0_draws(seed: int, d: int, shot_ids, n: int) -> np.ndarray:
    """uint8[len(shot_ids), n]: initial z frame of every shot (frame sampler)."""
    shot_ids = np.asarray(shot_ids, dtype=np.uint64).reshape(-1, 1)
    slots = np.arange(n, dtype=np.uint64).reshape(1, -1)
    w0 = _words(seed, shot_ids, slots, STREAM_FRAME_Z0)[0].astype(np.uint64)
    return ((w0 * np.uint64(d)) >> _S32).astype(np.uint8)


def frame_zm_draws(seed: int, d: int, shot_ids, n_meas: int) -> np.ndarray:
    """uint8[len(shot_ids), n_meas]: z frame redrawn on the measured qudit after measurement k (frame sampler)."""
    shot_ids = np.asarray(shot_ids, dtype=np.uint64).reshape(-1, 1)
    slots = np.arange(n_meas, dtype=np.uint64).reshape(1, -1)
    w0 = _words(seed, shot_ids, slots, STREAM_FRAME_ZM)[0].astype(np.uint64)
    return ((w0 * np.uint64(d)) >> _S32).astype(np.uint8)
