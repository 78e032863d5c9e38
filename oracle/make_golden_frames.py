"""TEST INFRASTRUCTURE ONLY: golden vectors for the Pauli-frame sampler from the UNMODIFIED reference.

    python oracle/make_golden_frames.py        # needs /root/reference (this container only)

Calls the reference's own `simulate_frame` (sdim/program.py:45-165) on seeded random circuits.  Its only randomness is
`np.random.randint(0, d, size=...)` — one call for the initial z frame, one per M / M_X / RESET — so the module's `np`
is swapped for a proxy whose `random.randint` hands out pre-drawn arrays in call order; the draws are stored with the
outputs so that any implementation can be replayed against them.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402
from oracle.make_golden import random_ops, final_measure_all  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frame_cases.json")


class _FeedRandom:
    def __init__(self, arrays):
        self.arrays, self.i = arrays, 0

    def randint(self, low, high, size=None):
        a = self.arrays[self.i]
        self.i += 1
        assert tuple(np.shape(a)) == (tuple(size) if isinstance(size, tuple) else (size,)), (np.shape(a), size)
        return np.array(a, dtype=np.int64)


class _NumpyProxy:
    def __init__(self, feed):
        self.random = feed

    def __getattr__(self, name):
        return getattr(np, name)


def make_case(seed, n, d, depth, shots):
    sdim = rh.load_reference()
    import sdim.program as sp
    rng = random.Random(seed)
    ops, k, j = random_ops(rng, n, depth, d, p_meas=0.15, p_noise=0.12)
    k = final_measure_all(ops, n, k)
    # the reference's frame path needs every qudit 0 measurement table to exist (program.py:398-399): final M on all
    recs, _ = rh.ref_run(n, d, ops, None, draw_seed=seed)          # noiseless reference shot (N1 ignored)
    # reference_results structured array [n_qudits, rounds]
    count, per = {}, [[] for _ in range(n)]
    for q, det, m in recs:
        per[q].append((q, len(per[q]), 0, det, m))
    rounds = max(len(p) for p in per)
    ref_arr = np.zeros((n, rounds), dtype=sp.MEASUREMENT_DTYPE)
    for q in range(n):
        for r, t in enumerate(per[q]):
            ref_arr[q, r] = t
    ir = np.array([(o[0], o[1], o[2]) for o in ops if o[0] != 0],
                  dtype=np.dtype([("gate_id", np.int64), ("qudit_index", np.int64), ("target_index", np.int64)]))
    npr = np.random.RandomState(seed)
    z0 = npr.randint(0, d, size=(n, shots))
    meas_ops = [o for o in ops if o[0] in (14, 15, 16)]
    zm = [npr.randint(0, d, size=shots) for _ in meas_ops]
    noise = npr.randint(0, d, size=(max(j, 1), shots, 2)) * (npr.rand(max(j, 1), shots, 1) < 0.6)
    feed = _FeedRandom([z0] + zm)
    saved = sp.np
    sp.np = _NumpyProxy(feed)
    try:
        fr = sp.simulate_frame(ir, ref_arr, n, d, shots, noise.astype(np.int64))
    finally:
        sp.np = saved
    # chronological records of every extra shot
    cnt = {}
    out = np.zeros((shots, len(meas_ops)), dtype=np.int64)
    for kk, o in enumerate(meas_ops):
        q = o[1]
        r = cnt.get(q, 0)
        cnt[q] = r + 1
        out[:, kk] = fr[q, r]["measurement_value"] | (fr[q, r]["deterministic"].astype(np.int64) << 7)
    return {"seed": seed, "n": n, "d": d, "ops": ops,
            "reference": [(m & 0x7F) | (0x80 if det else 0) for _, det, m in recs],
            "z0": z0.T.tolist(), "zm": np.array(zm).T.reshape(shots, len(meas_ops)).tolist(),
            "noise_ab": noise[:j].transpose(1, 0, 2).tolist() if j else [],
            "records": out.tolist()}


def main():
    cases = []
    seed = 5000
    for d in (2, 3, 5, 7):
        for n, depth in ((1, 20), (2, 40), (4, 80), (7, 140)):
            cases.append(make_case(seed, n, d, depth, shots=6))
            seed += 1
    with open(GOLDEN, "w") as fh:
        json.dump({"generator": "oracle/make_golden_frames.py",
                   "reference": "sdim.program.simulate_frame @ /root/reference (RESET records the reference value)",
                   "cases": cases}, fh, separators=(",", ":"))
    print(f"wrote {len(cases)} cases -> {GOLDEN} ({os.path.getsize(GOLDEN)} bytes)")


if __name__ == "__main__":
    main()
