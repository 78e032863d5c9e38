"""TEST INFRASTRUCTURE: a CPU stand-in for sdim_b200.engine.TableauEngine built on the numpy oracle.

It lets the CPU suite exercise the HOST logic of sdim_b200.program (result grouping, host-stepped modes, snapshots,
apply_gate, gate folding) without a GPU.  It is never imported by the product; the GPU tests run the same scenarios
on the real engine."""
import numpy as np
import torch

from oracle.tableau_oracle import OracleTableau, run_shot
from sdim_b200 import _native as N
from sdim_b200 import records as R
from oracle.frame_oracle import simulate_frames
from sdim_b200.rng import frame_z0_draws, frame_zm_draws, measurement_draws, noise_draws
from sdim_b200.tableau import ExtendedTableau


def unpack(img: np.ndarray, n: int, d: int, np_pad: int) -> OracleTableau:
    """Inverse of ExtendedTableau.pack: device image of one shot (include/sdimb.h) -> oracle tableau."""
    W = 2 * np_pad
    img = np.ascontiguousarray(img)
    if d > 127:
        img = img.view(np.uint16)                       # uint16 lanes: two bytes per entry
    img = img.astype(np.int64).reshape(2 * n + 1, W)
    rows = img[: 2 * n].reshape(n, 2, W)
    t = OracleTableau(n, d)
    t.x, t.z = rows[:, 0, :n].copy(), rows[:, 1, :n].copy()
    t.dx, t.dz = rows[:, 0, np_pad:np_pad + n].copy(), rows[:, 1, np_pad:np_pad + n].copy()
    t.p, t.dp = img[2 * n, :n].copy(), img[2 * n, np_pad:np_pad + n].copy()
    return t


def pack(t: OracleTableau, np_pad: int) -> np.ndarray:
    x, z, p, dx, dz, dp = t.arrays()
    return ExtendedTableau(t.n, t.d, phase_vector=p, z_block=z, x_block=x, destab_phase_vector=dp,
                           destab_z_block=dz, destab_x_block=dx).pack(np_pad)


class FakeEngine:
    def __init__(self, prog, device=None):
        self.prog, self.device = prog, torch.device("cpu")
        self.layout = N.layout(prog.num_qudits, prog.dimension)
        self.tableau, self.tableau_shots = None, 0
        self.rec_dtype = R.torch_dtype(prog.dimension)
        self.runs = []                                   # (shots, op_range, n_ops) of every call, for assertions

    def plan(self, mode=None, fresh=True, keep_tableau=False):
        return "fake", (not fresh) or keep_tableau

    def device_bytes_per_shot(self, shots, mode=None, fresh=True, keep_tableau=False):
        return self.layout.shot_bytes if ((not fresh) or keep_tableau) else 0

    def alloc_tableau(self, shots):
        return torch.zeros((shots, self.layout.shot_bytes), dtype=torch.uint8)

    def init_tableau(self, tab):
        image = torch.from_numpy(pack(OracleTableau(self.prog.num_qudits, self.prog.dimension), self.layout.np))
        tab[:] = image

    def run(self, shots, shot_offset=0, seed=0, replay_meas=None, replay_noise=None, keep_tableau=False, mode=None,
            tableau=None, fresh=True, op_range=None, records=None):
        prog, n, d = self.prog, self.prog.num_qudits, self.prog.dimension
        lo, hi = (0, prog.n_ops) if op_range is None else op_range
        self.runs.append((shots, op_range, prog.n_ops))
        if records is None:
            records = torch.zeros((shots, prog.n_meas), dtype=self.rec_dtype)
        det_bit, val_mask = R.masks(R.np_dtype(d))
        if (not fresh or keep_tableau) and tableau is None:
            tableau = self.alloc_tableau(shots)
        for s in range(shots):
            gid = shot_offset + s
            md = R.unsigned(replay_meas[s].numpy()) if replay_meas is not None else measurement_draws(seed, d, [gid], prog.n_meas)[0]
            nd = None
            if prog.n_noise:
                nd = R.unsigned(replay_noise[s].numpy()) if replay_noise is not None else \
                    noise_draws(seed, d, [gid], prog.noise_thresh24, prog.noise_channel)[0]
            start = None if fresh else unpack(tableau[s].numpy(), n, d, self.layout.np)
            recs, t = run_shot(n, d, prog.ops[lo:hi], lambda k: int(md[k]), nd, tableau=start)
            slots = [int(o[3]) for o in prog.ops[lo:hi] if int(o[0]) in (14, 15, 16)]
            for slot, (_, det, m) in zip(slots, recs):
                v = (m & val_mask) | (det_bit if det else 0)
                records[s, slot] = v - 0x10000 if v >= 0x8000 else v        # int16 bits of a uint16 record
            if keep_tableau:
                tableau[s] = torch.from_numpy(pack(t, self.layout.np))
        if keep_tableau:
            self.tableau, self.tableau_shots = tableau, shots
        return records

    def export(self, tableau, shot):
        t = unpack(tableau[shot].numpy(), self.prog.num_qudits, self.prog.dimension, self.layout.np)
        x, z, p, dx, dz, dp = t.arrays()
        return {"x": x, "z": z, "p": p, "dx": dx, "dz": dz, "dp": dp}

    def run_frames(self, shots, reference, shot_offset=1, seed=0, replay_z0=None, replay_zm=None, replay_noise=None):
        prog, n, d = self.prog, self.prog.num_qudits, self.prog.dimension
        ids = np.arange(shot_offset, shot_offset + shots)
        z0 = replay_z0.numpy() if replay_z0 is not None else frame_z0_draws(seed, d, ids, n)
        zm = replay_zm.numpy() if replay_zm is not None else frame_zm_draws(seed, d, ids, prog.n_meas)
        nd = None
        if prog.n_noise:
            nd = replay_noise.numpy() if replay_noise is not None else \
                noise_draws(seed, d, ids, prog.noise_thresh24, prog.noise_channel)
        return torch.from_numpy(simulate_frames(n, d, prog.ops, reference.numpy(), z0, zm, nd))
