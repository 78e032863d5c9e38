"""Shot sharding across the GPUs of one box (one process per GPU, torch.distributed).

Shots are independent (the reference runs them one after another, sdim/program.py:308), so the
path shards with no data-path collective: rank r of R simulates the contiguous global shot range
[r*S/R, (r+1)*S/R) — Philox counters use the GLOBAL shot id, so the records do not depend on R —
and the only exchange is one all_gather of the packed uint8 record matrix at the end (NCCL over
NVLink on GPUs; gloo in the CPU tests, which inject a stand-in runner).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(shots: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank`; the first shots % world ranks take one extra shot."""
    base, extra = divmod(shots, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(local, shots: int, n_meas: int, group=None):
    """all_gather the per-rank [local_shots, n_meas] record tensors (uint8; int16 bits for d > 127) into [shots, n_meas] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    per = -(-shots // world)                         # ceil: ranks pad to a common size
    padded = torch.zeros((per, n_meas), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * per, n_meas), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_range(shots, r, world)
        pieces.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(pieces, dim=0)


def broadcast_seed(seed: Optional[int], group=None) -> int:
    """One seed for the whole job: rank 0's value (drawn there from `random` when the caller gave none) reaches every
    rank, so that the gathered table is ONE run whose records depend on the global shot id only."""
    import random
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return random.getrandbits(63) if seed is None else int(seed)
    box = [random.getrandbits(63) if seed is None else int(seed)]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return int(box[0])


def simulate_sharded(program, compiled, shots: int, seed: int, mode: Optional[str] = None,
                     runner: Optional[Callable] = None, group=None, shot_offset: int = 0,
                     replay_meas=None, replay_noise=None) -> np.ndarray:
    """Simulate this rank's shot range and return the gathered uint8[shots, n_meas] records (packed bytes).

    The shard runs through `Program._run_local` — the same initial tableau, replay slices, global shot ids
    (`shot_offset + lo + local`) and wave sizing as a single-process run — with the records left on the device, so the
    only exchange is the NCCL all-gather over NVLink.  `seed` must already be the job-wide seed (`broadcast_seed`).
    runner(lo, hi) -> torch uint8 [hi-lo, n_meas] overrides the GPU engine (CPU tests).
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        world, rank = 1, 0
    else:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    for name, arr, width in (("replay_meas", replay_meas, compiled.n_meas), ("replay_noise", replay_noise, compiled.n_noise)):
        if arr is not None and (np.asarray(arr).shape[0] != shots or np.asarray(arr).shape[1] != width):
            raise ValueError(f"{name} must cover all {shots} shots of the job (every rank passes the full array)")
    lo, hi = shard_range(shots, rank, world)
    if runner is not None:
        local = runner(lo, hi)
    else:
        local = program._run_local(compiled, hi - lo, shot_offset + lo, seed,
                                   None if replay_meas is None else np.asarray(replay_meas)[lo:hi],
                                   None if replay_noise is None else np.asarray(replay_noise)[lo:hi],
                                   mode, on_device=True)
    if world > 1:
        local = gather_records(local, shots, compiled.n_meas, group)
    return local.cpu().numpy() if isinstance(local, torch.Tensor) else np.asarray(local)
