"""TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/liboracle.so (the C restatement, oracle/oracle.c)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def _load():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.oracle_run.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p,
                                 C.c_void_p, C.c_int]
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def threads() -> int:
    """Host threads the baseline uses: every core this process may run on (torchrun exports OMP_NUM_THREADS=1,
    which must not shrink the CPU baseline, so the count is passed to OpenMP explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _p(a):
    return None if a is None or a.size == 0 else a.ctypes.data


def run(n, d, ops, shots, shot_offset=0, seed=0, replay_meas=None, replay_noise=None, thresh24=None, channel=None,
        want_final=False, nthreads=0, meas_nnz=None):
    """Returns (records uint8[shots, n_meas] (uint16 for d > 127), final dict or None).  Replay arrays as in the CUDA path."""
    ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 4)
    n_meas = int(np.isin(ops[:, 0], (14, 15, 16)).sum())
    n_noise = int((ops[:, 0] == 17).sum())
    rdt = np.uint16 if d > 127 else np.uint8       # element width of records / replay arrays (include/sdimb.h)
    rec = np.zeros((shots, n_meas), dtype=rdt)
    rm = None if replay_meas is None else np.ascontiguousarray(replay_meas, dtype=rdt)
    rn = None if replay_noise is None else np.ascontiguousarray(replay_noise, dtype=rdt)
    th = None if thresh24 is None else np.ascontiguousarray(thresh24, dtype=np.uint32)
    ch = None if channel is None else np.ascontiguousarray(channel, dtype=np.uint8)
    if n_noise and rn is None and (th is None or ch is None):
        raise ValueError("noise events need replay_noise or (thresh24, channel)")
    final = np.zeros(4 * n * n + 2 * n, dtype=np.int64) if want_final else None
    rc = _load().oracle_run(n, d, shots, shot_offset, _p(ops), ops.shape[0], _p(rec), n_meas, _p(rm), _p(rn),
                            _p(th), _p(ch), n_noise, seed & 0xFFFFFFFFFFFFFFFF, _p(final), _p(meas_nnz),
                            nthreads if nthreads > 0 else threads())
    if rc != 0:
        raise RuntimeError(f"oracle_run failed ({rc})")
    out = None
    if want_final:
        nn = n * n
        out = {"x": final[:nn].reshape(n, n), "z": final[nn:2 * nn].reshape(n, n),
               "dx": final[2 * nn:3 * nn].reshape(n, n), "dz": final[3 * nn:4 * nn].reshape(n, n),
               "p": final[4 * nn:4 * nn + n], "dp": final[4 * nn + n:]}
    return rec, out


def run_philox(prog, shots, shot_offset, seed, nthreads=0):
    """Free-running mode on a sdim_b200 CompiledProgram-like object (fields ops, noise_thresh24, noise_channel)."""
    rec, _ = run(prog.num_qudits, prog.dimension, prog.ops, shots, shot_offset, seed,
                 thresh24=prog.noise_thresh24, channel=prog.noise_channel, nthreads=nthreads)
    return rec


def measurement_factor_counts(prog, seed=0):
    """int32[n_meas]: generators with a non-zero factor in each measurement of one shot (shot-invariant: the
    x/z blocks do not depend on outcomes or Pauli noise)."""
    nnz = np.zeros(max(prog.n_meas, 1), dtype=np.int32)
    run(prog.num_qudits, prog.dimension, prog.ops, 1, 0, seed, thresh24=prog.noise_thresh24,
        channel=prog.noise_channel, meas_nnz=nnz)
    return nnz[: prog.n_meas]
