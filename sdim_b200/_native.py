"""ctypes binding of libsdimb (include/sdimb.h).  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

OK, EINVAL, EDIM, EOP, ECUDA, ETOOBIG = 0, -1, -2, -3, -4, -5
FRESH, WRITEBACK, FORCE_GLOBAL, FORCE_RESIDENT, FORCE_LANES, FORCE_PLANES = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20
SCHEDULED, CLUSTER, NO_CLUSTER, TIME_KERNELS, NO_TILE = 0x40, 0x80, 0x100, 0x200, 0x400
OP_BARRIER = 18
KERNEL_NAMES = {0: "lanes-global", 1: "lanes-resident", 2: "planes-resident", 3: "planes-global", 4: "lanes16-global",
                5: "planes-tile"}
REC_DET, REC_VALUE = 0x80, 0x7F

EXPORTED_SYMBOLS = ("sdimb_version", "sdimb_strerror", "sdimb_layout", "sdimb_init", "sdimb_run",
                    "sdimb_export", "sdimb_simulate_host", "sdimb_launch_count", "sdimb_plan", "sdimb_schedule", "sdimb_release_workspace", "sdimb_frames", "sdimb_scratch_bytes", "sdimb_cluster_size",
                    "sdimb_scratch_bytes_shots", "sdimb_tail_run", "sdimb_kernel_times", "sdimb_gate_stream")


class SdimbLayout(C.Structure):
    _fields_ = [("n", C.c_int32), ("d", C.c_int32), ("np", C.c_int32), ("lanes", C.c_int32),
                ("order", C.c_int32), ("phase_order", C.c_int32),
                ("row_bytes", C.c_int64), ("phase_offset", C.c_int64), ("shot_bytes", C.c_int64),
                ("elem_bytes", C.c_int32), ("rec_bytes", C.c_int32)]


class SdimbRunArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("n", C.c_int32), ("d", C.c_int32),
                ("shots", C.c_int64), ("shot_offset", C.c_int64), ("tableau", C.c_void_p),
                ("ops", C.c_void_p), ("n_ops", C.c_int64), ("records", C.c_void_p), ("n_meas", C.c_int64),
                ("rec_stride", C.c_int64), ("replay_meas", C.c_void_p), ("replay_noise", C.c_void_p),
                ("noise_thresh24", C.c_void_p), ("noise_channel", C.c_void_p), ("n_noise", C.c_int64),
                ("seed", C.c_uint64), ("stream", C.c_void_p), ("scratch", C.c_void_p), ("scratch_bytes", C.c_int64),
                ("tail_run_len", C.c_int64), ("gate_stream", C.c_void_p), ("gate_stream_rows", C.c_int64)]


class NativeError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"libsdimb: {what} (code {code})")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load libsdimb.so (built in-tree by `python -m sdim_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing. The tableau path runs only on the CUDA library "
            "(no CPU fallback); build it with `python -m sdim_b200.build`.")
    L = C.CDLL(LIB_PATH)
    L.sdimb_version.restype = C.c_int
    L.sdimb_strerror.restype = C.c_char_p
    L.sdimb_strerror.argtypes = [C.c_int]
    L.sdimb_layout.argtypes = [C.c_int, C.c_int, C.POINTER(SdimbLayout)]
    L.sdimb_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]
    L.sdimb_run.argtypes = [C.POINTER(SdimbRunArgs)]
    L.sdimb_export.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 6 + [C.c_void_p]
    L.sdimb_simulate_host.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_uint64, C.c_uint32, C.POINTER(C.c_float)]
    L.sdimb_launch_count.restype = C.c_int64
    L.sdimb_schedule.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.sdimb_frames.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p]
    L.sdimb_scratch_bytes.argtypes = [C.c_int, C.c_int, C.c_uint32]
    L.sdimb_scratch_bytes.restype = C.c_int64
    L.sdimb_scratch_bytes_shots.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int64]
    L.sdimb_scratch_bytes_shots.restype = C.c_int64
    L.sdimb_tail_run.argtypes = [C.c_void_p, C.c_int64]
    L.sdimb_tail_run.restype = C.c_int64
    L.sdimb_gate_stream.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.sdimb_kernel_times.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.sdimb_cluster_size.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_uint32]
    L.sdimb_plan.argtypes = [C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L


def check(code: int) -> None:
    """Map a negative return code to the exception type the reference raises for that condition."""
    if code == OK:
        return
    what = lib().sdimb_strerror(code).decode()
    if code in (EINVAL, EDIM, EOP, ETOOBIG):
        raise ValueError(what)          # reference errors on this path are ValueError (sdim/program.py:382)
    raise NativeError(code, what)


def layout(n: int, d: int) -> SdimbLayout:
    out = SdimbLayout()
    check(lib().sdimb_layout(n, d, C.byref(out)))
    return out


def plan(n: int, d: int, flags: int):
    """(kernel id, needs_tableau) that sdimb_run would use for these flags."""
    k, need = C.c_int(0), C.c_int(0)
    check(lib().sdimb_plan(n, d, flags, C.byref(k), C.byref(need)))
    return k.value, bool(need.value)


def schedule(n: int, ops):
    """Layered op stream (sdimb_schedule) of an int32[n_ops, 4] array; host-only, no GPU needed."""
    import numpy as np
    ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 4)
    out = np.empty((2 * ops.shape[0] + 1, 4), dtype=np.int32)
    count = C.c_int64(0)
    check(lib().sdimb_schedule(n, ops.ctypes.data if ops.size else None, ops.shape[0], out.ctypes.data,
                               out.shape[0], C.byref(count)))
    return out[: count.value].copy()


def kernel_times():
    """(interpreter ms, tail-run ms) of the last run that carried TIME_KERNELS and split its stream; None otherwise."""
    a, b = C.c_float(0.0), C.c_float(0.0)
    if lib().sdimb_kernel_times(C.byref(a), C.byref(b)) != OK:
        return None
    return float(a.value), float(b.value)


def gate_stream(n: int, d: int, sched_front):
    """Pre-decoded per-warp streams (sdimb_gate_stream) of a gate-only stretch of a scheduled stream, int32[rows, 4];
    None when the library does not compile this shape (d > 3, n > 512, a measurement in the stretch)."""
    import numpy as np
    sched_front = np.ascontiguousarray(sched_front, dtype=np.int32).reshape(-1, 4)
    ptr = sched_front.ctypes.data if sched_front.size else None
    rows = C.c_int64(0)
    if lib().sdimb_gate_stream(n, d, ptr, sched_front.shape[0], None, 0, C.byref(rows)) != OK or rows.value <= 0:
        return None
    out = np.empty((rows.value, 4), dtype=np.int32)
    check(lib().sdimb_gate_stream(n, d, ptr, sched_front.shape[0], out.ctypes.data, out.shape[0], C.byref(rows)))
    return out


def tail_run(sched) -> int:
    """Length of the marked run of M ops that ends a scheduled stream (sdimb_tail_run), 0 if there is none."""
    import numpy as np
    sched = np.ascontiguousarray(sched, dtype=np.int32).reshape(-1, 4)
    return int(lib().sdimb_tail_run(sched.ctypes.data if sched.size else None, sched.shape[0]))
