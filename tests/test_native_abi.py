"""CPU: the C-ABI library loads and exports every symbol include/sdimb.h declares; argument validation.
No compute call is made here (there is no GPU in the CPU test environment)."""
import ctypes as C
import os
import re

import pytest

from sdim_b200 import _native as N
from sdim_b200.build import LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "sdimb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdimb_[a-z_0-9]+)\s*\(", text)))


def test_library_exists_and_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), "build with `python -m sdim_b200.build` (or __graft_entry__.build())"
    lib = N.lib()
    declared = _declared_functions()
    assert set(declared) == set(N.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sdimb_version() == 4


def test_layout_matches_header_contract():
    L = N.layout(256, 3)
    assert (L.np, L.lanes, L.row_bytes, L.phase_offset, L.shot_bytes) == (256, 512, 1024, 262144, 262656)
    assert (L.order, L.phase_order) == (3, 1)
    L = N.layout(13, 2)
    assert (L.np, L.lanes, L.order, L.phase_order) == (16, 32, 4, 2)
    L = N.layout(97, 2)
    assert L.np == 112 and L.shot_bytes == 97 * 448 + 224
    L = N.layout(4096, 7)
    assert L.shot_bytes == 4096 * 16384 + 8192                   # 64 MiB + 8 KiB


def test_layout_of_the_uint16_lanes():
    """127 < d < 2^15: two bytes per entry and per record, everything else as the uint8 store."""
    a, b = N.layout(20, 127), N.layout(20, 131)
    assert (a.elem_bytes, a.rec_bytes, b.elem_bytes, b.rec_bytes) == (1, 1, 2, 2)
    assert (b.np, b.lanes, b.order, b.phase_order) == (a.np, a.lanes, 131, 1)
    assert (b.row_bytes, b.phase_offset, b.shot_bytes) == (2 * a.row_bytes, 2 * a.phase_offset, 2 * a.shot_bytes)
    assert N.plan(20, 131, 0) == (4, True) and N.KERNEL_NAMES[4] == "lanes16-global"
    assert N.layout(3, 32749).order == 32749


def test_error_codes_map_to_value_errors():
    for n, d in ((4, 4), (4, 1), (4, 128), (4, 32771), (4, 32767), (20000, 131), (0, 3)):
        with pytest.raises(ValueError):
            N.layout(n, d)
    lib = N.lib()
    assert lib.sdimb_strerror(N.EOP) == b"Invalid gate value"      # sdim/program.py:382
    assert lib.sdimb_run(None) == N.EINVAL
    a = N.SdimbRunArgs()
    a.struct_size = 3
    assert lib.sdimb_run(C.byref(a)) == N.EINVAL
    a.struct_size = C.sizeof(N.SdimbRunArgs)
    a.n, a.d, a.shots = 4, 6, 1
    assert lib.sdimb_run(C.byref(a)) == N.EDIM
    a.d, a.shots = 3, -1
    assert lib.sdimb_run(C.byref(a)) == N.EINVAL
    # callers built against the struct of ABI versions 1-3 (without tail_run_len / without the gate-stream fields) are
    # accepted; any other size is not.  shots = 0 returns before anything touches the device.
    a.shots = 0
    full = C.sizeof(N.SdimbRunArgs)
    for size, want in ((full, N.OK), (N.SdimbRunArgs.gate_stream.offset, N.OK), (N.SdimbRunArgs.tail_run_len.offset, N.OK),
                       (full - 4, N.EINVAL), (full + 8, N.EINVAL)):
        a.struct_size = size
        assert lib.sdimb_run(C.byref(a)) == want, size
    # host entry validates the op stream before touching the device
    import numpy as np
    ops = np.array([[99, 0, -1, -1]], dtype=np.int32)
    rc = lib.sdimb_simulate_host(4, 3, 1, 0, ops.ctypes.data, 1, None, 0, None, None, None, None, 0, 0, 0, None)
    assert rc == N.EOP
    ops = np.array([[9, 0, 0, -1]], dtype=np.int32)
    rc = lib.sdimb_simulate_host(4, 3, 1, 0, ops.ctypes.data, 1, None, 0, None, None, None, None, 0, 0, 0, None)
    assert rc == N.EOP


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sdim_b200 import Circuit, Program
    c = Circuit(2, 3)
    c.add_gate("M", [0, 1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Program(c).simulate()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sdim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|liboracle", text, flags=re.M), \
                    f"{f} imports or loads the oracle"


def test_kernel_selection_rule_without_a_gpu():
    """sdimb_plan is host logic: d = 2, 3 run the bit-plane interpreter — image in shared memory
    while four or more fit per SM, else in a per-CTA global slab; other primes run uint8 lanes, resident while one
    tableau fits in shared memory."""
    from sdim_b200 import _native as N
    fresh = N.FRESH
    names = {k: N.KERNEL_NAMES[N.plan(*k, fresh)[0]] for k in
             ((64, 3), (160, 3), (200, 3), (224, 3), (256, 3), (500, 3), (97, 2), (290, 2), (320, 2), (700, 2),
              (64, 5), (200, 7), (256, 5), (4096, 7), (6, 127))}
    assert names[(64, 3)] == "planes-tile" and N.KERNEL_NAMES[N.plan(49, 3, fresh)[0]] == "planes-tile"     # n <= 128: several shots per warp
    assert N.KERNEL_NAMES[N.plan(64, 3, fresh | N.NO_TILE)[0]] == "planes-resident"
    assert N.KERNEL_NAMES[N.plan(128, 2, fresh | N.FORCE_PLANES)[0]] == "planes-tile"
    assert N.KERNEL_NAMES[N.plan(129, 2, fresh | N.FORCE_PLANES)[0]] == "planes-resident"
    assert names[(160, 3)] == names[(200, 3)] == "planes-resident"
    assert names[(224, 3)] == names[(256, 3)] == names[(500, 3)] == "planes-global"
    assert names[(97, 2)] == "planes-tile" and names[(290, 2)] == "planes-resident"
    assert names[(320, 2)] == names[(700, 2)] == "planes-global"
    assert names[(64, 5)] == names[(200, 7)] == names[(6, 127)] == "lanes-resident"
    assert names[(256, 5)] == names[(4096, 7)] == "lanes-global"
    # a fresh run that keeps nothing needs no HBM store unless the lanes run from global memory
    assert N.plan(256, 3, fresh) == (3, False) and N.plan(64, 5, fresh) == (1, False)
    assert N.plan(256, 5, fresh) == (0, True) and N.plan(256, 3, fresh | N.WRITEBACK)[1] is True
    # forced modes
    assert N.KERNEL_NAMES[N.plan(256, 3, fresh | N.FORCE_PLANES)[0]] == "planes-resident"
    assert N.KERNEL_NAMES[N.plan(256, 3, fresh | N.FORCE_LANES)[0]] == "lanes-global"
    assert N.KERNEL_NAMES[N.plan(64, 3, fresh | N.FORCE_PLANES | N.FORCE_GLOBAL)[0]] == "planes-global"
    with pytest.raises(ValueError):
        N.plan(500, 3, fresh | N.FORCE_PLANES)          # does not fit in shared memory
    with pytest.raises(ValueError):
        N.plan(256, 5, fresh | N.FORCE_RESIDENT)


def test_host_entry_validates_before_touching_cuda():
    """sdimb_simulate_host checks dimension, sizes and every op row first: the error a user gets for a bad stream
    does not depend on a GPU being present ("Invalid gate value", sdim/program.py:381-382)."""
    import numpy as np
    from sdim_b200 import _native as N
    L = N.lib()
    rec = np.zeros((2, 4), dtype=np.uint8)

    def call(n, d, ops, n_meas=0, n_noise=0, thr=None, ch=None, shots=2):
        ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 4)
        return L.sdimb_simulate_host(n, d, shots, 0, ops.ctypes.data if ops.size else None, ops.shape[0],
                                     rec.ctypes.data, n_meas, None, None,
                                     None if thr is None else thr.ctypes.data, None if ch is None else ch.ctypes.data,
                                     n_noise, 1, 0, None)

    assert call(4, 3, [[18, 0, -1, -1]]) == N.EOP                       # BARRIER is not a user opcode
    assert call(4, 3, [[99, 0, -1, -1]]) == N.EOP
    assert call(4, 3, [[5, 4, -1, -1]]) == N.EOP                        # qudit out of range
    assert call(4, 3, [[9, 1, 1, -1]]) == N.EOP                         # control == target
    assert call(4, 3, [[9, 1, 7, -1]]) == N.EOP
    assert call(4, 3, [[14, 0, -1, 3]], n_meas=1) == N.EOP              # record slot out of range
    assert call(4, 3, [[17, 0, -1, 0]], n_noise=0) == N.EOP             # noise slot out of range
    assert call(4, 4, [[5, 0, -1, -1]]) == N.EDIM and call(4, 32771, [[5, 0, -1, -1]]) == N.EDIM
    assert call(0, 3, [[5, 0, -1, -1]]) == N.EINVAL and call(4, 3, [[5, 0, -1, -1]], shots=-1) == N.EINVAL
    assert call(4, 3, [[17, 0, -1, 0]], n_noise=1) == N.EINVAL          # Philox mode without noise tables
    assert call(4, 3, [[5, 0, -1, -1]], shots=0) == N.OK                # nothing to do
    assert N.lib().sdimb_strerror(N.EOP).decode() == "Invalid gate value"
