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
    assert lib.sdimb_version() == 1


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


def test_error_codes_map_to_value_errors():
    for n, d in ((4, 4), (4, 1), (4, 128), (4, 131), (0, 3)):
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
