import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_ready():
    """(ok, why not): a CUDA device is visible and the in-tree library exists."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device visible"
    except Exception as exc:          # pragma: no cover
        return False, f"torch unavailable: {exc}"
    from sdim_b200.build import LIB_PATH
    if not os.path.exists(LIB_PATH):
        return False, f"{LIB_PATH} not built"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: the gpu-marked tests are skipped instead of failing in
    TableauEngine's no-fallback check.  With `-m gpu` on such a box they still FAIL (the driver asked for them)."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason=f"needs the CUDA path: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_random():
    with open(os.path.join(GOLDEN_DIR, "random_circuits.json")) as fh:
        return json.load(fh)["cases"]


@pytest.fixture(scope="session")
def golden_shipped():
    with open(os.path.join(GOLDEN_DIR, "shipped_circuits.json")) as fh:
        return json.load(fh)["circuits"]


@pytest.fixture(scope="session")
def golden_large_primes():
    with open(os.path.join(GOLDEN_DIR, "large_primes.json")) as fh:
        return json.load(fh)["cases"]


@pytest.fixture(scope="session")
def golden_wide_primes():
    """tests/golden/wide_primes.json (oracle/make_golden.py --wide): d in {131, 251, 257, 1031, 32749}."""
    with open(os.path.join(GOLDEN_DIR, "wide_primes.json")) as fh:
        return json.load(fh)["cases"]


def _load_npz_cases(filename):
    import numpy as np
    out = []
    with np.load(os.path.join(GOLDEN_DIR, filename)) as z:
        for name in z["names"]:
            name = str(name)
            n, d = (int(v) for v in z[name + "/nd"])
            out.append({"name": name, "n": n, "d": d, "ops": z[name + "/ops"], "noise_ab": z[name + "/noise_ab"],
                        "records": z[name + "/records"],
                        "final": {k: z[name + "/" + k].astype(np.int64) for k in ("x", "z", "p", "dx", "dz", "dp")}})
    return out


@pytest.fixture(scope="session")
def golden_lanes_sizes():
    """tests/golden/lanes_sizes.npz (oracle/make_golden.py --lanes): one shot of the UNMODIFIED reference per case for the
    uint8-lane kernels at multi-word sizes (d = 5, 7, 11; n = 97 ... 256), same format as golden_config_sizes."""
    return _load_npz_cases("lanes_sizes.npz")


@pytest.fixture(scope="session")
def golden_config_sizes():
    """tests/golden/config_sizes.npz (oracle/make_golden.py --configs): one shot of the UNMODIFIED reference per
    BASELINE.json config size.  -> list of dicts (name, n, d, ops, noise_ab, records, final)."""
    import numpy as np
    out = []
    with np.load(os.path.join(GOLDEN_DIR, "config_sizes.npz")) as z:
        for name in z["names"]:
            name = str(name)
            n, d = (int(v) for v in z[name + "/nd"])
            out.append({"name": name, "n": n, "d": d, "ops": z[name + "/ops"], "noise_ab": z[name + "/noise_ab"],
                        "records": z[name + "/records"],
                        "final": {k: z[name + "/" + k].astype(np.int64) for k in ("x", "z", "p", "dx", "dz", "dp")}})
    return out
