"""`sdim` shim: put this directory's parent (`<repo>/shim`) on PYTHONPATH and code written against events555/sdim
(`from sdim import Circuit, Program, read_circuit`, `from sdim.tableau.dataclasses import MeasurementResult`) runs on
sdim_b200's CUDA path.  See sdim_b200/compat.py."""
import os
import sys

_repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _repo not in sys.path:
    sys.path.insert(0, _repo)
__sdim_b200_shim__ = True

from sdim_b200.compat import install_as_sdim  # noqa: E402

install_as_sdim()
