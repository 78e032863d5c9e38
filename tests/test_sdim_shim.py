"""CPU: code written against the reference's module paths imports and builds circuits through the `sdim` shim."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sdim
from sdim import Circuit, Program, read_circuit, write_circuit, generate_random_clifford_circuit, MeasurementResult
from sdim.program import Program as P2, SimulationOptions
from sdim.circuit import CircuitInstruction
from sdim.gatedata import GateData
from sdim.tableau.dataclasses import MeasurementResult as M2, Tableau
from sdim.tableau.tableau_prime import ExtendedTableau
import sdim_b200
assert Program is P2 is sdim_b200.Program and MeasurementResult is M2 and ExtendedTableau is sdim_b200.ExtendedTableau
c = Circuit(3, 3)
c.add_gate("H", 0); c.add_gate("CNOT", 0, 1); c.add_gate("N1", 2, prob=0.1, noise_channel="d"); c.add_gate("M", [0, 1, 2])
p = Program(c)                       # constructing needs no GPU; simulate() does
assert len(p.circuits[0].operations) == 6 and str(MeasurementResult(1, True, 2)) == "Measured qudit (1) as (2) and was deterministic"
try:
    sdim.WeylTableau(2, 4)
except NotImplementedError as e:
    assert "composite" in str(e)
else:
    raise SystemExit("WeylTableau must be out of scope")
print("SHIM_OK", sdim.__sdim_b200_shim__)
'''


def test_reference_style_imports_resolve_to_sdim_b200():
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "shim"))
    out = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, cwd="/tmp", timeout=120)
    assert out.returncode == 0 and "SHIM_OK True" in out.stdout, out.stderr[-1500:]


def test_install_refuses_to_shadow_a_real_sdim():
    code = ("import sys, types; m = types.ModuleType('sdim'); m.__file__ = '/x/sdim/__init__.py'; sys.modules['sdim'] = m\n"
            "import sdim_b200.compat as c\n"
            "try:\n    c.install_as_sdim()\nexcept RuntimeError as e:\n    print('REFUSED')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert "REFUSED" in out.stdout, out.stderr[-800:]
