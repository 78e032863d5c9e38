"""`import sdim` compatibility: expose sdim_b200 under the reference's module paths.

The reference's public names live in `sdim` and a few submodules (sdim/__init__.py:67-73; user code and the
reference's own tests also import `sdim.program`, `sdim.circuit`, `sdim.circuit_io`, `sdim.gatedata`,
`sdim.random_circuit`, `sdim.tableau.dataclasses`, `sdim.tableau.tableau_prime`).  `install_as_sdim()` registers
aliases of the sdim_b200 modules under those names, so that code written against the reference runs on the GPU path
without edits.  Out-of-scope names (WeylTableau, the Cirq helpers: composite dimensions / unitary.py) raise on use.

Two ways in:   import sdim_b200.compat; sdim_b200.compat.install_as_sdim()
        or:    PYTHONPATH=<repo>/shim   (a one-file `sdim` package that does the same at import)
It refuses to replace a real `sdim` that is already imported.
"""
from __future__ import annotations

import sys
import types


def _out_of_scope(name: str, why: str):
    def raiser(*_a, **_k):
        raise NotImplementedError(f"sdim.{name} is outside the prime-dimension tableau path sdim_b200 implements ({why})")
    raiser.__name__ = name
    return raiser


def install_as_sdim(force: bool = False) -> types.ModuleType:
    import sdim_b200
    from . import circuit, circuit_io, gatedata, program, random_circuit, results, tableau
    have = sys.modules.get("sdim")
    if have is not None and not getattr(have, "__sdim_b200_shim__", False) and not force:
        raise RuntimeError(f"a different `sdim` is already imported from {getattr(have, '__file__', '?')}")
    root = have if have is not None and getattr(have, "__sdim_b200_shim__", False) else types.ModuleType("sdim")
    root.__sdim_b200_shim__ = True
    root.__path__ = []                                       # a package: submodule imports resolve through sys.modules
    root.__doc__ = "sdim_b200 under the reference's name (see sdim_b200/compat.py)"
    for name in sdim_b200.__all__:
        setattr(root, name, getattr(sdim_b200, name))
    root.WeylTableau = _out_of_scope("WeylTableau", "composite dimensions")
    root.circuit_to_cirq_circuit = _out_of_scope("circuit_to_cirq_circuit", "Cirq bridge")
    root.cirq_statevector_from_circuit = _out_of_scope("cirq_statevector_from_circuit", "Cirq bridge")
    tab_pkg = types.ModuleType("sdim.tableau")
    tab_pkg.__path__ = []
    dataclasses_mod = types.ModuleType("sdim.tableau.dataclasses")
    for name in ("MeasurementResult", "MEASUREMENT_DTYPE"):
        setattr(dataclasses_mod, name, getattr(results, name))
    dataclasses_mod.Tableau = tableau.Tableau
    prime_mod = types.ModuleType("sdim.tableau.tableau_prime")
    prime_mod.ExtendedTableau = tableau.ExtendedTableau
    tab_pkg.dataclasses, tab_pkg.tableau_prime = dataclasses_mod, prime_mod
    tab_pkg.ExtendedTableau, tab_pkg.Tableau, tab_pkg.MeasurementResult = \
        tableau.ExtendedTableau, tableau.Tableau, results.MeasurementResult
    mods = {"sdim": root, "sdim.circuit": circuit, "sdim.circuit_io": circuit_io, "sdim.gatedata": gatedata,
            "sdim.program": program, "sdim.random_circuit": random_circuit, "sdim.tableau": tab_pkg,
            "sdim.tableau.dataclasses": dataclasses_mod, "sdim.tableau.tableau_prime": prime_mod}
    for full, mod in mods.items():
        sys.modules[full] = mod
        if full != "sdim":
            parent, _, leaf = full.rpartition(".")
            setattr(sys.modules[parent], leaf, mod)
    return root
