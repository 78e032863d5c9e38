"""sdim_b200 — B200-native implementation of sdim's prime-dimension stabilizer-tableau path.

Drop-in for that path of events555/sdim: same `Circuit` / `Program.simulate` / `.chp` /
`MeasurementResult` surface (reference: sdim/__init__.py:67-73); the tableau work runs in
hand-written CUDA kernels for sm_100a behind the C ABI of include/sdimb.h.
"""
from .circuit import Circuit, CircuitInstruction
from .circuit_io import read_circuit, write_circuit
from .gatedata import Gate, GateData
from .program import Program, RecordTable, SimulationOptions
from .random_circuit import generate_and_write_random_circuit, generate_random_clifford_circuit
from .results import MEASUREMENT_DTYPE, MeasurementResult
from .tableau import ExtendedTableau, Tableau

__all__ = [
    "Circuit", "CircuitInstruction", "read_circuit", "write_circuit", "Gate", "GateData", "Program",
    "RecordTable", "SimulationOptions", "generate_random_clifford_circuit",
    "generate_and_write_random_circuit", "MeasurementResult", "MEASUREMENT_DTYPE", "ExtendedTableau", "Tableau",
]
