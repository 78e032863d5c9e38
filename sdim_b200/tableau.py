"""Host-side view of one tableau in the reference's own shape.

`ExtendedTableau` here is a plain container for the six int64 arrays the
reference keeps per shot (sdim/tableau/dataclasses.py:24-39,
sdim/tableau/tableau_prime.py:24-26), oriented `[qudit, generator]`, with the
same derived properties and print helpers.  It is what `Program.stabilizer_tableau`
and `MeasurementResult.stabilizer_tableau` hold after a device run, and what a
caller may pass as `Program(circuit, tableau=...)` to start from a state other
than |0...0>.  All gate arithmetic happens on the device store, not here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np


@dataclass
class Tableau:
    num_qudits: int = 1
    dimension: int = 2
    phase_vector: Optional[np.ndarray] = None
    z_block: Optional[np.ndarray] = None
    x_block: Optional[np.ndarray] = None

    def __post_init__(self):
        n = self.num_qudits
        if self.phase_vector is None:
            self.phase_vector = np.zeros(n, dtype=np.int64)
        if self.z_block is None:
            self.z_block = np.eye(n, dtype=np.int64)
        if self.x_block is None:
            self.x_block = np.zeros((n, n), dtype=np.int64)

    @property
    def even(self) -> bool:
        return self.dimension % 2 == 0

    @property
    def order(self) -> int:
        return self.dimension * 2 if self.even else self.dimension

    @property
    def phase_order(self) -> int:
        return 2 if self.even else 1

    @property
    def pauli_size(self) -> int:
        return 2 * self.num_qudits + 1

    @property
    def stab_tableau(self) -> np.ndarray:
        return np.vstack((self.phase_vector, self.z_block, self.x_block))

    def modulo(self) -> None:
        self.z_block %= self.dimension
        self.x_block %= self.dimension
        self.phase_vector %= self.order

    @staticmethod
    def _print_labeled_matrix(label: str, matrix: np.ndarray) -> None:
        print(f"{label}:")
        print(matrix)

    def print_phase_vector(self):
        self._print_labeled_matrix("Phase Vector", self.phase_vector)

    def print_z_block(self):
        self._print_labeled_matrix("Z Block", self.z_block)

    def print_x_block(self):
        self._print_labeled_matrix("X Block", self.x_block)

    def print_tableau(self):
        self.print_phase_vector()
        self.print_z_block()
        self.print_x_block()


@dataclass
class ExtendedTableau(Tableau):
    destab_phase_vector: Optional[np.ndarray] = None
    destab_z_block: Optional[np.ndarray] = None
    destab_x_block: Optional[np.ndarray] = None

    def __post_init__(self):
        super().__post_init__()
        n = self.num_qudits
        if self.destab_z_block is None:
            self.destab_z_block = np.zeros((n, n), dtype=np.int64)
        if self.destab_x_block is None:
            self.destab_x_block = np.eye(n, dtype=np.int64)
        if self.destab_phase_vector is None:
            self.destab_phase_vector = np.zeros(n, dtype=np.int64)

    @property
    def destab_tableau(self) -> np.ndarray:
        return np.vstack((self.destab_phase_vector, self.destab_z_block, self.destab_x_block))

    @property
    def tableau(self) -> np.ndarray:
        return np.hstack((self.stab_tableau, self.destab_tableau))

    def modulo(self) -> None:
        super().modulo()
        self.destab_x_block %= self.dimension
        self.destab_z_block %= self.dimension
        self.destab_phase_vector %= self.order

    def print_destab_phase_vector(self):
        self._print_labeled_matrix("Destabilizer Phase Vector", self.destab_phase_vector)

    def print_destab_z_block(self):
        self._print_labeled_matrix("Destabilizer Z Block", self.destab_z_block)

    def print_destab_x_block(self):
        self._print_labeled_matrix("Destabilizer X Block", self.destab_x_block)

    def print_tableau(self):
        super().print_tableau()
        self.print_destab_phase_vector()
        self.print_destab_z_block()
        self.print_destab_x_block()

    # ---- the reference's per-gate interface (sdim/tableau/tableau_prime.py:97-292) -----------------------
    # Kept for callers that drive a tableau by hand.  Every call packs this tableau into the device store, runs
    # ONE op through the CUDA interpreter and reads the six arrays back (there is no host implementation of the
    # arithmetic); use Program.simulate for anything performance sensitive.
    def _device_op(self, gate_name: str, qudit: int, target: Optional[int] = None, device=None):
        from .circuit import Circuit
        from .program import Program
        circuit = Circuit(self.num_qudits, self.dimension)
        prog = Program(circuit, tableau=self, device=device)
        from .circuit import CircuitInstruction
        result = prog.apply_gate(CircuitInstruction(circuit.gate_data, gate_name, qudit, target))
        t = prog.stabilizer_tableau
        self.x_block, self.z_block, self.phase_vector = t.x_block, t.z_block, t.phase_vector
        self.destab_x_block, self.destab_z_block = t.destab_x_block, t.destab_z_block
        self.destab_phase_vector = t.destab_phase_vector
        return result

    def hadamard(self, qudit_index: int):
        self._device_op("H", qudit_index)

    def hadamard_inv(self, qudit_index: int):
        self._device_op("H_INV", qudit_index)

    def phase(self, qudit_index: int):
        self._device_op("P", qudit_index)

    def phase_inv(self, qudit_index: int):
        self._device_op("P_INV", qudit_index)

    def cnot(self, control: int, target: int):
        self._device_op("CNOT", control, target)

    def cnot_inv(self, control: int, target: int):
        self._device_op("CNOT_INV", control, target)

    def measure(self, qudit_index: int):
        """Z-basis measurement of one qudit (tableau_prime.py:262-292); returns a MeasurementResult."""
        return self._device_op("M", qudit_index)

    # ---- conversion to / from the device store -------------------------------------------------
    @classmethod
    def from_arrays(cls, n: int, d: int, arrays: Dict[str, np.ndarray]) -> "ExtendedTableau":
        return cls(n, d, phase_vector=arrays["p"], z_block=arrays["z"], x_block=arrays["x"],
                   destab_phase_vector=arrays["dp"], destab_z_block=arrays["dz"], destab_x_block=arrays["dx"])

    def pack(self, np_pad: int) -> np.ndarray:
        """Byte image of this tableau in the device layout of include/sdimb.h (one shot): uint8 entries, uint16 for
        d > 127."""
        n, d, o = self.num_qudits, self.dimension, self.order
        W = 2 * np_pad
        img = np.zeros((2 * n + 1, W), dtype=np.uint16 if d > 127 else np.uint8)
        rows = img[: 2 * n].reshape(n, 2, W)
        rows[:, 0, :n] = np.asarray(self.x_block) % d
        rows[:, 1, :n] = np.asarray(self.z_block) % d
        rows[:, 0, np_pad:np_pad + n] = np.asarray(self.destab_x_block) % d
        rows[:, 1, np_pad:np_pad + n] = np.asarray(self.destab_z_block) % d
        img[2 * n, :n] = np.asarray(self.phase_vector) % o
        img[2 * n, np_pad:np_pad + n] = np.asarray(self.destab_phase_vector) % o
        return img.reshape(-1).view(np.uint8)
