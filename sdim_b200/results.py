"""Measurement record types (reference: sdim/tableau/dataclasses.py:166-210)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

# Structured dtype of the reference's record arrays (sdim/program.py:34-40).
MEASUREMENT_DTYPE = np.dtype([
    ("qudit_index", np.int64),
    ("meas_round", np.int64),
    ("shot", np.int64),
    ("deterministic", np.bool_),
    ("measurement_value", np.int64),
])


@dataclass
class MeasurementResult:
    """Outcome of one measurement; equality compares all four fields, as in the reference."""

    qudit_index: int
    deterministic: bool
    measurement_value: int
    stabilizer_tableau: Optional[Any] = None

    def __str__(self) -> str:
        kind = "deterministic" if self.deterministic else "random"
        return f"Measured qudit ({self.qudit_index}) as ({self.measurement_value}) and was {kind}"

    def __repr__(self) -> str:
        return str(self)

    def get_tableau(self):
        if self.stabilizer_tableau is None:
            raise ValueError("Stabilizer tableau not recorded during measurement")
        return self.stabilizer_tableau
