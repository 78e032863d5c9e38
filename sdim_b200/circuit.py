"""Circuit model: the user-facing container the simulator consumes.

Behavioural mirror of the reference `Circuit` / `CircuitInstruction`
(reference: sdim/circuit.py:5-242): same constructor, same `add_gate`
broadcasting rules and error messages, same operator overloads including
their quirks (`*` mutates and returns self, two-qudit `add_gate` drops
kwargs, `from_operation_list` drops params — SURVEY Appendix B-9), so that a
program written against the reference builds the identical operation list.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Union

from .gatedata import GateData, shared_gate_data


@dataclass
class CircuitInstruction:
    """One gate application (reference: sdim/circuit.py:5-40).

    `gate_name` keeps the spelling the user gave (upper-cased, possibly an
    alias); `name` is the canonical name and `gate_id` the integer opcode.
    """

    gate_data: GateData
    gate_name: str
    qudit_index: int
    target_index: Optional[int] = None
    gate_id: Optional[int] = None
    name: Optional[str] = None
    params: Optional[dict] = None

    def __post_init__(self):
        self.gate_id = self.gate_data.get_gate_id(self.gate_name)
        if self.gate_id is None:
            raise ValueError(f"Gate {self.gate_name} not found")
        self.name = self.gate_data.get_gate_name(self.gate_id)

    def __str__(self) -> str:
        return f"{self.gate_id} {self.qudit_index} {self.target_index}"


@dataclass
class Circuit:
    """A sequence of gate applications on `num_qudits` qudits of dimension `dimension`."""

    num_qudits: int
    dimension: int = 2
    operations: Optional[list] = None
    gate_data: Optional[GateData] = None

    def __post_init__(self):
        if self.num_qudits < 1:
            raise ValueError("Number of qudits must be greater than 0")
        if self.dimension < 2:
            raise ValueError("Dimension must be greater than 1")
        self.operations = self.operations or []
        self.gate_data = self.gate_data or shared_gate_data(self.dimension)

    def add_gate(self, gate_name: str, control: Union[int, List[int]],
                 target: Union[int, List[int], None] = None, **kwargs) -> None:
        """Append gate(s); lists broadcast as in the reference (sdim/circuit.py:76-126).

        One control with k targets, k controls with one target, or equal-length
        lists zipped pairwise; anything else raises ValueError.  Keyword
        arguments become the instruction's `params` for single-qudit gates
        (after merging the gate's defaults); two-qudit gates carry no params.
        """
        controls = [control] if isinstance(control, int) else control
        targets = [target] if isinstance(target, int) else target
        upper = gate_name.upper()
        gd = self.gate_data
        gate = gd.gateMap.get(gd.aliasMap.get(upper, upper))
        if gate is not None and gate.defaults:
            for key, value in gate.defaults.items():
                kwargs.setdefault(key, value)
        if targets is None:
            for c in controls:
                self.operations.append(CircuitInstruction(gd, upper, c, None, params=kwargs))
            return
        if len(controls) == 1:
            pairs = [(controls[0], t) for t in targets]
        elif len(targets) == 1:
            pairs = [(c, targets[0]) for c in controls]
        elif len(controls) == len(targets):
            pairs = list(zip(controls, targets))
        else:
            raise ValueError("Invalid combination of control and target qubits")
        for c, t in pairs:
            self.operations.append(CircuitInstruction(gd, upper, c, t))

    def __mul__(self, repetitions: int) -> "Circuit":
        # In-place, like the reference (sdim/circuit.py:128-141).
        base = list(self.operations)
        for _ in range(repetitions - 1):
            self.operations.extend(base)
        return self

    def __imul__(self, repetitions: int) -> "Circuit":
        return self.__mul__(repetitions)

    def __add__(self, other: "Circuit") -> "Circuit":
        if self.dimension != other.dimension:
            raise ValueError("Cannot add circuits with different dimensions")
        out = Circuit(max(self.num_qudits, other.num_qudits), self.dimension)
        out.operations = list(self.operations) + list(other.operations)
        return out

    def __iadd__(self, other: "Circuit") -> "Circuit":
        if self.dimension != other.dimension:
            raise ValueError("Cannot add circuits with different dimensions")
        self.num_qudits = max(self.num_qudits, other.num_qudits)
        self.operations.extend(other.operations)
        return self

    def __str__(self) -> str:
        return "\n".join(str(op) for op in self.operations)

    def print_gateData(self) -> None:
        print(self.gate_data)

    @classmethod
    def from_operation_list(cls, operation_list, num_qudits: int, dimension: int) -> "Circuit":
        """Build from `(name, [qudits])` tuples or CircuitInstructions (sdim/circuit.py:212-242)."""
        circuit = cls(num_qudits, dimension)
        for op in operation_list:
            if isinstance(op, tuple):
                name, qudits = op[0], op[1]
                if len(qudits) == 1:
                    circuit.add_gate(name, qudits[0])
                elif len(qudits) == 2:
                    circuit.add_gate(name, qudits[0], qudits[1])
                else:
                    raise ValueError(f"Unsupported number of qudits for gate {name}")
            elif isinstance(op, CircuitInstruction):
                circuit.add_gate(op.gate_name, op.qudit_index, op.target_index)
            else:
                raise ValueError(f"Unsupported operation type: {type(op)}")
        return circuit
