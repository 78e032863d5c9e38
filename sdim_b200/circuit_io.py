"""`.chp` circuit reader / writer (reference: sdim/circuit_io.py:6-139).

Same dialect: free-text header terminated by a line holding only `#`, an
optional `d <dim>` line, then one gate per line — name, one or two integer
qudit indices, optional `key=value` parameters (kept as strings, as the
reference does; Appendix B-7).  The Cirq conversion helpers of the reference
file are not part of the tableau path and are not provided.
"""
from __future__ import annotations

import os
from typing import Optional

from .circuit import Circuit

_PKG_PARENT = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..")


def _resolve(filename: str) -> str:
    # The reference joins every path to `<package dir>/..` (circuit_io.py:20-24); an
    # absolute filename survives os.path.join unchanged, so both spellings work.
    return os.path.join(_PKG_PARENT, filename)


def read_circuit(filename: str) -> Circuit:
    """Parse a `.chp` file into a Circuit.

    Raises ValueError for a malformed `key=value` token or a gate line with
    zero or more than two integer arguments, like the reference
    (circuit_io.py:69,87).
    """
    with open(_resolve(filename), "r") as fh:
        lines = fh.readlines()
    start = next(i for i, line in enumerate(lines) if line.strip() == "#")
    body = lines[start + 1:]

    dimension = 2
    head = body[0].split() if body else []
    if head and head[0].upper() == "D":
        dimension = int(head[1])
        body = body[1:]

    highest = max(int(tok) for line in body for tok in line.split() if tok.isdigit())
    circuit = Circuit(highest + 1, dimension)

    for line in body:
        tokens = line.split()
        if not tokens:
            continue
        name = tokens[0].upper()
        qudits = [int(tok) for tok in tokens[1:] if tok.isdigit()]
        params: Optional[dict] = None
        for tok in tokens[1:]:
            if "=" not in tok:
                continue
            pieces = tok.split("=")
            if len(pieces) != 2:
                raise ValueError("Extra parameter doesn't have the correct format.")
            if params is None:
                params = {}
            params[pieces[0]] = pieces[1]
        extra = params or {}
        if len(qudits) == 1:
            circuit.add_gate(name, qudits[0], **extra)
        elif len(qudits) == 2:
            circuit.add_gate(name, qudits[0], qudits[1], **extra)
        else:
            raise ValueError(f"Unexpected number of arguments for gate {name}")
    return circuit


def write_circuit(circuit: Circuit, output_file: str = "random_circuit.chp",
                  comment: str = "", directory: Optional[str] = None) -> str:
    """Serialise a Circuit to `.chp`; returns the path written.

    Emits the header comment, `#`, `d <dim>`, then each operation under the
    name the user gave it plus its `key=value` params (circuit_io.py:91-139).
    Unlike the reference it does not print a line per gate (circuit_io.py:113).
    """
    header = comment if comment else "Randomly-generated Clifford group quantum circuit"
    out = [header, "#", f"d {circuit.dimension}"]
    for op in circuit.operations:
        text = f"{op.gate_name} {op.qudit_index}"
        if op.target_index is not None:
            text += f" {op.target_index}"
        if op.params is not None:
            for key, value in op.params.items():
                text += f" {key}={value}"
        out.append(text)
    if directory is None:
        directory = os.path.join(_PKG_PARENT, "circuits")
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, output_file)
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    return path
