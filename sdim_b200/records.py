"""Packed measurement records (include/sdimb.h: SDIMB_REC_*).

d <= 127: one byte per record, low 7 bits = value, bit 7 = deterministic flag.
d  > 127: uint16 records (uint16 lanes), low 15 bits = value, bit 15 = deterministic flag.  On the device the 16-bit
buffers are torch.int16 (same bits; PyTorch's uint16 has few operators), on the host they are numpy uint16.
"""
from __future__ import annotations

import numpy as np

WIDE_MIN_DIMENSION = 128      # first dimension that needs two bytes per entry / record


def np_dtype(dimension: int):
    return np.uint16 if dimension >= WIDE_MIN_DIMENSION else np.uint8


def torch_dtype(dimension: int):
    import torch
    return torch.int16 if dimension >= WIDE_MIN_DIMENSION else torch.uint8


def unsigned(arr: np.ndarray) -> np.ndarray:
    """Host view of a record array that came off the device (int16 bits -> uint16)."""
    arr = np.asarray(arr)
    return arr.view(np.uint16) if arr.dtype == np.int16 else arr


def masks(dtype) -> tuple:
    """(deterministic flag, value mask) of a packed record dtype."""
    return (0x8000, 0x7FFF) if np.dtype(dtype).itemsize == 2 else (0x80, 0x7F)


def split(packed: np.ndarray):
    """packed records -> (values, deterministic)."""
    packed = unsigned(packed)
    det, val = masks(packed.dtype)
    return packed & packed.dtype.type(val), (packed & packed.dtype.type(det)) != 0


def to_device(arr, dimension: int):
    """numpy / array-like of replayed outcomes or exponents -> CPU torch tensor of the record dtype."""
    import torch
    a = np.ascontiguousarray(np.asarray(arr), dtype=np_dtype(dimension))
    return torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
