"""Shared helpers for the parity tests."""
import numpy as np


def compare(a, b):
    """Parity metrics between two float32 arrays: (#differing bit patterns, max |delta|, #texels over 1e-4)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    diff_bits = int((a.view(np.uint32) != b.view(np.uint32)).sum())
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    d = np.where(both_inf, 0.0, np.abs(a.astype(np.float64) - b.astype(np.float64)))
    d = np.nan_to_num(d, nan=np.inf)
    return diff_bits, float(d.max()) if d.size else 0.0, int((d > 1e-4).sum())


def to_bytes(a):
    """pixelFloatToByte (/root/reference/core/pixel-conversion.hpp:8-10) in numpy."""
    a = np.asarray(a, dtype=np.float32)
    c = np.where((a >= 0) & (a <= 1), a, (a > 0).astype(np.float32)).astype(np.float32)
    v = (np.float32(255.5) - np.float32(255.0) * c).astype(np.float32)
    return (~v.astype(np.int32)).astype(np.uint8)
