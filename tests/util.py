"""Helpers shared by the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
KIND_NAME = {0: "normal", 1: "uniform", 2: "exponential"}


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_params(g):
    """Rebuild the (kind, p1, p2, dmax) list a golden file was generated with."""
    cplx = bool(int(g["is_complex"])) if "is_complex" in g.files else None
    out = []
    for k, a, b, d in zip(g["kinds"], g["p1"], g["p2"], g["dmax"]):
        if cplx is False:
            a, b = float(a.real), float(b.real)
        out.append((KIND_NAME[int(k)], a, b, int(d)))
    return out


def bits(a):
    return np.ascontiguousarray(a).view(np.float64).view(np.int64)


def same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def normwise(a, b):
    """max|a-b| / max|b|  — the parity measure of SURVEY.md §7 (normwise relative error)."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))
