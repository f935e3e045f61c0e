"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")

# Parity tolerances (SURVEY.md §8(d) / BASELINE.json north_star), fp32 path vs reference:
TOL_PROTO = 1e-5   # prototypes: max|ours-ref| / max|ref| per vector set
TOL_LOSS = 1e-4    # scalar losses, relative
TOL_GRAD = 1e-4    # input gradients: max|ours-ref| / max|ref| per tensor


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def golden():
    z = np.load(GOLDEN)
    cases = {}
    for key in z.files:
        case, name = key.split("/", 1)
        cases.setdefault(case, {})[name] = z[key]
    return cases


TN_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transnorm_golden.npz")
TOL_TN = 1e-5      # TransNorm outputs, running estimates and gradients: max|ours-ref| / max|ref| per tensor


def transnorm_golden():
    z = np.load(TN_GOLDEN)
    cases = {}
    for key in z.files:
        case, name = key.split("/", 1)
        cases.setdefault(case, {})[name] = z[key]
    return cases
