"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerances (BASELINE.json): forces / single-step positions 1e-12 relative, 100-step RMS 1e-8
RTOL_STEP = 1e-12
RTOL_TRAJ = 1e-8


def load_model(name):
    from spatialpy_b200 import FlatModel
    return FlatModel.load(os.path.join(GOLDEN, f"{name}.model.npz"))


def load_ref(name):
    return np.load(os.path.join(GOLDEN, f"{name}.ref.npz"))


def load_ens(name):
    return np.load(os.path.join(GOLDEN, f"{name}.ens.npz"))


def rel_err(a, b):
    """max |a-b| relative to the largest magnitude of the reference array (norm-wise relative error)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(b))), 1e-300)
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b)
    d = np.where(np.isnan(a) & np.isnan(b), 0.0, d)
    return float(np.max(d)) / scale


def csr_sets(ptr, idx):
    return [np.sort(idx[ptr[i]:ptr[i + 1]]) for i in range(len(ptr) - 1)]


def csr_sorted(ptr, idx, *vals):
    """Reorder every CSR row by neighbour id so two engines' lists can be compared value by value."""
    order = np.empty_like(idx, dtype=np.int64)
    for i in range(len(ptr) - 1):
        b, e = ptr[i], ptr[i + 1]
        order[b:e] = b + np.argsort(idx[b:e], kind="stable")
    return (idx[order],) + tuple(v[order] for v in vals)
