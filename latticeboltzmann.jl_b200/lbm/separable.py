"""Separable form of a field sampled on the grid.

`lbm_reduce_errors` takes the analytic fields of a problem as sums of at most two products X(x) Y(y).  The shipped
problems provide that form analytically (`expected_separable`); for any other problem -- and in the Julia binding, which
only has the package's pointwise functions `density(q, problem, x, y, t)` ... to call -- the form is recovered from the
sampled field by a skeleton (cross) decomposition with full pivoting: exact up to round-off for every field of rank <= 2.
"""
import numpy as np


def cross_decompose(E, max_terms=2, rtol=1e-13):
    """E (NX, NY) -> [(a, X, Y), ...] with E ~= sum a X[:, None] Y[None, :], or None when E is not of rank <= max_terms
    (residual above rtol * max |E|)."""
    E = np.asarray(E, dtype=np.float64)
    R = E.copy()
    scale = max(float(np.abs(E).max()), 1e-300)
    terms = []
    for _ in range(max_terms):
        k = int(np.argmax(np.abs(R)))
        i, j = np.unravel_index(k, R.shape)
        pivot = R[i, j]
        if abs(pivot) <= 1e-14 * scale:
            break
        X, Y = R[:, j].copy(), R[i, :].copy()
        terms.append((1.0 / pivot, X, Y))
        R -= np.outer(X, Y) / pivot
    if np.abs(R).max() > rtol * scale:
        return None
    return terms


def separable_from_fields(fields):
    """fields: 8 arrays (NX, NY) -> the `expected` argument of Context.reduce_errors, or None if some field is not
    separable."""
    out = []
    for E in fields:
        t = cross_decompose(E)
        if t is None:
            return None
        out.append((0.0, t))
    return out
