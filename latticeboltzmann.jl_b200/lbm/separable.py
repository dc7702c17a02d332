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


def decompose_function(fn, xs, ys, sample=48, probes=512, rtol=1e-12, seed=0):
    """Separable form of an analytic field WITHOUT evaluating it on the whole grid: fn(X, Y) (vectorised, broadcasting)
    -> (0.0, [(a, X(xs), Y(ys)), ...]) with at most two terms, or None.

    Cross approximation needs only rows and columns of the field: the pivots are found on a coarse sample, every factor
    is one analytic evaluation along a full grid line (len(xs) + len(ys) evaluations per term), and the result is
    verified at random probe points of the full grid.  Cost O(NX + NY) -- what lets a 32768^2 grid be initialised on the
    device from a few tables."""
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    ix = np.unique(np.linspace(0, len(xs) - 1, min(sample, len(xs))).round().astype(int))
    iy = np.unique(np.linspace(0, len(ys) - 1, min(sample, len(ys))).round().astype(int))

    def ev(xv, yv):
        xv, yv = np.asarray(xv, dtype=np.float64), np.asarray(yv, dtype=np.float64)
        return np.asarray(fn(xv, yv), dtype=np.float64) + 0.0 * (xv + yv)

    S = ev(xs[ix][:, None], ys[iy][None, :])
    scale = max(float(np.abs(S).max()), 1e-300)
    terms = []  # (1 / pivot, X over xs, Y over ys)

    def residual(xv, yv, X_at, Y_at):
        """fn minus the terms found so far at points (xv, yv) whose factor values are X_at[k], Y_at[k]"""
        r = ev(xv, yv)
        for (a, _, _), Xa, Ya in zip(terms, X_at, Y_at):
            r = r - a * Xa * Ya
        return r

    R = S.copy()
    for _ in range(2):
        k = int(np.argmax(np.abs(R)))
        i, j = np.unravel_index(k, R.shape)
        if abs(R[i, j]) <= 1e-14 * scale:
            break
        gi, gj = ix[i], iy[j]
        # full column (all x at y = ys[gj]) and row (all y at x = xs[gi]) of the current residual
        X = residual(xs, ys[gj], [t[1] for t in terms], [t[2][gj] for t in terms])
        Y = residual(xs[gi], ys, [t[1][gi] for t in terms], [t[2] for t in terms])
        pivot = X[gi]
        terms.append((1.0 / pivot, X, Y))
        R = R - np.outer(X[ix], Y[iy]) / pivot
    rng = np.random.default_rng(seed)
    px = rng.integers(0, len(xs), probes)
    py = rng.integers(0, len(ys), probes)
    approx = np.zeros(probes)
    for a, X, Y in terms:
        approx = approx + a * X[px] * Y[py]
    if np.abs(approx - ev(xs[px], ys[py])).max() > rtol * scale or np.abs(R).max() > rtol * scale:
        return None
    return (0.0, terms)
