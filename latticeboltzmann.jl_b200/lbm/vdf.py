"""Per-node velocity-distribution-function helpers of the Julia API, host side.

These mirror src/hermite_polynomials.jl and src/velocity_distribution_function/*.jl for
single population vectors (and broadcast over leading grid axes); they build initial
conditions and serve the reference's unit-style calls (`density(q, f)`, `equilibrium(q, rho, u, T)`).
Grid-sized collide/stream/BC/diagnostic work never runs here -- that is the CUDA library.
Population axis is the LAST axis: f[..., i].
"""
import math

import numpy as np

from .quadratures import order

D = 2


def _delta(a, b):
    return 1 if a == b else 0


def hermite(n, xi, q=None):
    """hermite(Val{n}, xi[, q])  (hermite_polynomials.jl:3-82)."""
    cs = 1.0 if q is None else 1 / q.speed_of_sound_squared
    xi = [v for v in xi]
    if n == 0:
        return 1.0
    if n == 1:
        return np.array([xi[0] * 1.0, xi[1] * 1.0])
    H = np.empty((D,) * n)
    for idx in np.ndindex(*H.shape):
        if n == 2:
            a, b = idx
            H[idx] = xi[b] * xi[a] - cs * _delta(a, b)
        elif n == 3:
            a, b, c = idx
            H[idx] = xi[c] * xi[b] * xi[a] - cs * (xi[a] * _delta(b, c) + xi[b] * _delta(a, c) + xi[c] * _delta(a, b))
        elif n == 4:
            a, b, c, e = idx
            H[idx] = (xi[e] * xi[c] * xi[b] * xi[a]
                      - cs * (xi[a] * xi[b] * _delta(c, e) + xi[a] * xi[c] * _delta(b, e) + xi[a] * xi[e] * _delta(b, c)
                              + xi[b] * xi[c] * _delta(a, e) + xi[b] * xi[e] * _delta(a, c) + xi[c] * xi[e] * _delta(a, b))
                      + cs * cs * (_delta(a, b) * _delta(c, e) + _delta(a, c) * _delta(b, e) + _delta(a, e) * _delta(b, c)))
        else:
            raise ValueError(n)
    return H


_HERMITE_TABLES = {}


def _hermite_table(q, n):
    """hermite(Val{n}, c_i, q) for every population; constant per lattice, so built once (the reference recomputes it per
    node per population inside pressure / deviatoric_tensor, moments.jl:27,90)."""
    key = (q.name, float(q.speed_of_sound_squared), n)
    if key not in _HERMITE_TABLES:
        _HERMITE_TABLES[key] = [hermite(n, (int(q.abscissae[0, i]), int(q.abscissae[1, i])), q) for i in range(q.Q)]
    return _HERMITE_TABLES[key]


def density(q, f):
    """density(q, f) = sum(f)  (moments.jl:1-3)."""
    f = np.asarray(f)
    rho = f[..., 0].copy()
    for i in range(1, q.Q):
        rho = rho + f[..., i]
    return rho


def velocity(q, f, rho=None):
    """velocity!(q, f, rho, u) returning u with last axis 2 (moments.jl:5-19)."""
    f = np.asarray(f)
    rho = density(q, f) if rho is None else rho
    u = np.zeros(f.shape[:-1] + (2,))
    for d in range(2):
        acc = np.zeros(f.shape[:-1])
        for i in range(q.Q):
            acc = acc + f[..., i] * float(q.abscissae[d, i])
        u[..., d] = acc / rho
    return u


def velocity_(q, f, rho, u):
    """In-place form: velocity!(q, f, rho, u)."""
    u[...] = velocity(q, f, rho)


def _a_bar_2(q, f):
    H2 = _hermite_table(q, 2)
    f = np.asarray(f)
    a = np.zeros(f.shape[:-1] + (2, 2))
    for i in range(q.Q):
        a = a + f[..., i, None, None] * H2[i]
    return a


def pressure(q, f, rho, u):
    """moments.jl:21-33; 1.0 for D2Q4/D2Q5 (quadratures.jl:127, D2Q5.jl:48)."""
    if q.name in ("D2Q4", "D2Q5"):
        return 1.0 + 0.0 * np.asarray(rho)
    a2 = _a_bar_2(q, f)
    u = np.asarray(u)
    return ((a2[..., 0, 0] + a2[..., 1, 1]) - rho * (u[..., 0] ** 2 + u[..., 1] ** 2 - D)) / D


def temperature(q, f, rho, u):
    """moments.jl:67-74."""
    return pressure(q, f, rho, u) / rho


def momentum_flux(q, f, rho, u):
    """moments.jl:35-56 (returns css * P)."""
    f = np.asarray(f)
    u = np.asarray(u)
    P = np.zeros(f.shape[:-1] + (2, 2))
    for a in range(2):
        for b in range(2):
            for i in range(q.Q):
                P[..., a, b] = P[..., a, b] + f[..., i] * (q.abscissae[a, i] - u[..., a]) * (q.abscissae[b, i] - u[..., b])
    return q.speed_of_sound_squared * P


def equilibrium_coefficient(n, q, rho, u, T):
    """velocity_distribution_function/hermite.jl:37-77, including the Val{4} delta quirk (:69,71)."""
    cs = 1 / q.speed_of_sound_squared
    u = np.asarray(u, dtype=np.float64)
    rho = np.asarray(rho, dtype=np.float64)
    ux = [u[..., 0], u[..., 1]]
    if n == 0:
        return rho
    if n == 1:
        return rho[..., None] * u
    out = np.zeros(rho.shape + (D,) * n)
    for idx in np.ndindex(*((D,) * n)):
        if n == 2:
            a, b = idx
            v = ux[a] * ux[b] + cs * (T - 1) * _delta(a, b)
        elif n == 3:
            a, b, c = idx
            v = ux[a] * ux[b] * ux[c] + cs * (T - 1) * (ux[a] * _delta(b, c) + ux[b] * _delta(a, c) + ux[c] * _delta(a, b))
        elif n == 4:
            a, b, c, d = idx
            v = (ux[a] * ux[b] * ux[c] * ux[d]
                 + cs * (T - 1) * (ux[a] * ux[b] * _delta(c, d) + ux[a] * ux[c] * _delta(b, d) + ux[a] * ux[d] * _delta(b, d)
                                   + ux[b] * ux[c] * _delta(a, d) + ux[b] * ux[d] * _delta(a, d) + ux[c] * ux[d] * _delta(a, b))
                 + cs ** 2 * (T - 1) ** 2 * (_delta(a, b) * _delta(c, d) + _delta(a, c) * _delta(b, d) + _delta(a, d) * _delta(b, c)))
        else:
            raise ValueError(n)
        out[(Ellipsis,) + idx] = rho * v
    return out


def deviatoric_tensor(q, tau, f, rho, u):
    """moments.jl:81-96."""
    a_bar = _a_bar_2(q, f)
    a_eq = equilibrium_coefficient(2, q, rho, u, 1.0)
    s = (a_bar - a_eq) / (1 + 1 / (2 * tau))
    tr = (s[..., 0, 0] + s[..., 1, 1]) / D
    s = s.copy()
    s[..., 0, 0] -= tr
    s[..., 1, 1] -= tr
    return s


def hermite_based_equilibrium(q, rho, u, T):
    """hermite_based_equilibrium!(q, rho, u, T, f)  (hermite.jl:10-33); broadcasts over grids.
    rho, T: (...,)  u: (..., 2)  ->  f: (..., Q)"""
    N = order(q) // 2
    cs = 1 / q.speed_of_sound_squared
    rho = np.asarray(rho, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    T = np.asarray(T, dtype=np.float64) if not np.isscalar(T) else T
    Hs = [_hermite_table(q, n) for n in range(1, N + 1)]
    a_eq = [equilibrium_coefficient(n, q, rho, u, T) for n in range(1, N + 1)]
    f = np.empty(rho.shape + (q.Q,))
    for i in range(q.Q):
        s = 0.0
        for n in range(1, N + 1):
            H = Hs[n - 1][i]
            dot = np.tensordot(a_eq[n - 1], H, axes=(list(range(-n, 0)), list(range(n)))) if n > 1 else (
                a_eq[0][..., 0] * H[0] + a_eq[0][..., 1] * H[1])
            s = s + dot / (math.factorial(n) * cs ** n)
        f[..., i] = q.weights[i] * (rho + s)
    return f


def equilibrium(q, rho, u, T):
    """equilibrium(q, rho, u, T) -> f  (maxwell_boltzmann_equilibrium.jl:1-10; Hermite series)."""
    return hermite_based_equilibrium(q, rho, u, T)


def _pow4(x):
    x2 = x * x
    return x2 * x2


def equilibrium_(q, rho, u, T, f):
    """equilibrium!(q, rho, u, T, f): the truncated polynomial used by collide!
    (maxwell_boltzmann_equilibrium.jl:12-66; velocity_distribution_function/quadratures.jl)."""
    cs = q.speed_of_sound_squared
    u = np.asarray(u, dtype=np.float64)
    ux, uy = u[..., 0], u[..., 1]
    u2 = ux * ux + uy * uy
    eq_order = {"D2Q4": 1, "D2Q5": 1, "D2Q9": 2, "D2Q13": 2, "D2Q17": 3, "D2Q21": 3, "D2Q37": 4}[q.name]
    for i in range(q.Q):
        udx = float(q.abscissae[0, i]) * ux + float(q.abscissae[1, i]) * uy
        poly = 1.0 + cs * udx
        if eq_order >= 2:
            poly = poly + 0.5 * (cs * cs * (udx * udx) + -cs * u2)
        if eq_order >= 3:
            poly = poly + (1 / 6) * (cs * udx * (cs * cs * (udx * udx) - 3 * cs * u2))
        if eq_order >= 4:
            poly = poly + (1 / 24) * (_pow4(cs) * _pow4(udx) - 6 * cs ** 3 * u2 * udx ** 2 + 3 * cs ** 2 * u2 ** 2)
        f[..., i] = rho * q.weights[i] * poly
    return f
