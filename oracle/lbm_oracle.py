"""ORACLE -- CPU restatement of LatticeBoltzmann.jl's collide/stream/BC hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path
(``latticeboltzmann.jl_b200``) never does, and fails loudly without its CUDA
library.

Parity pin: the reference itself (Julia) cannot run in this image, so the
oracle is pinned by outputs the reference left in its repository:
(1) the numeric golden table it prints,
``examples/notebooks/trt_magic_parameter.ipynb:109-176`` (committed under
``tests/golden/trt_magic_parameter.json``; ``tests/test_oracle_golden.py``);
(2) the numbers behind its notebook FIGURES, recovered from the stored SVG to
~1e-5 relative (``tests/golden/extract_notebook_plots.py`` ->
``tests/golden/notebook_figures.json``; ``tests/test_oracle_figures.py``):
shear-wave convergence on all 7 lattices, TGV convergence, 841-step error time
series for six initialisation strategies incl. the Mei et al. iteration,
Couette with MovingWall + BounceBack on D2Q9..D2Q37, the 950-point TRT(tau, tau)
Poiseuille sweep, snapshot stress profiles under the time-dependent force;
(3) every identity the reference's own tests assert (``test/*.jl``;
``tests/test_oracle_identities.py``).
Not covered by any reference number or figure (pinned by restatement and
identities only): MRT with tau != 1, TRT with tau_s != tau_a on the multi-speed
lattices.

Array convention: Julia ``f[x, y, i]`` (column-major, 1-based) is stored here as
``f[i, y, x]`` (numpy C order, 0-based) -- the same bytes.  Every function
keeps the reference's floating-point operation order (left folds, no FMA) so
that the C restatement (``oracle/lbm_oracle.c``, ``-ffp-contract=off``) and
the ``exact`` CUDA kernels can be compared bit-for-bit in Float64.

All citations are file:line under /root/reference.
"""
from fractions import Fraction

import numpy as np

from . import lattices as L  # noqa: F401  (re-export)
from .lattices import Lattice

D = 2  # dimension(q)  (src/quadratures.jl:3)


# --------------------------------------------------------------------------
# Hermite tensors (src/hermite_polynomials.jl:47-82), per population
# --------------------------------------------------------------------------
def _delta(a, b):
    return 1 if a == b else 0  # src/LatticeBoltzmann.jl:8


def hermite(n, xi, q=None):
    """hermite(Val{n}, xi[, q]) for integer/float xi of length 2.

    hermite_polynomials.jl:7-8 (n=0,1), :14-45 (no q), :47-82 (with q:
    delta terms scaled by cs = 1/css).
    """
    cs = 1.0 if q is None else 1 / q.css
    xi = [x for x in xi]
    if n == 0:
        return 1.0
    if n == 1:
        return np.array([xi[0] * 1.0, xi[1] * 1.0])
    if n == 2:
        H = np.empty((D, D))
        for a in range(D):
            for b in range(D):
                H[a, b] = xi[b] * xi[a] - cs * _delta(a, b)
        return H
    if n == 3:
        H = np.empty((D, D, D))
        for a in range(D):
            for b in range(D):
                for c in range(D):
                    H[a, b, c] = xi[c] * xi[b] * xi[a] - cs * (
                        xi[a] * _delta(b, c) + xi[b] * _delta(a, c) + xi[c] * _delta(a, b))
        return H
    if n == 4:
        H = np.empty((D, D, D, D))
        cs2 = cs * cs
        for a in range(D):
            for b in range(D):
                for c in range(D):
                    for e in range(D):
                        H[a, b, c, e] = (
                            xi[e] * xi[c] * xi[b] * xi[a]
                            - cs * (xi[a] * xi[b] * _delta(c, e) + xi[a] * xi[c] * _delta(b, e)
                                    + xi[a] * xi[e] * _delta(b, c) + xi[b] * xi[c] * _delta(a, e)
                                    + xi[b] * xi[e] * _delta(a, c) + xi[c] * xi[e] * _delta(a, b))
                            + cs2 * (_delta(a, b) * _delta(c, e) + _delta(a, c) * _delta(b, e)
                                     + _delta(a, e) * _delta(b, c)))
        return H
    raise ValueError(n)


def hermite_table(q, n):
    """[hermite(Val{n}, c_i, q) for i in 1:Q] (mrt.jl:31)."""
    return [hermite(n, (int(q.cx[i]), int(q.cy[i])), q) for i in range(q.Q)]


# --------------------------------------------------------------------------
# Moments (src/velocity_distribution_function/moments.jl)
# --------------------------------------------------------------------------
def density(q, f):
    """moments.jl:3 -- sum(f) as a left fold over populations."""
    rho = f[0].copy() if isinstance(f[0], np.ndarray) else f[0]
    for i in range(1, q.Q):
        rho = rho + f[i]
    return rho


def velocity(q, f, rho):
    """velocity!(q, f, rho, u) moments.jl:5-19."""
    u = []
    for c in (q.cx, q.cy):
        acc = 0.0 * f[0]
        for i in range(q.Q):
            acc = acc + f[i] * float(c[i])
        u.append(acc / rho)
    return u[0], u[1]


def _a_bar_2(q, f):
    """sum(f[idx] * hermite(Val{2}, c_idx, q)) -- moments.jl:27-28,90-92."""
    H2 = hermite_table(q, 2)
    out = {}
    for a in range(D):
        for b in range(D):
            acc = f[0] * H2[0][a, b]
            for i in range(1, q.Q):
                acc = acc + f[i] * H2[i][a, b]
            out[a, b] = acc
    return out


def pressure(q, f, rho, ux, uy):
    """moments.jl:21-33; D2Q4/D2Q5 overrides return 1.0
    (velocity_distribution_function/quadratures.jl:127, quadratures/D2Q5.jl:48)."""
    if q.name in ("D2Q4", "D2Q5"):
        return 1.0 + 0.0 * rho
    a2 = _a_bar_2(q, f)
    return ((a2[0, 0] + a2[1, 1]) - rho * ((ux * ux + uy * uy) - D)) / D


def temperature(q, f, rho, ux, uy):
    """moments.jl:67-74."""
    return pressure(q, f, rho, ux, uy) / rho


def equilibrium_coefficient(n, q, rho, u, T):
    """velocity_distribution_function/hermite.jl:37-77 (incl. the Val{4} delta typo, :69,71)."""
    cs = 1 / q.css
    if n == 0:
        return rho
    if n == 1:
        return [rho * u[0], rho * u[1]]
    if n == 2:
        lam = cs * (T - 1)
        return {(a, b): rho * (u[a] * u[b] + (lam if a == b else 0.0))
                for a in range(D) for b in range(D)}
    if n == 3:
        out = {}
        for a in range(D):
            for b in range(D):
                for c in range(D):
                    out[a, b, c] = rho * (u[a] * u[b] * u[c] + cs * (T - 1) * (
                        u[a] * _delta(b, c) + u[b] * _delta(a, c) + u[c] * _delta(a, b)))
        return out
    if n == 4:
        out = {}
        for a in range(D):
            for b in range(D):
                for c in range(D):
                    for d in range(D):
                        out[a, b, c, d] = rho * (
                            u[a] * u[b] * u[c] * u[d]
                            + cs * (T - 1) * (
                                u[a] * u[b] * _delta(c, d) + u[a] * u[c] * _delta(b, d)
                                + u[a] * u[d] * _delta(b, d) + u[b] * u[c] * _delta(a, d)
                                + u[b] * u[d] * _delta(a, d) + u[c] * u[d] * _delta(a, b))
                            + cs * cs * (T - 1) * (T - 1) * (
                                _delta(a, b) * _delta(c, d) + _delta(a, c) * _delta(b, d)
                                + _delta(a, d) * _delta(b, c)))
        return out
    raise ValueError(n)


def deviatoric_tensor(q, tau, f, rho, ux, uy):
    """moments.jl:81-96.  tau = css*nu (no +0.5, callers track_hydrodynamic_errors.jl:176-177)."""
    a_bar = _a_bar_2(q, f)
    a_eq = equilibrium_coefficient(2, q, rho, (ux, uy), 1.0)
    den = 1 + 1 / (2 * tau)
    s = {k: (a_bar[k] - a_eq[k]) / den for k in a_bar}
    tr = (s[0, 0] + s[1, 1]) / D
    return {(0, 0): s[0, 0] - tr, (0, 1): s[0, 1], (1, 0): s[1, 0], (1, 1): s[1, 1] - tr}


# --------------------------------------------------------------------------
# Equilibria
# --------------------------------------------------------------------------
def _pow4(x):
    # Julia lowers x^4 to pow(x, 4.0); restated as (x*x)*(x*x) (<= 1 ulp apart).
    x2 = x * x
    return x2 * x2


def equilibrium_collision(q, rho, ux, uy):
    """equilibrium!(q, rho, u, T=1, feq): the truncated polynomial used inside collide!.

    Generic order 2: maxwell_boltzmann_equilibrium.jl:12-66 (D2Q9, D2Q13).
    Order 1: velocity_distribution_function/quadratures.jl:129-159 (D2Q4, D2Q5).
    Order 3: quadratures.jl:3-19, :45-61 (D2Q17, D2Q21).  Order 4: :63-124 (D2Q37).
    """
    cs = q.css
    u2 = (0.0 + ux * ux) + uy * uy
    feq = []
    for i in range(q.Q):
        udx = float(q.cx[i]) * ux + float(q.cy[i]) * uy
        a1 = cs * udx
        poly = 1.0 + a1
        if q.eq_order >= 2:
            a2 = ((cs * cs) * (udx * udx) + 0.0) + (-cs) * u2
            poly = poly + (1 / 2) * a2
        if q.eq_order >= 3:
            a3 = (cs * udx) * ((((cs * cs) * (udx * udx)) - (3 * cs) * u2) + 0.0)
            poly = poly + (1 / 6) * a3
        if q.eq_order >= 4:
            cs3 = cs * cs * cs
            a4 = ((_pow4(cs) * _pow4(udx)) - ((6 * cs3) * u2) * (udx * udx)) + (3 * (cs * cs)) * (u2 * u2)
            poly = poly + (1 / 24) * a4
        feq.append((rho * q.w[i]) * poly)
    return feq


def hermite_based_equilibrium(q, rho, ux, uy, T):
    """hermite_based_equilibrium!(q, rho, u, T, f): hermite.jl:10-33 (used by `equilibrium`,
    maxwell_boltzmann_equilibrium.jl:1-10, i.e. by initialisation)."""
    cs = 1 / q.css
    N = q.N
    Hs = [hermite_table(q, n) for n in range(1, N + 1)]
    a_eq = [equilibrium_coefficient(n, q, rho, (ux, uy), T) for n in range(1, N + 1)]
    fact = [1, 1, 2, 6, 24]
    f = []
    for i in range(q.Q):
        terms = []
        for n in range(1, N + 1):
            H = Hs[n - 1][i]
            a = a_eq[n - 1]
            if n == 1:
                dot = a[0] * H[0] + a[1] * H[1]
            else:
                dot = None
                # column-major order of the tensor entries (first index fastest)
                idxs = list(np.ndindex(*([D] * n)))
                idxs.sort(key=lambda t: t[::-1])
                for t in idxs:
                    term = a[t] * H[t]
                    dot = term if dot is None else dot + term
            terms.append(dot / (fact[n] * cs ** n))
        s = terms[0]
        for t in terms[1:]:
            s = s + t
        f.append(q.w[i] * (rho + s))
    return f


# --------------------------------------------------------------------------
# Collision models (src/collision_models/{srt,trt,mrt}.jl)
# --------------------------------------------------------------------------
class SRT:
    """srt.jl:1-5.  force: None | (Fx, Fy) scalars/arrays[NY,NX] | callable(time)->(Fx,Fy)."""

    def __init__(self, tau, force=None):
        self.tau = float(tau)
        self.force = force


class TRT:
    """trt.jl:1-5 -- 3-arg ctor order (tau_symmetric, tau_asymmetric, force)."""

    def __init__(self, tau_s, tau_a, force=None):
        self.tau_s = float(tau_s)
        self.tau_a = float(tau_a)
        self.force = force


class MRT:
    """mrt.jl:1-34 -- taus per Hermite order (1-based taus[n-1] relaxes a^(n))."""

    def __init__(self, q, taus, force=None):
        if np.isscalar(taus):
            # MRT(q, tau): N = round(Int, order(q) / 2) -- half-to-even, so 4 entries for order 7
            # (mrt.jl:19-22; `force` is dropped there)
            taus = [float(taus)] * round(q.order / 2)
        self.taus = [float(t) for t in taus]
        self.force = force


class IterativeInitializationCollisionModel:
    """collision_models/iterative_initialization.jl:1-41 -- SRT towards an equilibrium whose velocity is pinned to the
    problem's lattice velocity; only the density is taken from f.  `nonlinear_term` is precomputed exactly as the
    reference does (:21-35): w_i rho_0 (a_H_1 + a_H_2 / 2), rho_0 = 1."""
    force = None

    def __init__(self, q, tau, problem):
        self.tau = float(tau)
        X, Y = problem.grid()
        vx, vy = problem.velocity(X, Y)
        ux, uy = problem.u_max * vx, problem.u_max * vy  # lattice_velocity, problems.jl:99-100
        u_squared = ux * ux + uy * uy
        cs = q.css
        self.u0 = (ux, uy)
        self.nonlinear_term = []
        for i in range(q.Q):
            u_dot_xi = float(q.cx[i]) * ux + float(q.cy[i]) * uy
            a_H_1 = cs * u_dot_xi
            a_H_2 = (cs * cs) * (u_dot_xi * u_dot_xi) - cs * u_squared
            self.nonlinear_term.append(q.w[i] * 1.0 * (a_H_1 + a_H_2 / 2))


def _force_at(force, time, shape):
    if force is None:
        return None
    F = force(time) if callable(force) else force
    return F[0], F[1]


def collide(cm, q, f_in, time=0.0):
    """collide!(cm, q, f_in, f_out; time) -> new array f_out."""
    f = [f_in[i] for i in range(q.Q)]
    rho = density(q, f)
    if isinstance(cm, IterativeInitializationCollisionModel):
        # iterative_initialization.jl:42-60
        out = np.empty_like(f_in)
        tau = cm.tau
        for i in range(q.Q):
            feq = q.w[i] * rho + cm.nonlinear_term[i]
            out[i] = (1 - 1 / tau) * f[i] + (1 / tau) * feq
        return out
    ux, uy = velocity(q, f, rho)
    F = _force_at(cm.force, time, rho.shape)
    out = np.empty_like(f_in)
    if isinstance(cm, SRT):
        # srt.jl:18-62
        tau = cm.tau
        if F is not None:
            ux, uy = ux + tau * F[0], uy + tau * F[1]
        feq = equilibrium_collision(q, rho, ux, uy)
        a, b = (1 - 1 / tau), (1 / tau)
        for i in range(q.Q):
            out[i] = a * f[i] + b * feq[i]
        return out
    if isinstance(cm, TRT):
        # trt.jl:42-97 (force shift uses tau_a, :79)
        ts, ta = cm.tau_s, cm.tau_a
        if F is not None:
            ux, uy = ux + ta * F[0], uy + ta * F[1]
        feq = equilibrium_collision(q, rho, ux, uy)
        ws, wa = -(1 / ts), (1 / ta)
        for i in range(q.Q):
            o = q.opp[i]
            feq_s = 0.5 * (feq[i] + feq[o])
            feq_a = 0.5 * (feq[i] - feq[o])
            f_s = 0.5 * (f[i] + f[o])
            f_a = 0.5 * (f[i] - f[o])
            out[i] = f[i] + (ws * (f_s - feq_s) - wa * (f_a - feq_a))
        return out
    if isinstance(cm, MRT):
        # mrt.jl:56-118
        cs = q.css
        N = q.N
        taus = cm.taus
        if F is not None:
            ux, uy = ux + taus[1] * F[0], uy + taus[1] * F[1]  # tau_s[2], mrt.jl:94
        Hs = [hermite_table(q, n) for n in range(1, N + 1)]
        fact = [1, 1, 2, 6, 24]
        a_coll = [None] * (N + 1)
        for n in range(2, N + 1):
            a_eq = equilibrium_coefficient(n, q, rho, (ux, uy), 1.0)
            idxs = list(np.ndindex(*([D] * n)))
            a_f = {}
            for t in idxs:
                acc = f[0] * Hs[n - 1][0][t]
                for i in range(1, q.Q):
                    acc = acc + f[i] * Hs[n - 1][i][t]
                a_f[t] = acc
            tn = taus[n - 1]
            a_coll[n] = {t: (1 - 1 / tn) * a_f[t] + (1 / tn) * a_eq[t] for t in idxs}
        for i in range(q.Q):
            first = (cs * rho) * (ux * Hs[0][i][0] + uy * Hs[0][i][1])
            acc = rho + first
            if N >= 2:
                hs = None
                for n in range(2, N + 1):
                    idxs = list(np.ndindex(*([D] * n)))
                    idxs.sort(key=lambda t: t[::-1])
                    dot = None
                    for t in idxs:
                        term = a_coll[n][t] * Hs[n - 1][i][t]
                        dot = term if dot is None else dot + term
                    term_n = (cs ** n) * dot / fact[n]
                    hs = term_n if hs is None else hs + term_n
                acc = acc + hs
            out[i] = q.w[i] * acc
        return out
    raise TypeError(cm)


# --------------------------------------------------------------------------
# Streaming (src/stream.jl:19-30, :69-74) -- periodic pull with mod1
# --------------------------------------------------------------------------
def stream(q, f):
    out = np.empty_like(f)
    for i in range(q.Q):
        out[i] = np.roll(f[i], shift=(int(q.cy[i]), int(q.cx[i])), axis=(0, 1))
    return out


def stream_push(q, f):
    """stream(q, f, f_new) scatter variant, stream.jl:6-16,44-61 (single wrap only)."""
    Q, NY, NX = f.shape
    out = f.copy()
    for x in range(1, NX + 1):
        for y in range(1, NY + 1):
            for i in range(Q):
                nx_ = x + int(q.cx[i])
                if nx_ > NX:
                    nx_ -= NX
                elif nx_ < 1:
                    nx_ += NX
                ny_ = y + int(q.cy[i])
                if ny_ > NY:
                    ny_ -= NY
                elif ny_ < 1:
                    ny_ += NY
                out[i, ny_ - 1, nx_ - 1] = f[i, y - 1, x - 1]
    return out


# --------------------------------------------------------------------------
# Boundary conditions (src/boundary_conditions/*.jl)
# --------------------------------------------------------------------------
class BounceBack:
    """bounce_back.jl:2-6.  direction in 'N','S','E','W'; xs, ys 1-based inclusive (lo, hi)."""

    def __init__(self, direction, xs, ys):
        self.direction = direction
        self.xs = xs
        self.ys = ys


class MovingWall:
    """moving_wall.jl:5-15 (only a North apply! exists, :17-38; it ignores xs/ys)."""

    def __init__(self, direction, xs, ys, u, rho=1.0, T=1.0):
        self.direction = direction
        self.xs = xs
        self.ys = ys
        self.u = (float(u[0]), float(u[1]))
        self.rho = float(rho)
        self.T = float(T)


def apply_bcs(bcs, q, f_new, f_old, time=0.0):
    """apply!(bcs, q, f_new, f_old; time): in list order (boundary_conditions.jl:6-16). In place on f_new."""
    Q, NY, NX = f_new.shape
    for bc in bcs:
        if isinstance(bc, BounceBack):
            x0, x1 = bc.xs
            y0, y1 = bc.ys
            for i in range(Q):
                o = q.opp[i]
                if bc.direction in ("N", "S"):
                    for y in range(y0, y1 + 1):
                        if bc.direction == "N":
                            if y + int(q.cy[o]) <= NY:  # bounce_back.jl:15
                                continue
                        else:
                            if y + int(q.cy[o]) > 0:  # :31
                                continue
                        f_new[i, y - 1, x0 - 1:x1] = f_old[o, y - 1, x0 - 1:x1]
                else:
                    for x in range(x0, x1 + 1):
                        if bc.direction == "E":
                            if x + int(q.cx[o]) <= NX:  # :48
                                continue
                        else:
                            if x + int(q.cx[o]) > 0:  # :64
                                continue
                        f_new[i, y0 - 1:y1, x - 1] = f_old[o, y0 - 1:y1, x - 1]
        elif isinstance(bc, MovingWall):
            if bc.direction != "N":
                raise NotImplementedError("MovingWall: only North exists (moving_wall.jl:17)")
            cs = q.css
            a1eq = (bc.rho * bc.u[0], bc.rho * bc.u[1])
            for i in range(Q):
                a_1 = q.w[i] * cs * (a1eq[0] * float(q.cx[i]) + a1eq[1] * float(q.cy[i]))
                o = q.opp[i]
                for y in range(1, NY + 1):
                    if y + int(q.cy[o]) <= NY:
                        continue
                    f_new[i, y - 1, :] = f_old[o, y - 1, :] + 2 * a_1
        else:
            raise TypeError(bc)
    return f_new


def step(cm, q, bcs, f_stream, time=0.0):
    """One iteration body of simulate(model, time): lattice_boltzmann_model.jl:65-67.
    Returns (f_stream_new, f_collision)."""
    f_coll = collide(cm, q, f_stream, time)
    f_new = stream(q, f_coll)
    apply_bcs(bcs, q, f_new, f_coll, time)
    return f_new, f_coll


# --------------------------------------------------------------------------
# Problems (src/problems/*.jl) -- analytic fields evaluated on arrays
# --------------------------------------------------------------------------
def _julia_range(a, b, n):
    """range(a, stop=b, length=n): correctly-rounded a + (i-1)(b-a)/(n-1) (Julia's
    twice-precision StepRangeLen), and Float64(range.step)."""
    if n == 1:
        return np.array([a]), 0.0
    fa, fb = Fraction(a), Fraction(b)
    st = (fb - fa) / (n - 1)
    return np.array([float(fa + i * st) for i in range(n)]), float(st)


class Problem:
    """problems.jl:1-128 -- shared unit-scaling helpers."""
    static = False

    def has_external_force(self):
        return False

    def range(self):
        # problems.jl:18-26 (y_range stop uses domain_size[1] -- quirk)
        dx = self.domain_size[0] / self.NX
        dy = self.domain_size[1] / self.NY
        xr, xs = _julia_range(dx / 2, self.domain_size[0] - dx / 2, self.NX)
        yr, ys = _julia_range(dy / 2, self.domain_size[0] - dy / 2, self.NY)
        return xr, yr, xs, ys

    def grid(self):
        xr, yr, _, _ = self.range()
        X, Y = np.meshgrid(xr, yr)  # [NY, NX]
        return X, Y

    def delta_x(self):
        # problems.jl:89-95
        if self.NX > self.NY:
            return self.domain_size[0] * (1 / self.NX)
        return self.domain_size[1] * (1 / self.NY)

    def delta_t(self):
        return self.delta_x() * self.u_max  # problems.jl:85-87

    def viscosity(self):
        return self.nu * self.delta_x() ** 2 / self.delta_t()  # problems.jl:80

    def lattice_viscosity(self):
        return self.nu  # problems.jl:97

    def velocity_gradient(self, x, y, t=0.0):
        z = 0.0 * x
        return z, z, z, z  # (u_x, v_x, u_y, v_y) = [u_x v_x; u_y v_y]  problems.jl:8-16

    def deviatoric(self, q, x, y, t=0.0):
        # problems.jl:30-47:  a = [u_x v_x; u_y v_y]
        a11, a12, a21, a22 = self.velocity_gradient(x, y, t)
        nu = self.viscosity()
        return {(0, 0): -nu * (2 * a11), (0, 1): -nu * (a12 + a21),
                (1, 0): -nu * (a12 + a21), (1, 1): -nu * (2 * a22)}

    def force_field(self, t=0.0):
        """force(problem, x_idx, y_idx, t) over the grid (problems.jl:62-75)."""
        X, Y = self.grid()
        return self.force(X, Y, t)

    def lattice_force_field(self, t=0.0):
        # problems.jl:103-104
        Fx, Fy = self.force_field(t)
        s = self.u_max * self.delta_t()
        return s * Fx, s * Fy

    def boundary_conditions(self):
        return []


class TGV(Problem):
    """second_order_convergence.jl:1-131."""

    def __init__(self, q, tau, scale=2, NX=None, NY=None, u_max=None):
        NX = 16 * scale if NX is None else NX
        NY = NX if NY is None else NY
        u_max = 0.02 / scale if u_max is None else u_max
        self.q = q
        self.rho_0 = 1.0
        self.u_max = 1.0
        self.u_0 = u_max
        self.tau = tau
        self.nu = (tau - 0.5) / q.css
        self.NX, self.NY = NX, NY
        self.static = False
        self.domain_size = (1.0, 1.0)

    def _k(self):
        kx = 2 * np.pi / self.NX
        ky = 2 * np.pi / self.NY
        return kx, ky, 1 / (self.nu * (kx ** 2 + ky ** 2))

    def density(self, q, x, y, t=0.0):
        kx, ky, td = self._k()
        x = x * self.NX
        y = y * self.NY
        return self.rho_0 * (1.0 - q.css * (self.u_0 ** 2 / 4)
                             * ((ky / kx) * np.cos(2 * kx * x) + (kx / ky) * np.cos(2 * ky * y))
                             * np.exp(-2 * t / td))

    def pressure(self, q, x, y, t=0.0):
        kx, ky, td = self._k()
        x = x * self.NX
        y = y * self.NY
        return self.rho_0 - (q.css * 1.0 * (self.u_0 ** 2 / 4)
                             * ((ky / kx) * np.cos(2 * kx * x) + (kx / ky) * np.cos(2 * ky * y))
                             * np.exp(-2 * t / td))

    def velocity(self, x, y, t=0.0):
        kx, ky, td = self._k()
        x = x * self.NX
        y = y * self.NY
        s = self.u_0 * np.exp(-t / td)
        return (s * (-np.sqrt(ky / kx) * np.cos(kx * x) * np.sin(ky * y)),
                s * (np.sqrt(kx / ky) * np.sin(kx * x) * np.cos(ky * y)))

    def velocity_gradient(self, x, y, t=0.0):
        kx, ky, td = self._k()
        x = x * self.NX
        y = y * self.NY
        u_x = np.sqrt(ky * kx) * np.sin(kx * x) * np.sin(ky * y)
        v_y = -np.sqrt(ky * kx) * np.sin(kx * x) * np.sin(ky * y)
        u_y = -np.sqrt(ky ** 3 / kx) * np.cos(kx * x) * np.cos(ky * y)
        v_x = np.sqrt(kx ** 3 / ky) * np.cos(kx * x) * np.cos(ky * y)
        s = np.exp(-t / td) * self.u_0
        return s * u_x, s * v_x, s * u_y, s * v_y

    def viscosity(self):
        return self.nu

    def delta_x(self):
        return 1.0

    def delta_t(self):
        return 1.0

    def decay_time(self):
        return self._k()[2]


class TaylorGreenVortex(Problem):
    """taylor_green_vortex.jl:1-118."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(2 * np.pi, 2 * np.pi),
                 static=True, A=1, B=-1, a=1, b=1):
        NX = 16 * scale if NX is None else NX
        NY = NX if NY is None else NY
        self.rho_0 = 1.0
        self.u_max = 0.01 / scale
        self.nu = nu
        self.NX, self.NY = NX, NY
        self.domain_size = tuple(float(v) for v in domain_size)
        self.static = static
        self.A, self.B, self.a, self.b = float(A), float(B), float(a), float(b)

    def has_external_force(self):
        return self.static

    def decay(self, t):
        return 1.0 if self.static else np.exp(-(self.a ** 2 + self.b ** 2) * self.viscosity() * t)

    def pressure(self, q, x, y, t=0.0):
        P = -(1 / 4) * self.rho_0 * self.decay(t) ** 2 * (
            self.A ** 2 * np.cos(2 * self.a * x) + self.B ** 2 * np.cos(2 * self.b * y))
        return 1.0 + q.css * self.u_max ** 2 * P

    def density(self, q, x, y, t=0.0):
        return self.pressure(q, x, y, t)

    def velocity(self, x, y, t=0.0):
        d = self.decay(t)
        return (d * (self.A * np.cos(self.a * x) * np.sin(self.b * y)),
                d * (self.B * np.sin(self.a * x) * np.cos(self.b * y)))

    def velocity_gradient(self, x, y, t=0.0):
        a, A, b, B = self.a, self.A, self.b, self.B
        u_x = -a * A * np.sin(a * x) * np.sin(b * y)
        v_y = -b * B * np.sin(a * x) * np.sin(b * y)
        u_y = b * A * np.cos(a * x) * np.cos(b * y)
        v_x = a * B * np.cos(a * x) * np.cos(b * y)
        d = self.decay(t)
        return d * u_x, d * v_x, d * u_y, d * v_y

    def force(self, x, y, t=0.0):
        if not self.static:
            return 0.0 * x, 0.0 * x
        ux, uy = self.velocity(x, y, 0.0)
        s = 2 * self.viscosity()
        return s * ux, s * uy


class DecayingShearFlow(Problem):
    """decaying_shear_flow.jl:1-149."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(2 * np.pi, 2 * np.pi),
                 static=True, A=1.0, B=1.0, k_x=1.0, k_y=0.0, u_max=None, convenience=True):
        NX = 8 * scale if NX is None else NX
        NY = NX if NY is None else NY
        if convenience:  # keyword ctor :14-37
            if k_y == 0.0:
                NY = 3
            if k_x == 0.0:
                NX = 3
        self.rho_0 = 1.0
        self.u_max = 0.02 / scale if u_max is None else u_max
        self.nu = nu
        self.NX, self.NY = NX, NY
        self.domain_size = tuple(float(v) for v in domain_size)
        self.static = static
        self.A, self.B, self.k_x, self.k_y = float(A), float(B), float(k_x), float(k_y)

    def has_external_force(self):
        return self.static

    def decay(self, t):
        if self.static:
            return 1.0
        return np.exp(-1.0 * self.k_x ** 2 * self.viscosity() * t)

    def density(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * x

    def pressure(self, q, x, y, t=0.0):
        return 1 + (self.B * 0.025 * q.css * self.u_max ** 2 * self.B
                    * np.sin(self.k_x * (x - self.A * t)) ** 2 * self.decay(t) ** 2)

    def velocity(self, x, y, t=0.0):
        A, B, kx, ky = self.A, self.B, self.k_x, self.k_y
        ux = A * np.cos(ky * y - ky * B * t)
        uy = B * np.cos(kx * x - kx * A * t)
        if self.static:
            return ux, uy
        return (ux * np.exp(-1.0 * ky ** 2 * self.viscosity() * t),
                uy * np.exp(-1.0 * kx ** 2 * self.viscosity() * t))

    def velocity_gradient(self, x, y, t=0.0):
        A, B, kx, ky = self.A, self.B, self.k_x, self.k_y
        u_x = 0.0 * x
        u_y = -A * ky * np.sin(ky * (y - B * t))
        v_x = -B * kx * np.sin(kx * (x - A * t))
        v_y = 0.0 * x
        if not self.static:
            u_y = u_y * np.exp(-1.0 * ky ** 2 * self.viscosity() * t)
            v_x = v_x * np.exp(-1.0 * kx ** 2 * self.viscosity() * t)
        return u_x, v_x, u_y, v_y

    def force(self, x, y, t=0.0):
        if not self.static:
            return 0.0 * x, 0.0 * x
        A, B, kx, ky = self.A, self.B, self.k_x, self.k_y
        nu = self.viscosity()
        return (nu * ky ** 2 * A * np.cos(ky * y - ky * B * t),
                nu * kx ** 2 * B * np.cos(kx * x - kx * A * t))


class PoiseuilleFlow(Problem):
    """poiseuille.jl:1-96."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(1.0, 1.0),
                 u_max=None, k=1.0, G=1.0, convenience=True):
        if convenience:  # :12-26
            NXd = 5 * scale if NX is None else NX
            NY = NXd if NY is None else NY
            NX = 3
            u_max = 0.1 / scale
        self.rho_0 = 1.0
        self.u_max = u_max
        self.nu = nu
        self.NX, self.NY = NX, NY
        self.k = k
        self.domain_size = tuple(float(v) for v in domain_size)
        self.G = G

    def has_external_force(self):
        return True

    def delta_x(self):
        return self.domain_size[1] / self.NY  # :84-86

    def density(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * x

    def pressure(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * x

    def velocity(self, x, y, t=0.0):
        return y * (self.domain_size[1] - y) * (self.G / 2), 0.0 * x

    def velocity_gradient(self, x, y, t=0.0):
        z = 0.0 * x
        return z, z, (self.domain_size[1] - 2 * y) * (self.G / 2), z

    def force_field(self, t=0.0):
        # force(problem, x::Int, y::Int, t) :72-82 -- index based, uniform
        return self.viscosity() * self.G, 0.0

    def boundary_conditions(self):
        return [BounceBack("N", (1, self.NX), (1, self.NY)),
                BounceBack("S", (1, self.NX), (1, self.NY))]


class CouetteFlow(Problem):
    """couette_flow.jl:1-85."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(1.0, 1.0),
                 u_max=None, convenience=True):
        if convenience:  # :10-25
            NXd = 5 * scale if NX is None else NX
            NY = NXd if NY is None else NY
            NX = 1
            u_max = 0.01 / scale
        self.rho_0 = 1.0
        self.u_max = u_max
        self.nu = nu
        self.NX, self.NY = NX, NY
        self.domain_size = tuple(float(v) for v in domain_size)

    def density(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * x

    def pressure(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * x

    def velocity(self, x, y, t=0.0):
        return y + 0.0 * x, 0.0 * x

    def velocity_gradient(self, x, y, t=0.0):
        z = 0.0 * x
        return z, z, 1.0 + z, z

    def boundary_conditions(self):
        return [BounceBack("S", (1, self.NX), (1, self.NY)),
                MovingWall("N", (1, self.NX), (1, self.NY), [self.u_max, 0])]


class LidDrivenCavityFlow(Problem):
    """lid_driven_cavity.jl:1-66."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(1.0, 1.0)):
        NX = 16 * scale if NX is None else NX
        NY = NX if NY is None else NY
        self.rho_0 = 1.0
        self.u_max = 0.01 / scale
        self.nu = nu
        self.NX, self.NY = NX, NY
        self.domain_size = tuple(float(v) for v in domain_size)

    def density(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * x

    def pressure(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * x

    def velocity(self, x, y, t=0.0):
        return 0.0 * x, 0.0 * x

    def boundary_conditions(self):
        return [BounceBack("E", (1, self.NX), (1, self.NY)),
                BounceBack("S", (1, self.NX), (1, self.NY)),
                BounceBack("W", (1, self.NX), (1, self.NY)),
                MovingWall("N", (1, self.NX), (1, self.NY), [self.u_max, 0])]


class LinearizedThermalDiffusion(Problem):
    """linear_hydrodynamics_modes.jl:15-99."""

    def __init__(self, nu, kappa, scale, NY=None):
        NY = 4 * scale if NY is None else NY
        self.rho_0, self.theta_0, self.rho_t, self.u_t = 1.0, 1.0, 0.001, 0.001
        self.nu, self.kappa = nu, kappa
        self.domain_size = (2 * np.pi, 2 * np.pi)
        self.u_max = 0.01 / scale
        self.NX = self.NY = NY

    def delta_x(self):
        return self.domain_size[1] / self.NY

    def heat_diffusion(self):
        return self.kappa * self.delta_x() ** 2 / self.delta_t()

    def density(self, q, x, y, t=0.0):
        return self.rho_0 + self.rho_t * np.sin(y) * np.exp(-self.heat_diffusion() * t) + 0.0 * x

    def pressure(self, q, x, y, t=0.0):
        return self.rho_0 * self.theta_0 + 0.0 * x

    def velocity(self, x, y, t=0.0):
        return 0.0 * x, 0.0 * x


class LinearizedTransverseShearWave(LinearizedThermalDiffusion):
    """linear_hydrodynamics_modes.jl:101-178."""

    def __init__(self, nu, kappa, scale, NY=None):
        super().__init__(nu, kappa, scale, 8 * scale if NY is None else NY)

    def density(self, q, x, y, t=0.0):
        return self.rho_0 + 0.0 * x

    def pressure(self, q, x, y, t=0.0):
        return self.theta_0 + 0.0 * x

    def velocity(self, x, y, t=0.0):
        return self.u_t * np.sin(y) * np.exp(-self.viscosity() * t) + 0.0 * x, 0.0 * x


# --------------------------------------------------------------------------
# Collision-model factories (srt.jl:7-16, trt.jl:8-21,35-40, mrt.jl:36-47)
# --------------------------------------------------------------------------
def _problem_force(problem):
    if not problem.has_external_force():
        return None
    if isinstance(problem, PoiseuilleFlow):
        Fx, Fy = problem.force_field(0.0)
        s = problem.u_max * problem.delta_t()
        return (s * Fx, s * Fy)
    if isinstance(problem, DecayingShearFlow):
        return lambda t: problem.lattice_force_field(t)
    return problem.lattice_force_field(0.0)


def collision_model(kind, q, problem, Lambda=1 / 4):
    """CollisionModel(SRT|TRT|MRT, q, problem)."""
    tau = q.css * problem.lattice_viscosity() + 0.5
    force = _problem_force(problem)
    if kind == "SRT":
        return SRT(tau, force)
    if kind == "TRT":
        return TRT(tau, 0.5 + Lambda / (tau - 0.5), force)
    if kind == "MRT":
        return MRT(q, [tau] * q.order, force)
    raise ValueError(kind)


# --------------------------------------------------------------------------
# Initial conditions (src/initial_conditions{.jl,/*.jl})
# --------------------------------------------------------------------------
def _lattice_fields(q, problem, X, Y):
    rho = problem.density(q, X, Y)                       # lattice_density  problems.jl:98
    vx, vy = problem.velocity(X, Y)
    ux, uy = problem.u_max * vx, problem.u_max * vy       # lattice_velocity :99-100
    T = problem.pressure(q, X, Y) / problem.density(q, X, Y)  # lattice_temperature :105-106
    return rho, ux, uy, T


def initialize(strategy, q, problem):
    """initialize(strategy, q, problem) -> f[Q, NY, NX]  (initial_conditions.jl:7-22)."""
    X, Y = problem.grid()
    one = np.ones_like(X)
    if strategy == "ZeroVelocityInitialCondition":
        # initial_conditions.jl:44-53
        return np.stack([q.w[i] * one for i in range(q.Q)])
    if strategy == "AnalyticalEquilibrium":
        # analytical_equilibrium.jl:8-17 -> problems.jl:121-128
        rho, ux, uy, T = _lattice_fields(q, problem, X, Y)
        return np.stack(hermite_based_equilibrium(q, rho, ux, uy, T))
    if strategy == "ConstantDensity":
        # constant_density.jl:10-20
        _, ux, uy, _ = _lattice_fields(q, problem, X, Y)
        return np.stack(hermite_based_equilibrium(q, one, ux, uy, one))
    if strategy in ("AnalyticalVelocityAndStress", "AnalyticalEquilibriumAndOffEquilibrium"):
        rho, ux, uy, T = _lattice_fields(q, problem, X, Y)
        cs = q.css
        tau = cs * problem.lattice_viscosity()
        H2 = hermite_table(q, 2)
        if strategy == "AnalyticalVelocityAndStress":
            # analytical_velocity_stress.jl:5-31
            f = hermite_based_equilibrium(q, one, ux, uy, one)
            g = [problem.u_max ** 2 * c for c in problem.velocity_gradient(X, Y, 0.0)]
            coef = lambda w: w * (cs * (tau + 0.5) * 1.0 * 1.0) / 2  # noqa: E731
        elif isinstance(problem, TGV):
            # analytical_offequilibrium.jl:50-87
            f = hermite_based_equilibrium(q, rho, ux, uy, T)
            g = list(problem.velocity_gradient(X, Y, 0.0))
            coef = None
        else:
            # analytical_offequilibrium.jl:10-49
            f = hermite_based_equilibrium(q, rho, ux, uy, T)
            g = [problem.u_max ** 2 * c for c in problem.velocity_gradient(X, Y, 0.0)]
            coef = None
        # grad = [u_x v_x; u_y v_y];  S = grad + grad'
        S = {(0, 0): g[0] + g[0], (0, 1): g[1] + g[2], (1, 0): g[2] + g[1], (1, 1): g[3] + g[3]}
        out = []
        for i in range(q.Q):
            dot = (H2[i][0, 0] * S[0, 0] + H2[i][1, 0] * S[1, 0]) + H2[i][0, 1] * S[0, 1]
            dot = dot + H2[i][1, 1] * S[1, 1]
            if strategy == "AnalyticalVelocityAndStress":
                out.append(f[i] + -(coef(q.w[i]) * dot))
            elif isinstance(problem, TGV):
                out.append(f[i] - (q.w[i] * (cs * (tau + 0.5) * rho * 1.0) / 2 * dot))
            else:
                factor = problem.domain_size[0] * problem.domain_size[1]
                out.append(f[i] + (-factor * q.w[i] * 0.5 * ((tau + 0.5) * cs) * dot))
        return np.stack(out)
    raise ValueError(strategy)


# --------------------------------------------------------------------------
# Stop criteria (processing_methods/stopping_criteria/stopping_criteria.jl)
# --------------------------------------------------------------------------
class NoStoppingCriteria:
    def should_stop(self, q, f):
        return False


class MeanVelocityStoppingCriteria:
    """:3-55."""

    def __init__(self, tolerance, old=0.0):
        self.old_mean_velocity = old
        self.tolerance = tolerance

    def should_stop(self, q, f):
        rho = density(q, f)
        ux, _ = velocity(q, f, rho)
        # sequential sum in the reference's x-outer / y-inner order
        u_mean = _fsum(ux) / ux.size
        with np.errstate(divide="ignore", invalid="ignore"):
            converged = abs(np.float64(u_mean) / np.float64(self.old_mean_velocity) - 1)
        if converged < self.tolerance:
            return True
        if np.isnan(u_mean):
            return True
        self.old_mean_velocity = u_mean
        return False


class VelocityConvergenceStoppingCriteria:
    """:57-115 (denominator NOT sqrt'ed, :101)."""

    def __init__(self, tolerance, problem):
        self.old_ux = np.zeros((problem.NY, problem.NX))
        self.old_uy = np.zeros((problem.NY, problem.NX))
        self.tolerance = tolerance

    def should_stop(self, q, f):
        rho = density(q, f)
        ux, uy = velocity(q, f, rho)
        error = _fsum((ux - self.old_ux) ** 2 + (uy - self.old_uy) ** 2)
        old_norm = _fsum(self.old_ux ** 2 + self.old_uy ** 2)
        self.old_ux, self.old_uy = ux.copy(), uy.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            converged = np.sqrt(np.float64(error)) / np.float64(old_norm)
        if converged < self.tolerance:
            return True
        if np.isnan(converged):
            return True
        return False


class DensityConvergence:
    """stopping_criteria/density_convergence.jl:1-17, restated literally: the loop `for x_idx in nx, y_idx in ny` (:9)
    iterates over the two integers themselves, i.e. visits only the node (NX, NY); every other entry of rho / rho_old
    stays 0, so norm(rho - rho_old) is the density change of that one node.  `whole_field=True` gives the criterion
    the code evidently intended (all nodes)."""

    def __init__(self, eps, problem, whole_field=False):
        self.eps = float(eps)
        self.rho_old = np.zeros((problem.NY, problem.NX))
        self.rho = np.zeros((problem.NY, problem.NX))
        self.whole_field = whole_field

    def should_stop(self, q, f):
        rho_now = density(q, [f[i] for i in range(q.Q)])
        if self.whole_field:
            self.rho_old[...] = self.rho
            self.rho[...] = rho_now
        else:
            self.rho_old[-1, -1] = self.rho[-1, -1]
            self.rho[-1, -1] = rho_now[-1, -1]
        d = float(np.sqrt(np.sum((self.rho - self.rho_old) ** 2)))
        return d < self.eps or d > 100.0


class ProcessIterativeInitialization:
    """processing_methods/process_iterative_initialization.jl:1-26 (the inner process method is never called, :18)."""

    def __init__(self, eps, problem, whole_field=False):
        self.stop = DensityConvergence(eps, problem, whole_field)
        self.calls = 0

    def next(self, q, f, t):
        self.calls += 1
        return self.stop.should_stop(q, f)


def initialize_mei_et_al(q, problem, tau=1.0, eps=1e-7, max_steps=10000, whole_field=False):
    """initialize(::IterativeInitializationMeiEtAl, q, problem) (initial_conditions/mei_et_al.jl:11-40):
    f = w everywhere, then simulate(model, 1:10000) with the constant-velocity collision operator until the density
    criterion fires.  Returns (f_stream, number of collide-stream steps taken)."""
    f = np.stack([q.w[i] * np.ones((problem.NY, problem.NX)) for i in range(q.Q)])
    pm = ProcessIterativeInitialization(eps, problem, whole_field)
    model = Model(f, q, IterativeInitializationCollisionModel(q, tau, problem), problem.boundary_conditions(), pm)
    simulate_model(model, range(1, max_steps + 1))
    return model.f_stream, min(pm.calls, max_steps)


def stop_criteria(problem):
    """StopCriteria(problem) :8-15."""
    if isinstance(problem, PoiseuilleFlow):
        return MeanVelocityStoppingCriteria(1e-12)
    if isinstance(problem, CouetteFlow):
        return MeanVelocityStoppingCriteria(1e-7)
    if isinstance(problem, LidDrivenCavityFlow):
        return MeanVelocityStoppingCriteria(1e-5)
    if isinstance(problem, DecayingShearFlow) and problem.static:
        return MeanVelocityStoppingCriteria(1e-8)
    return NoStoppingCriteria()


# --------------------------------------------------------------------------
# Processing methods
# --------------------------------------------------------------------------
def _fsum(a):
    """Sequential (left-fold) sum over the grid in the reference's loop order
    (x outer, y inner); cumsum accumulates strictly left to right."""
    return float(np.cumsum(np.asarray(a, dtype=np.float64).T.ravel())[-1])


def hydrodynamic_fields(q, problem, f):
    """Per-node rho, u, p, sigma as TrackHydrodynamicErrors computes them
    (track_hydrodynamic_errors.jl:134-186), scaled to dimensionless units."""
    fl = [f[i] for i in range(q.Q)]
    rho = density(q, fl)
    ux, uy = velocity(q, fl, rho)
    tau = q.css * problem.lattice_viscosity()
    a_bar = _a_bar_2(q, fl)
    a_eq = equilibrium_coefficient(2, q, rho, (ux, uy), 1.0)
    den = 1 + 1 / (2 * tau)
    a_2 = {k: (a_bar[k] + (1 / (2 * tau)) * a_eq[k]) / den for k in a_bar}
    P00 = a_2[0, 0] - rho * (ux * ux - 1)
    P11 = a_2[1, 1] - rho * (uy * uy - 1)
    p = (P00 + P11) / D
    sig = deviatoric_tensor(q, tau, fl, rho, ux, uy)
    factor = 1 / problem.u_max ** 2
    return dict(rho=rho, ux=ux / problem.u_max, uy=uy / problem.u_max, p=p,
                sxx=sig[0, 0] * factor, sxy=sig[0, 1] * factor,
                syx=sig[1, 0] * factor, syy=sig[1, 1] * factor)


class TrackHydrodynamicErrors:
    """track_hydrodynamic_errors.jl:1-242 (visualisation omitted)."""

    def __init__(self, problem, should_process, n_steps, stop=None):
        self.problem = problem
        self.should_process = should_process
        self.n_steps = n_steps
        self.stop_criteria = stop_criteria(problem) if stop is None else stop
        self.df = []

    def next(self, q, f, t):
        should_stop = False
        if t % 100 == 0:
            if self.stop_criteria.should_stop(q, f):
                should_stop = True
        if (not should_stop) and t != self.n_steps:
            if not self.should_process:
                return False
        pr = self.problem
        NY, NX = f.shape[1:]
        X, Y = pr.grid()
        _, _, xstep, ystep = pr.range()
        time = t * pr.delta_t()
        Delta = ystep * xstep
        if NX == 1:
            Delta = ystep
            if NY == 1:
                Delta = 1.0
        elif NY == 1:
            Delta = xstep
        Delta_ = Delta
        e_rho = pr.density(q, X, Y, time)
        e_ux, e_uy = pr.velocity(X, Y, time)
        e_p = pr.pressure(q, X, Y, time)
        e_s = pr.deviatoric(q, X, Y, time)
        h = hydrodynamic_fields(q, pr, f)
        with np.errstate(divide="ignore", invalid="ignore"):
            rec = dict(
                timestep=t,
                error_rho=np.sqrt(_fsum((h["rho"] - e_rho) ** 2)),
                error_u=np.sqrt(np.float64(_fsum((h["ux"] - e_ux) ** 2 + (h["uy"] - e_uy) ** 2))
                                / np.float64(_fsum(e_ux ** 2 + e_uy ** 2))),
                error_p=np.sqrt(np.float64(_fsum((h["p"] - e_p) ** 2)) / np.float64(_fsum(e_p ** 2))),
                error_sxx=np.sqrt(np.float64(_fsum((e_s[0, 0] - h["sxx"]) ** 2)) / np.float64(_fsum(e_s[0, 0] ** 2))),
                error_sxy=np.sqrt(np.float64(_fsum((e_s[0, 1] - h["sxy"]) ** 2)) / np.float64(_fsum(e_s[0, 1] ** 2))),
                error_syy=np.sqrt(np.float64(_fsum((e_s[1, 1] - h["syy"]) ** 2)) / np.float64(_fsum(e_s[1, 1] ** 2))),
                error_syx=np.sqrt(np.float64(_fsum((e_s[1, 0] - h["syx"]) ** 2)) / np.float64(_fsum(e_s[1, 0] ** 2))),
                mass=Delta_ * _fsum(h["rho"]),
                momentum=Delta_ * _fsum(h["rho"] * (h["ux"] + h["uy"])),
                energy=Delta_ * _fsum(h["rho"] * (h["ux"] ** 2 + h["uy"] ** 2)),
            )
        self.df.append(rec)
        return should_stop


class CompareWithAnalyticalSolution:
    """processing_methods.jl:31-269 (visualisation omitted)."""

    def __init__(self, problem, should_process, n_steps, stop=None):
        self.problem = problem
        self.should_process = should_process
        self.n_steps = n_steps
        self.stop_criteria = stop_criteria(problem) if stop is None else stop
        self.df = []

    def next(self, q, f, t):
        pr = self.problem
        if t % 100 == 0:
            if self.stop_criteria.should_stop(q, f):
                self.process(q, f, t * pr.delta_t())
                return True
        if not self.should_process:
            if t != self.n_steps:
                return False
        self.process(q, f, t * pr.delta_t())
        return False

    def process(self, q, f, time):
        pr = self.problem
        X, Y = pr.grid()
        _, _, xstep, ystep = pr.range()
        fl = [f[i] for i in range(q.Q)]
        rho = density(q, fl)
        ux, uy = velocity(q, fl, rho)
        T = temperature(q, fl, rho, ux, uy)
        p = pressure(q, fl, rho, ux, uy)
        ux, uy = ux / pr.u_max, uy / pr.u_max
        kin = (ux ** 2 + uy ** 2) * rho
        e_rho = pr.density(q, X, Y, time)
        e_p = pr.pressure(q, X, Y, time)
        e_ux, e_uy = pr.velocity(X, Y, time)
        e_T = e_p / e_rho
        e_kin = e_ux ** 2 + e_uy ** 2
        opp = ystep * xstep
        self.df.append(dict(
            density=_fsum(rho), momentum=_fsum((ux + uy) * rho),
            total_energy=_fsum(kin + T), kinetic_energy=_fsum(kin), internal_energy=_fsum(T),
            density_a=_fsum(e_rho), momentum_a=_fsum(e_rho * (e_ux + e_uy)),
            total_energy_a=_fsum(e_kin + e_T), kinetic_energy_a=_fsum(e_kin),
            internal_energy_a=_fsum(e_T),
            error_u=np.sqrt(_fsum(opp * ((ux - e_ux) ** 2 + (uy - e_uy) ** 2))),
            error_p=np.sqrt(_fsum(opp * (p - e_p) ** 2)),
            error_sxx=0.0, error_sxy=0.0, error_syy=0.0, error_syx=0.0))


class TakeSnapshots:
    """take_snapshots.jl:3-29."""

    def __init__(self, problem, every_t):
        self.problem = problem
        self.every_t = every_t
        self.snapshots = []
        self.timesteps = []

    def next(self, q, f, t):
        if isinstance(self.every_t, int):
            if t % self.every_t != 0:
                return False
        elif t not in self.every_t:
            return False
        self.snapshots.append(f.copy())
        self.timesteps.append(t)
        return False


def processing_method(problem, should_process, n_steps, stop=None):
    """ProcessingMethod(problem, should_process, n_steps) processing_methods.jl:10-29."""
    if isinstance(problem, (TaylorGreenVortex, DecayingShearFlow, TGV)):
        return TrackHydrodynamicErrors(problem, should_process, n_steps, stop)
    return CompareWithAnalyticalSolution(problem, should_process, n_steps, stop)


# --------------------------------------------------------------------------
# Driver (src/lattice_boltzmann_model.jl)
# --------------------------------------------------------------------------
class Model:
    def __init__(self, f_stream, q, cm, bcs, pm):
        self.f_stream = f_stream
        self.f_collision = f_stream.copy()
        self.q = q
        self.cm = cm
        self.bcs = bcs
        self.pm = pm


def make_model(problem, q, collision="SRT", strategy="AnalyticalEquilibrium", pm=None):
    """LatticeBoltzmannModel(problem, q; ...) :15-33.  `collision` is a kind string
    ("SRT"/"TRT"/"MRT" -> factory) or an instance (used as is: collision_models.jl:18)."""
    f = initialize(strategy, q, problem)
    cm = collision_model(collision, q, problem) if isinstance(collision, str) else collision
    return Model(f, q, cm, problem.boundary_conditions(), pm)


def simulate_model(model, times):
    """simulate(model, time) :60-77."""
    pr = getattr(model.pm, "problem", None)
    dt = pr.delta_t() if pr is not None else 0.0
    times = list(times)
    for t in times:
        model.f_stream, model.f_collision = step(model.cm, model.q, model.bcs, model.f_stream, t * dt)
        if model.pm is not None and model.pm.next(model.q, model.f_stream, t + 1):
            return model
    if model.pm is not None:
        model.pm.next(model.q, model.f_stream, times[-1] + 1)
    return model


def simulate(problem, q, pm=None, should_process=True, strategy="AnalyticalEquilibrium",
             t_end=1.0, collision="SRT"):
    """simulate(problem, q; ...) :34-59 -- runs 0:n_steps (n_steps+1 iterations)."""
    dt = problem.delta_t()
    n_steps = round(t_end / dt)  # round-half-even == Julia round(Int, x)
    if pm is None:
        pm = processing_method(problem, should_process, n_steps)
    model = make_model(problem, q, collision, strategy, pm)
    return simulate_model(model, range(0, n_steps + 1))
