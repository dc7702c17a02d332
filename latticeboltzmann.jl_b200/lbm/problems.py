"""Problem definitions of the Julia API (src/problems/*.jl), host side.

Analytic fields are evaluated with numpy on whole grids: x, y may be scalars or arrays
(grids have shape (NX, NY): `X[x, y] = x_range[x]`, `Y[x, y] = y_range[y]`).  Vector fields
return `(ux, uy)`; tensors return `((a11, a12), (a21, a22))` with the reference's layout
`[u_x v_x; u_y v_y]`.
"""
from fractions import Fraction

import numpy as np

from .boundary_conditions import BounceBack, East, MovingWall, North, South, West


def _julia_range(a, b, n):
    """range(a, stop=b, length=n): Julia's twice-precision range gives the correctly rounded
    a + (i-1)(b-a)/(n-1); returns (values, Float64(step))."""
    if n == 1:
        return np.array([float(a)]), 0.0
    fa, fb = Fraction(a), Fraction(b)
    st = (fb - fa) / (n - 1)
    return np.array([float(fa + i * st) for i in range(n)]), float(st)


class FluidFlowProblem:
    """problems/problems.jl:1-128."""
    static = False

    def has_external_force(self):
        return False

    def range(self):
        # problems.jl:18-26; y_range's stop uses domain_size[1] (sic)
        key = (self.NX, self.NY, tuple(self.domain_size))
        memo = self.__dict__.get("_range_memo")
        if memo is None or memo[0] != key:  # (the grid of a problem does not change; diagnostics ask for it every call)
            dx = self.domain_size[0] / self.NX
            dy = self.domain_size[1] / self.NY
            xr, xstep = _julia_range(dx / 2, self.domain_size[0] - dx / 2, self.NX)
            yr, ystep = _julia_range(dy / 2, self.domain_size[0] - dy / 2, self.NY)
            xr.setflags(write=False)
            yr.setflags(write=False)
            memo = self.__dict__["_range_memo"] = (key, xr, yr, xstep, ystep)
        _, xr, yr, self._xstep, self._ystep = memo
        return xr, yr

    def range_steps(self):
        self.range()
        return self._xstep, self._ystep

    def grid(self, y0=0, ny=None):
        """(X, Y) of shape (NX, ny) for rows y0 .. y0+ny-1 (slab support)."""
        xr, yr = self.range()
        ny = self.NY - y0 if ny is None else ny
        return np.meshgrid(xr, yr[y0:y0 + ny], indexing="ij")

    def delta_x(self):
        if self.NX > self.NY:
            return self.domain_size[0] * (1 / self.NX)
        return self.domain_size[1] * (1 / self.NY)

    def delta_t(self):
        return self.delta_x() * self.u_max

    def viscosity(self):
        return self.nu * self.delta_x() ** 2 / self.delta_t()

    def lattice_viscosity(self):
        return self.nu

    def velocity_gradient(self, x, y, t=0.0):
        z = 0.0 * np.asarray(x, dtype=np.float64)
        return ((z, z), (z, z))

    def deviatoric_tensor(self, q, x, y, t=0.0):
        """deviatoric_tensor(q, problem, x, y, t): problems.jl:30-47."""
        (a11, a12), (a21, a22) = self.velocity_gradient(x, y, t)
        nu = self.viscosity()
        return ((-nu * (2 * a11), -nu * (a12 + a21)), (-nu * (a12 + a21), -nu * (2 * a22)))

    def boundary_conditions(self):
        return []

    def expected_separable(self, q, t, y0=0, ny=None):
        """The analytic fields rho, u_x, u_y, p, sigma_xx, sigma_xy, sigma_yx, sigma_yy at time t as
        sums of <= 2 products X(x) Y(y): [(c0, [(a, X | None, Y | None), ...]), ...] -- the input of the
        on-device error norms (lbm_reduce_errors).  None: not available for this problem."""
        return None

    def _xy(self, y0=0, ny=None):
        xr, yr = self.range()
        ny = self.NY - y0 if ny is None else ny
        return xr, yr[y0:y0 + ny]

    # --- lattice units (problems.jl:97-106) ------------------------------------------------
    def lattice_density(self, q, x, y, t=0.0):
        return self.density(q, x, y, t)

    def lattice_velocity(self, q, x, y, t=0.0):
        ux, uy = self.velocity(x, y, t)
        return self.u_max * ux, self.u_max * uy

    def lattice_pressure(self, q, x, y, t=0.0):
        return self.u_max ** 2 * self.pressure(q, x, y, t)

    def lattice_temperature(self, q, x, y, t=0.0):
        return self.pressure(q, x, y) / self.density(q, x, y)

    def force_on_grid(self, t=0.0, y0=0, ny=None):
        """force(problem, x_idx, y_idx, t) for every node (problems.jl:62-75)."""
        X, Y = self.grid(y0, ny)
        return self.force(X, Y, t)

    # --- dimensionless (problems.jl:108-119) -----------------------------------------------
    def dimensionless_velocity(self, u):
        return u / self.u_max

    def dimensionless_stress(self, s):
        return s * (1 / self.u_max ** 2)


def has_external_force(p):
    return p.has_external_force()


# Julia's generic functions over problems (problems.jl:97-119), as free functions with the reference's argument order
def lattice_density(q, problem, x, y, t=0.0):
    return problem.lattice_density(q, x, y, t)


def lattice_velocity(q, problem, x, y, t=0.0):
    return problem.lattice_velocity(q, x, y, t)


def lattice_pressure(q, problem, x, y, t=0.0):
    return problem.lattice_pressure(q, x, y, t)


def lattice_temperature(q, problem, x, y, t=0.0):
    return problem.lattice_temperature(q, x, y, t)


def dimensionless_viscosity(problem):
    return problem.nu * problem.delta_x() ** 2 / problem.delta_t()


def dimensionless_density(problem, rho):
    return rho


def dimensionless_velocity(problem, u):
    return u / problem.u_max


def dimensionless_pressure(q, problem, p):
    return p


def dimensionless_temperature(q, problem, T):
    return T


def dimensionless_force(problem, F):
    return F / (problem.u_max * problem.delta_t())


def dimensionless_stress(problem, sigma):
    return sigma * (1 / problem.u_max ** 2)


def force(problem, x, y, t=0.0):
    """force(problem, x, y, t): the problem's own method, [0, 0] for the unforced ones (problems.jl:62-75)."""
    if hasattr(problem, "force"):
        return problem.force(x, y, t)
    z = 0.0 * (np.asarray(x, dtype=np.float64) + y)
    return z, z


def decay(problem, x, y, t):
    """decay(problem, x, y, t) (decaying_shear_flow.jl:117-129, taylor_green_vortex.jl)."""
    return problem.decay(x, y, t)


def range_(problem):
    """range(problem) -> (x_range, y_range), cell-centred (problems.jl:18-26)."""
    return problem.range()


def delta_t(p):
    return p.delta_t()


def delta_x(p):
    return p.delta_x()


def viscosity(p):
    return p.viscosity()


def lattice_viscosity(p):
    return p.lattice_viscosity()


def boundary_conditions(p):
    return p.boundary_conditions()


def lattice_force(problem, x_idx, y_idx, t=0.0):
    """lattice_force(problem, x_idx, y_idx, t) with 1-based indices (problems.jl:103-104)."""
    s = problem.u_max * problem.delta_t()
    if hasattr(problem, "force_idx"):
        F = problem.force_idx(x_idx, y_idx, t)
    else:
        xr, yr = problem.range()
        F = problem.force(xr[x_idx - 1], yr[y_idx - 1], t)
    return np.array([s * F[0], s * F[1]])


class TGV(FluidFlowProblem):
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
        self.nu = (tau - 0.5) / q.speed_of_sound_squared
        self.NX, self.NY = int(NX), int(NY)
        self.static = False
        self.domain_size = (1.0, 1.0)

    def _k(self):
        kx, ky = 2 * np.pi / self.NX, 2 * np.pi / self.NY
        return kx, ky, 1 / (self.nu * (kx ** 2 + ky ** 2))

    def density(self, q, x, y, t=0.0):
        kx, ky, td = self._k()
        x, y = x * self.NX, y * self.NY
        return self.rho_0 * (1.0 - q.speed_of_sound_squared * (self.u_0 ** 2 / 4)
                             * ((ky / kx) * np.cos(2 * kx * x) + (kx / ky) * np.cos(2 * ky * y)) * np.exp(-2 * t / td))

    def pressure(self, q, x, y, t=0.0):
        kx, ky, td = self._k()
        x, y = x * self.NX, y * self.NY
        return self.rho_0 - (q.speed_of_sound_squared * 1.0 * (self.u_0 ** 2 / 4)
                             * ((ky / kx) * np.cos(2 * kx * x) + (kx / ky) * np.cos(2 * ky * y)) * np.exp(-2 * t / td))

    def velocity(self, x, y, t=0.0):
        kx, ky, td = self._k()
        x, y = x * self.NX, y * self.NY
        s = self.u_0 * np.exp(-t / td)
        return (s * (-np.sqrt(ky / kx) * np.cos(kx * x) * np.sin(ky * y)),
                s * (np.sqrt(kx / ky) * np.sin(kx * x) * np.cos(ky * y)))

    def velocity_gradient(self, x, y, t=0.0):
        kx, ky, td = self._k()
        x, y = x * self.NX, y * self.NY
        u_x = np.sqrt(ky * kx) * np.sin(kx * x) * np.sin(ky * y)
        v_y = -np.sqrt(ky * kx) * np.sin(kx * x) * np.sin(ky * y)
        u_y = -np.sqrt(ky ** 3 / kx) * np.cos(kx * x) * np.cos(ky * y)
        v_x = np.sqrt(kx ** 3 / ky) * np.cos(kx * x) * np.cos(ky * y)
        s = np.exp(-t / td) * self.u_0
        return ((s * u_x, s * v_x), (s * u_y, s * v_y))

    def force(self, x, y, t=0.0):
        z = 0.0 * np.asarray(x, dtype=np.float64)
        return z, z

    def expected_separable(self, q, t, y0=0, ny=None):
        kx, ky, td = self._k()
        cs, u0, nu = q.speed_of_sound_squared, self.u_0, self.nu
        D1, D2 = np.exp(-t / td), np.exp(-2 * t / td)
        key = (self.NX, self.NY, y0, ny)
        memo = self.__dict__.get("_sep_memo")
        if memo is None or memo[0] != key:  # the spatial factors do not depend on t
            xr, yr = self._xy(y0, ny)
            X, Y = xr * self.NX, yr * self.NY
            memo = self.__dict__["_sep_memo"] = (key, np.cos(kx * X), np.sin(kx * X), np.cos(ky * Y), np.sin(ky * Y),
                                                  np.cos(2 * kx * X), np.cos(2 * ky * Y))
        _, cx, sx, cy, sy, c2x, c2y = memo
        K = cs * (u0 ** 2 / 4) * D2
        s = D1 * u0
        return [
            (self.rho_0, [(-self.rho_0 * K * (ky / kx), c2x, None), (-self.rho_0 * K * (kx / ky), None, c2y)]),
            (0.0, [(-s * np.sqrt(ky / kx), cx, sy)]),
            (0.0, [(s * np.sqrt(kx / ky), sx, cy)]),
            (self.rho_0, [(-K * (ky / kx), c2x, None), (-K * (kx / ky), None, c2y)]),
            (0.0, [(-nu * 2 * s * np.sqrt(ky * kx), sx, sy)]),
            (0.0, [(-nu * s * (np.sqrt(kx ** 3 / ky) - np.sqrt(ky ** 3 / kx)), cx, cy)]),
            (0.0, [(-nu * s * (np.sqrt(kx ** 3 / ky) - np.sqrt(ky ** 3 / kx)), cx, cy)]),
            (0.0, [(nu * 2 * s * np.sqrt(ky * kx), sx, sy)]),
        ]

    def viscosity(self):
        return self.nu

    def delta_x(self):
        return 1.0

    def delta_t(self):
        return 1.0


def decay_time(problem):
    """second_order_convergence.jl:126-131."""
    nu = problem.viscosity()
    kx, ky = 2 * np.pi / problem.NX, 2 * np.pi / problem.NY
    return 1 / (nu * (kx ** 2 + ky ** 2))


class TaylorGreenVortex(FluidFlowProblem):
    """taylor_green_vortex.jl."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(2 * np.pi, 2 * np.pi), static=True,
                 A=1, B=-1, a=1, b=1):
        NX = 16 * scale if NX is None else NX
        NY = NX if NY is None else NY
        self.rho_0 = 1.0
        self.u_max = 0.01 / scale
        self.nu = nu
        self.NX, self.NY = int(NX), int(NY)
        self.domain_size = (float(domain_size[0]), float(domain_size[1]))
        self.static = static
        self.A, self.B, self.a, self.b = float(A), float(B), float(a), float(b)

    def has_external_force(self):
        return self.static

    def decay(self, x, y, t):
        return 1.0 if self.static else np.exp(-(self.a ** 2 + self.b ** 2) * self.viscosity() * t)

    def pressure(self, q, x, y, t=0.0):
        P = -(1 / 4) * self.rho_0 * self.decay(x, y, t) ** 2 * (
            self.A ** 2 * np.cos(2 * self.a * x) + self.B ** 2 * np.cos(2 * self.b * y))
        return 1.0 + q.speed_of_sound_squared * self.u_max ** 2 * P

    def density(self, q, x, y, t=0.0):
        return self.pressure(q, x, y, t)

    def velocity(self, x, y, t=0.0):
        d = self.decay(x, y, t)
        return (d * (self.A * np.cos(self.a * x) * np.sin(self.b * y)),
                d * (self.B * np.sin(self.a * x) * np.cos(self.b * y)))

    def velocity_gradient(self, x, y, t=0.0):
        a, A, b, B = self.a, self.A, self.b, self.B
        u_x = -a * A * np.sin(a * x) * np.sin(b * y)
        v_y = -b * B * np.sin(a * x) * np.sin(b * y)
        u_y = b * A * np.cos(a * x) * np.cos(b * y)
        v_x = a * B * np.cos(a * x) * np.cos(b * y)
        d = self.decay(x, y, t)
        return ((d * u_x, d * v_x), (d * u_y, d * v_y))

    def force(self, x, y, t=0.0):
        if not self.static:
            z = 0.0 * np.asarray(x, dtype=np.float64)
            return z, z
        ux, uy = self.velocity(x, y, 0.0)
        s = 2 * self.viscosity()
        return s * ux, s * uy

    def expected_separable(self, q, t, y0=0, ny=None):
        x, y = self._xy(y0, ny)
        a, A, b, B = self.a, self.A, self.b, self.B
        d = self.decay(x, y, t)
        nu = self.viscosity()
        K = q.speed_of_sound_squared * self.u_max ** 2 * (-(1 / 4) * self.rho_0 * d ** 2)
        cx, sx, cy, sy = np.cos(a * x), np.sin(a * x), np.cos(b * y), np.sin(b * y)
        p = (1.0, [(K * A ** 2, np.cos(2 * a * x), None), (K * B ** 2, None, np.cos(2 * b * y))])
        return [p, (0.0, [(d * A, cx, sy)]), (0.0, [(d * B, sx, cy)]), p,
                (0.0, [(-nu * 2 * d * (-a * A), sx, sy)]),
                (0.0, [(-nu * d * (a * B + b * A), cx, cy)]),
                (0.0, [(-nu * d * (a * B + b * A), cx, cy)]),
                (0.0, [(-nu * 2 * d * (-b * B), sx, sy)])]


class DecayingShearFlow(FluidFlowProblem):
    """decaying_shear_flow.jl.  `DecayingShearFlow.fields(...)` is the positional struct constructor."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(2 * np.pi, 2 * np.pi), static=True,
                 A=1.0, B=1.0, k_x=1.0, k_y=0.0):
        NX = 8 * scale if NX is None else NX
        NY = NX if NY is None else NY
        if k_y == 0.0:
            NY = 3
        if k_x == 0.0:
            NX = 3
        self._set(1.0, 0.02 / scale, nu, NX, NY, domain_size, static, A, B, k_x, k_y)

    def _set(self, rho_0, u_max, nu, NX, NY, domain_size, static, A, B, k_x, k_y):
        self.rho_0, self.u_max, self.nu = rho_0, u_max, nu
        self.NX, self.NY = int(NX), int(NY)
        self.domain_size = (float(domain_size[0]), float(domain_size[1]))
        self.static = static
        self.A, self.B, self.k_x, self.k_y = float(A), float(B), float(k_x), float(k_y)

    @classmethod
    def fields(cls, rho_0, u_max, nu, NX, NY, domain_size, static, A, B, k_x, k_y):
        p = cls.__new__(cls)
        p._set(rho_0, u_max, nu, NX, NY, domain_size, static, A, B, k_x, k_y)
        return p

    def has_external_force(self):
        return self.static

    def decay(self, x, y, t):
        if self.static:
            return 1.0
        return np.exp(-1.0 * self.k_x ** 2 * self.viscosity() * t)

    def density(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * np.asarray(x, dtype=np.float64)

    def pressure(self, q, x, y, t=0.0):
        return 1 + (self.B * 0.025 * q.speed_of_sound_squared * self.u_max ** 2 * self.B
                    * np.sin(self.k_x * (x - self.A * t)) ** 2 * self.decay(x, y, t) ** 2)

    def velocity(self, x, y, t=0.0):
        A, B, kx, ky = self.A, self.B, self.k_x, self.k_y
        ux = A * np.cos(ky * y - ky * B * t) + 0.0 * x
        uy = B * np.cos(kx * x - kx * A * t) + 0.0 * y
        if self.static:
            return ux, uy
        return (ux * np.exp(-1.0 * ky ** 2 * self.viscosity() * t), uy * np.exp(-1.0 * kx ** 2 * self.viscosity() * t))

    def velocity_gradient(self, x, y, t=0.0):
        A, B, kx, ky = self.A, self.B, self.k_x, self.k_y
        z = 0.0 * (np.asarray(x, dtype=np.float64) + np.asarray(y, dtype=np.float64))
        u_y = -A * ky * np.sin(ky * (y - B * t)) + z
        v_x = -B * kx * np.sin(kx * (x - A * t)) + z
        if not self.static:
            u_y = u_y * np.exp(-1.0 * ky ** 2 * self.viscosity() * t)
            v_x = v_x * np.exp(-1.0 * kx ** 2 * self.viscosity() * t)
        return ((z, v_x), (u_y, z))

    def force(self, x, y, t=0.0):
        z = 0.0 * (np.asarray(x, dtype=np.float64) + np.asarray(y, dtype=np.float64))
        if not self.static:
            return z, z
        A, B, kx, ky = self.A, self.B, self.k_x, self.k_y
        nu = self.viscosity()
        return (nu * ky ** 2 * A * np.cos(ky * y - ky * B * t) + z, nu * kx ** 2 * B * np.cos(kx * x - kx * A * t) + z)

    def expected_separable(self, q, t, y0=0, ny=None):
        x, y = self._xy(y0, ny)
        A, B, kx, ky = self.A, self.B, self.k_x, self.k_y
        nu = self.viscosity()
        ex = 1.0 if self.static else np.exp(-1.0 * kx ** 2 * nu * t)
        ey = 1.0 if self.static else np.exp(-1.0 * ky ** 2 * nu * t)
        dec = self.decay(x, y, t)
        p = (1.0, [(B * 0.025 * q.speed_of_sound_squared * self.u_max ** 2 * B * dec ** 2, np.sin(kx * (x - A * t)) ** 2, None)])
        u_y = -A * ky * np.sin(ky * (y - B * t)) * ey   # d u_x / d y, function of y
        v_x = -B * kx * np.sin(kx * (x - A * t)) * ex   # d u_y / d x, function of x
        sxy = (0.0, [(-nu, v_x, None), (-nu, None, u_y)])
        return [(1.0, []), (0.0, [(A * ey, None, np.cos(ky * y - ky * B * t))]), (0.0, [(B * ex, np.cos(kx * x - kx * A * t), None)]),
                p, (0.0, []), sxy, sxy, (0.0, [])]

    def force_separable(self, t0, nsteps, y0=0, ny=None):
        """Lattice force of steps t0..t0+nsteps-1 as F_x(y, t), F_y(x, t) tables: the force above is
        a sum of a function of (y, t) and one of (x, t) (decaying_shear_flow.jl:131-147)."""
        dt = self.delta_t()
        return self.force_separable_times([(t0 + k) * dt for k in range(nsteps)], y0, ny)

    def force_separable_times(self, times, y0=0, ny=None):
        xr, yr = self.range()
        ny = self.NY - y0 if ny is None else ny
        yr = yr[y0:y0 + ny]
        s = self.u_max * self.delta_t()
        nu = self.viscosity()
        A, B, kx, ky = self.A, self.B, self.k_x, self.k_y
        fx = np.empty((len(times), len(yr)))
        fy = np.empty((len(times), len(xr)))
        for k, t in enumerate(times):
            fx[k] = s * (nu * ky ** 2 * A * np.cos(ky * yr - ky * B * t))
            fy[k] = s * (nu * kx ** 2 * B * np.cos(kx * xr - kx * A * t))
        return fx, fy


class PoiseuilleFlow(FluidFlowProblem):
    """poiseuille.jl.  `PoiseuilleFlow.fields(...)` is the positional struct constructor."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(1.0, 1.0), static=True):
        NXd = 5 * scale if NX is None else NX
        NY = NXd if NY is None else NY
        self._set(1.0, 0.1 / scale, nu, 3, NY, 1.0, domain_size, 1.0)

    def _set(self, rho_0, u_max, nu, NX, NY, k, domain_size, G):
        self.rho_0, self.u_max, self.nu = rho_0, u_max, nu
        self.NX, self.NY = int(NX), int(NY)
        self.k = k
        self.domain_size = (float(domain_size[0]), float(domain_size[1]))
        self.G = G

    @classmethod
    def fields(cls, rho_0, u_max, nu, NX, NY, k, domain_size, G):
        p = cls.__new__(cls)
        p._set(rho_0, u_max, nu, NX, NY, k, domain_size, G)
        return p

    def has_external_force(self):
        return True

    def delta_x(self):
        return self.domain_size[1] / self.NY

    def density(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def pressure(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def velocity(self, x, y, t=0.0):
        z = 0.0 * (np.asarray(x, dtype=np.float64) + y)
        return y * (self.domain_size[1] - y) * (self.G / 2) + z, z

    def velocity_gradient(self, x, y, t=0.0):
        z = 0.0 * (np.asarray(x, dtype=np.float64) + y)
        return ((z, z), ((self.domain_size[1] - 2 * y) * (self.G / 2) + z, z))

    def force_idx(self, x_idx, y_idx, t=0.0):
        # force(problem, x::Int, y::Int, t) poiseuille.jl:72-82 -- uniform
        return (self.viscosity() * self.G, 0.0)

    def expected_separable(self, q, t, y0=0, ny=None):
        x, y = self._xy(y0, ny)
        L, G, nu = self.domain_size[1], self.G, self.viscosity()
        sxy = (0.0, [(-nu, None, (L - 2 * y) * (G / 2))])
        return [(1.0, []), (0.0, [(1.0, None, y * (L - y) * (G / 2))]), (0.0, []), (1.0, []), (0.0, []), sxy, sxy, (0.0, [])]

    def force_uniform(self):
        s = self.u_max * self.delta_t()
        F = self.force_idx(1, 1)
        return s * F[0], s * F[1]

    def boundary_conditions(self):
        return [BounceBack(North(), (1, self.NX), (1, self.NY)), BounceBack(South(), (1, self.NX), (1, self.NY))]


class CouetteFlow(FluidFlowProblem):
    """couette_flow.jl.  `CouetteFlow.fields(...)` is the positional struct constructor."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(1.0, 1.0)):
        NXd = 5 * scale if NX is None else NX
        NY = NXd if NY is None else NY
        self._set(1.0, 0.01 / scale, nu, 1, NY, domain_size)

    def _set(self, rho_0, u_max, nu, NX, NY, domain_size):
        self.rho_0, self.u_max, self.nu = rho_0, u_max, nu
        self.NX, self.NY = int(NX), int(NY)
        self.domain_size = (float(domain_size[0]), float(domain_size[1]))

    @classmethod
    def fields(cls, rho_0, u_max, nu, NX, NY, domain_size):
        p = cls.__new__(cls)
        p._set(rho_0, u_max, nu, NX, NY, domain_size)
        return p

    def density(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def pressure(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def velocity(self, x, y, t=0.0):
        z = 0.0 * (np.asarray(x, dtype=np.float64) + y)
        return y + z, z

    def velocity_gradient(self, x, y, t=0.0):
        z = 0.0 * (np.asarray(x, dtype=np.float64) + y)
        return ((z, z), (1.0 + z, z))

    def force_idx(self, x_idx, y_idx, t=0.0):
        return (0.0, 0.0)

    def expected_separable(self, q, t, y0=0, ny=None):
        x, y = self._xy(y0, ny)
        nu = self.viscosity()
        return [(1.0, []), (0.0, [(1.0, None, y)]), (0.0, []), (1.0, []), (0.0, []), (-nu, []), (-nu, []), (0.0, [])]

    def boundary_conditions(self):
        return [BounceBack(South(), (1, self.NX), (1, self.NY)),
                MovingWall(North(), (1, self.NX), (1, self.NY), [self.u_max, 0])]


class LidDrivenCavityFlow(FluidFlowProblem):
    """lid_driven_cavity.jl."""

    def __init__(self, nu=1.0 / 6.0, scale=2, NX=None, NY=None, domain_size=(1.0, 1.0)):
        NX = 16 * scale if NX is None else NX
        NY = NX if NY is None else NY
        self.rho_0 = 1.0
        self.u_max = 0.01 / scale
        self.nu = nu
        self.NX, self.NY = int(NX), int(NY)
        self.domain_size = (float(domain_size[0]), float(domain_size[1]))

    def density(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def pressure(self, q, x, y, t=0.0):
        return 1.0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def velocity(self, x, y, t=0.0):
        z = 0.0 * (np.asarray(x, dtype=np.float64) + y)
        return z, z

    def boundary_conditions(self):
        return [BounceBack(East(), (1, self.NX), (1, self.NY)), BounceBack(South(), (1, self.NX), (1, self.NY)),
                BounceBack(West(), (1, self.NX), (1, self.NY)),
                MovingWall(North(), (1, self.NX), (1, self.NY), [self.u_max, 0])]


class _LinearizedMode(FluidFlowProblem):
    """Common fields of the two linearised hydrodynamic modes (linear_hydrodynamics_modes.jl)."""

    def __init__(self, nu, kappa, scale, NY):
        self.rho_0, self.theta_0 = 1.0, 1.0
        self.rho_tilde, self.u_tilde = 0.001, 0.001
        self.nu, self.kappa = nu, kappa
        self.domain_size = (2 * np.pi, 2 * np.pi)
        self.u_max = 0.01 / scale
        self.NX = self.NY = int(NY)

    def delta_x(self):
        return self.domain_size[1] / self.NY

    def heat_diffusion(self):
        return self.kappa * self.delta_x() ** 2 / self.delta_t()  # problems.jl:81-82


class LinearizedThermalDiffusion(_LinearizedMode):
    """LinearizedThermalDiffusion(nu, kappa, scale, NY = 4 scale): rho = rho_0 + rho~ sin(y) exp(-kappa t), u = 0,
    p = rho_0 theta_0 (linear_hydrodynamics_modes.jl:15-99)."""

    def __init__(self, nu, kappa, scale, NY=None):
        super().__init__(nu, kappa, scale, 4 * scale if NY is None else NY)

    def density(self, q, x, y, t=0.0):
        return self.rho_0 + self.rho_tilde * np.sin(y) * np.exp(-self.heat_diffusion() * t) + 0.0 * np.asarray(x, dtype=np.float64)

    def pressure(self, q, x, y, t=0.0):
        return self.rho_0 * self.theta_0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def velocity(self, x, y, t=0.0):
        z = 0.0 * (np.asarray(x, dtype=np.float64) + y)
        return z, z

    def expected_separable(self, q, t, y0=0, ny=None):
        x, y = self._xy(y0, ny)
        return [(self.rho_0, [(self.rho_tilde * np.exp(-self.heat_diffusion() * t), None, np.sin(y))]), (0.0, []), (0.0, []),
                (self.rho_0 * self.theta_0, []), (0.0, []), (0.0, []), (0.0, []), (0.0, [])]


class LinearizedTransverseShearWave(_LinearizedMode):
    """LinearizedTransverseShearWave(nu, kappa, scale, NY = 8 scale): rho = rho_0, u = [u~ sin(y) exp(-nu t), 0],
    p = theta_0 (linear_hydrodynamics_modes.jl:101-178)."""

    def __init__(self, nu, kappa, scale, NY=None):
        super().__init__(nu, kappa, scale, 8 * scale if NY is None else NY)

    def density(self, q, x, y, t=0.0):
        return self.rho_0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def pressure(self, q, x, y, t=0.0):
        return self.theta_0 + 0.0 * (np.asarray(x, dtype=np.float64) + y)

    def velocity(self, x, y, t=0.0):
        z = 0.0 * (np.asarray(x, dtype=np.float64) + y)
        return self.u_tilde * np.sin(y) * np.exp(-self.viscosity() * t) + z, z

    def expected_separable(self, q, t, y0=0, ny=None):
        x, y = self._xy(y0, ny)
        return [(self.rho_0, []), (0.0, [(self.u_tilde * np.exp(-self.viscosity() * t), None, np.sin(y))]), (0.0, []),
                (self.theta_0, []), (0.0, []), (0.0, []), (0.0, []), (0.0, [])]
