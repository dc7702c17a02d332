"""TEST INFRASTRUCTURE: an oracle-backed stand-in for lbm._abi.Context.

The host mirror (latticeboltzmann.jl_b200/lbm) only ever talks to the CUDA library through `_abi.Context`.  This class
implements the same methods on the CPU with the numpy oracle, so that the host-side logic of the drop-in (problems,
initialisation strategies, force data, batching of `simulate`, processing methods, stop criteria, unit scaling) can be run
end to end in the `-m "not gpu"` suite -- e.g. against the reference's notebook figures.  It is never importable from the
product: it lives under tests/ and is installed by monkeypatching `lbm.model.make_context` (see `emulated_backend`).
"""
import contextlib

import numpy as np

import oracle.lbm_oracle as O
from lbm import _abi

D = 2
_DIRS = {0: "N", 1: "E", 2: "S", 3: "W"}


def _to_oracle(f_xyq):
    return np.ascontiguousarray(np.transpose(np.asarray(f_xyq, dtype=np.float64), (2, 1, 0)))


def _to_host(f_qyx):
    return np.asfortranarray(np.transpose(f_qyx, (2, 1, 0)))


class _Pinned:  # prescribed velocity for the iterative-initialisation operator
    def __init__(self, ux, uy):
        self.NY, self.NX = ux.shape
        self.u_max = 1.0
        self._u = (ux, uy)

    def grid(self):
        return np.meshgrid(np.arange(self.NX, dtype=float), np.arange(self.NY, dtype=float))

    def velocity(self, X, Y, t=0.0):
        return self._u


class OracleContext:
    def __init__(self, nx, ny, lattice, collision, tau, bcs=(), dtype=_abi.F64, arith=_abi.ARITH_EXACT, device=0,
                 rank=0, world=1, nccl_id=None):
        assert world == 1
        self.q = O.L.BY_NAME[lattice]() if isinstance(lattice, str) else O.L.BY_NAME[[k for k, v in _abi.LATTICE_IDS.items() if v == lattice][0]]()
        self.nx, self.ny, self.y0, self.ny_local = int(nx), int(ny), 0, int(ny)
        self.Q = self.q.Q
        self.shape = (self.nx, self.ny, self.Q)
        self.collision, self.tau = int(collision), [float(t) for t in np.atleast_1d(tau)]
        self.bcs = []
        for b in bcs:
            xs, ys = (b.x0, b.x1), (b.y0, b.y1)
            if b.kind == 1:
                self.bcs.append(O.MovingWall(_DIRS[b.direction], xs, ys, [b.u[0], b.u[1]], b.rho, b.T))
            else:
                self.bcs.append(O.BounceBack(_DIRS[b.direction], xs, ys))
        self.f = np.zeros((self.Q, self.ny, self.nx))
        self.fc = None
        self.force = None  # None | ("uniform", fx, fy) | ("field", Fx, Fy) | ("sep", t0, fx_of_y, fy_of_x) | ("u0", ux, uy)
        self.u_old = None
        self.rho_old = None
        self.kernel_launches = 0
        self.halo_path = 0

    # -- plumbing -----------------------------------------------------------------------------
    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def sync(self):
        pass

    def set_option(self, key, value):
        pass

    def new_f(self):
        return np.empty(self.shape, order="F")

    def upload_f(self, f):
        self.f, self.fc = _to_oracle(f), None

    def upload_f_collision(self, f):
        self.fc = _to_oracle(f)

    def download_f(self, out=None):
        r = _to_host(self.f)
        if out is not None:
            out[...] = r
            return out
        return r

    def download_f_collision(self, out=None):
        return _to_host(self.fc)

    def init_equilibrium_rows(self, y0, rho, ux, uy, T):
        ny = np.shape(rho)[1]
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (self.nx, ny)).T) for a in (rho, ux, uy, T)]
        f = np.stack(O.hermite_based_equilibrium(self.q, *arrs))
        self.f[:, y0:y0 + ny] = f
        self.fc = None

    def init_analytic(self, fields, unit_density=False, unit_temperature=False, offeq=0, offeq_coef=0.0):
        """what lbm_init_analytic evaluates, with the oracle's equilibrium: fields -> rho, u, T (+ off-equilibrium part)"""
        e = []
        for c0, terms in fields:
            E = np.zeros((self.nx, self.ny)) + c0
            for a, X, Y in terms:
                E = E + a * np.outer(np.ones(self.nx) if X is None else X, np.ones(self.ny) if Y is None else Y)
            e.append(np.ascontiguousarray(E.T))
        rho = np.ones_like(e[0]) if unit_density else e[0]
        T = np.ones_like(e[0]) if unit_temperature else e[3] / e[0]
        f = np.stack(O.hermite_based_equilibrium(self.q, rho, e[1], e[2], T))
        if offeq:
            sxx, sxy, syy = e[4] + e[4], e[5] + e[6], e[7] + e[7]
            kk = offeq_coef * rho if offeq == 2 else offeq_coef
            for i in range(self.q.Q):
                H = O.hermite(2, (int(self.q.cx[i]), int(self.q.cy[i])), self.q)
                f[i] += self.q.w[i] * kk * (H[0][0] * sxx + H[0][1] * sxy + H[1][0] * sxy + H[1][1] * syy)
        self.f[:] = f
        self.fc = None

    # -- force --------------------------------------------------------------------------------
    def set_force_none(self):
        self.force = None

    def set_force_uniform(self, fx, fy):
        self.force = ("uniform", float(fx), float(fy))

    def set_force_field(self, Fx, Fy):
        self.force = ("field", np.asarray(Fx, dtype=np.float64).T.copy(), np.asarray(Fy, dtype=np.float64).T.copy())

    def set_velocity_field(self, ux, uy):
        self.force = ("u0", np.asarray(ux, dtype=np.float64).T.copy(), np.asarray(uy, dtype=np.float64).T.copy())

    def set_force_separable(self, t0, fx_of_y, fy_of_x):
        self.force = ("sep", int(t0), np.asarray(fx_of_y, dtype=np.float64), np.asarray(fy_of_x, dtype=np.float64))

    def _cm(self, step):
        F = None
        fm = self.force
        if fm is not None and fm[0] == "uniform":
            F = (fm[1], fm[2])
        elif fm is not None and fm[0] == "field":
            F = (fm[1], fm[2])
        elif fm is not None and fm[0] == "sep":
            s = step - fm[1]
            F = (np.broadcast_to(fm[2][s][:, None], (self.ny, self.nx)), np.broadcast_to(fm[3][s][None, :], (self.ny, self.nx)))
        if self.collision == _abi.SRT:
            return O.SRT(self.tau[0], F)
        if self.collision == _abi.TRT:
            return O.TRT(self.tau[0], self.tau[1], F)
        if self.collision == _abi.MRT:
            return O.MRT(self.q, self.tau, F)
        if fm is None or fm[0] != "u0":
            raise _abi.LbmError(-5, "LBM_ITERATIVE_INIT needs the prescribed velocity")
        return O.IterativeInitializationCollisionModel(self.q, self.tau[0], _Pinned(fm[1], fm[2]))

    # -- operators ----------------------------------------------------------------------------
    def collide(self, step=0, time=0.0):
        self.fc = O.collide(self._cm(step), self.q, self.f)

    def stream(self):
        self.f = O.stream(self.q, self.fc)

    def apply_bcs(self, time=0.0):
        O.apply_bcs(self.bcs, self.q, self.f, self.fc)

    def step(self, t0, nsteps, dt=1.0):
        cache = None
        for t in range(int(t0), int(t0) + int(nsteps)):
            if cache is None or (self.force is not None and self.force[0] == "sep"):
                cache = self._cm(t)
            self.f, self.fc = O.step(cache, self.q, self.bcs, self.f)
        self.kernel_launches += int(nsteps)

    # -- diagnostics --------------------------------------------------------------------------
    def _fields(self, tau_visc):
        q = self.q
        f = [self.f[i] for i in range(q.Q)]
        rho = O.density(q, f)
        ux, uy = O.velocity(q, f, rho)
        a = O._a_bar_2(q, f)
        den = 1 + 1 / (2 * tau_visc)
        exx, exy, eyy = rho * (ux * ux), rho * (ux * uy), rho * (uy * uy)
        bxx, byy = (a[0, 0] + exx / (2 * tau_visc)) / den, (a[1, 1] + eyy / (2 * tau_visc)) / den
        p_track = ((bxx - rho * (ux * ux - 1)) + (byy - rho * (uy * uy - 1))) / 2
        sxx, sxy, syy = (a[0, 0] - exx) / den, (a[0, 1] - exy) / den, (a[1, 1] - eyy) / den
        tr = (sxx + syy) / 2
        return dict(rho=rho, ux=ux, uy=uy, p=O.pressure(q, f, rho, ux, uy), p_track=p_track, sxx=sxx - tr, sxy=sxy, syy=syy - tr)

    def moments(self, tau_visc=1.0, fields=("rho", "ux", "uy")):
        h = self._fields(tau_visc)
        return {k: np.asfortranarray(h[k].T) for k in fields}

    def reduce(self, kind):
        h = self._fields(1.0)
        rho, ux, uy = h["rho"], h["ux"], h["uy"]
        if kind == _abi.REDUCE_MEAN_UX:
            return np.array([ux.sum(), ux.size, np.isnan(ux).sum(), 0.0])
        if kind == _abi.REDUCE_VELOCITY_CHANGE:
            if self.u_old is None:
                self.u_old = (np.zeros_like(ux), np.zeros_like(uy))
            ox, oy = self.u_old
            out = np.array([((ux - ox) ** 2 + (uy - oy) ** 2).sum(), (ox ** 2 + oy ** 2).sum(), 0.0, 0.0])
            self.u_old = (ux.copy(), uy.copy())
            return out
        if kind == _abi.REDUCE_DENSITY_CHANGE:
            if self.rho_old is None:
                self.rho_old = np.zeros_like(rho)
            out = np.array([((rho - self.rho_old) ** 2).sum(), rho[-1, -1], 0.0, 0.0])
            self.rho_old = rho.copy()
            return out
        return np.array([rho.sum(), (rho * (ux + uy)).sum(), (rho * (ux * ux + uy * uy)).sum(), 0.0])

    def reduce_errors(self, tau_visc, u_max, expected):
        h = self._fields(tau_visc)
        e = []
        for c0, terms in expected:
            v = c0 + np.zeros((self.ny, self.nx))
            for a, X, Y in terms:
                xs = np.ones(self.nx) if X is None else np.asarray(X, dtype=np.float64)
                ys = np.ones(self.ny) if Y is None else np.asarray(Y, dtype=np.float64)
                v = v + a * (ys[:, None] * xs[None, :])
            e.append(v)
        rho, vx, vy = h["rho"], h["ux"] / u_max, h["uy"] / u_max
        fac = 1 / (u_max * u_max)
        sxx, sxy, syy, pr = h["sxx"] * fac, h["sxy"] * fac, h["syy"] * fac, h["p_track"]
        s = [((rho - e[0]) ** 2), (vx - e[1]) ** 2 + (vy - e[2]) ** 2, e[1] ** 2 + e[2] ** 2, (pr - e[3]) ** 2, e[3] ** 2,
             (e[4] - sxx) ** 2, e[4] ** 2, (e[5] - sxy) ** 2, e[5] ** 2, (e[7] - syy) ** 2, e[7] ** 2, (e[6] - sxy) ** 2, e[6] ** 2,
             rho, rho * (vx + vy), rho * (vx * vx + vy * vy)]
        return np.array([np.sum(x) for x in s])


    def reduce_process(self, u_max, expected):
        """the 12 sums of process! (processing_methods.jl:177-239) from the oracle's moments"""
        h = self._fields(1.0)
        e = []
        for c0, terms in expected[:4]:
            v = c0 + np.zeros((self.ny, self.nx))
            for a, X, Y in terms:
                xs = np.ones(self.nx) if X is None else np.asarray(X, dtype=np.float64)
                ys = np.ones(self.ny) if Y is None else np.asarray(Y, dtype=np.float64)
                v = v + a * (ys[:, None] * xs[None, :])
            e.append(v)
        rho, vx, vy, pr = h["rho"], h["ux"] / u_max, h["uy"] / u_max, h["p"]
        T, kin = pr / rho, (vx ** 2 + vy ** 2) * rho
        eT, ekin = e[3] / e[0], e[1] ** 2 + e[2] ** 2
        s = [rho, (vx + vy) * rho, kin + T, kin, T, e[0], e[0] * (e[1] + e[2]), ekin + eT, ekin, eT,
             (vx - e[1]) ** 2 + (vy - e[2]) ** 2, (pr - e[3]) ** 2]
        return np.array([np.sum(x) for x in s])

@contextlib.contextmanager
def emulated_backend(monkeypatch):
    """Route lbm.model.make_context to the oracle-backed context for the duration of a test."""
    import lbm
    from lbm import model

    def make_context(q, cm, bcs, nx, ny, dtype="f64", arith="exact", comm=None, device=None):
        return OracleContext(nx, ny, q.name, model._cm_code(cm), cm.taus(), [b.to_abi() for b in bcs])

    monkeypatch.setattr(model, "make_context", make_context)
    yield lbm
