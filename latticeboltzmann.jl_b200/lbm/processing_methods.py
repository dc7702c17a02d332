"""Processing methods and stop criteria of the Julia API (src/processing_methods{.jl,/*}).

`next_(pm, q, f_in, t)` mirrors `next!(pm, q, f_in, t)::Bool`; `f_in` is the device-resident
state (`DeviceState`, what `simulate` passes) or a host array (uploaded to a scratch context).
Per-node moments and grid reductions run in the CUDA library (`lbm_moments`, `lbm_reduce`);
the host only combines them with the problem's analytic fields.  Plotting is out of scope.
"""
import math
import warnings

import numpy as np

from . import _abi
from .separable import decompose_function, separable_from_fields
from .problems import (CouetteFlow, DecayingShearFlow, LidDrivenCavityFlow, PoiseuilleFlow, TGV,
                       TaylorGreenVortex)


# ---------------------------------------------------------------------------------------------
# stop criteria (processing_methods/stopping_criteria/stopping_criteria.jl)
# ---------------------------------------------------------------------------------------------
class StopCriteriaBase:
    active = True

    def should_stop_(self, q, state):
        return False


class NoStoppingCriteria(StopCriteriaBase):
    active = False


class MeanVelocityStoppingCriteria(StopCriteriaBase):
    """:3-55."""

    def __init__(self, old_mean_velocity, tolerance, problem):
        self.old_mean_velocity = old_mean_velocity
        self.tolerance = tolerance
        self.problem = problem

    def should_stop_(self, q, state):
        s = state.reduce(_abi.REDUCE_MEAN_UX)
        u_mean = np.float64(s[0]) / np.float64(s[1])
        with np.errstate(divide="ignore", invalid="ignore"):
            converged = abs(u_mean / np.float64(self.old_mean_velocity) - 1)
        if converged < self.tolerance:
            return True
        if math.isnan(u_mean):
            warnings.warn("nan in velocity profile")
            return True
        self.old_mean_velocity = float(u_mean)
        return False


class VelocityConvergenceStoppingCriteria(StopCriteriaBase):
    """:57-115; the previous velocity field lives on the device."""

    def __init__(self, tolerance, problem):
        self.tolerance = tolerance
        self.problem = problem

    def should_stop_(self, q, state):
        s = state.reduce(_abi.REDUCE_VELOCITY_CHANGE)
        with np.errstate(divide="ignore", invalid="ignore"):
            converged = np.sqrt(np.float64(s[0])) / np.float64(s[1])  # denominator not sqrt'ed (:101)
        if converged < self.tolerance:
            return True
        if math.isnan(converged):
            warnings.warn("nan in velocity profile")
            return True
        return False


class DensityConvergence(StopCriteriaBase):
    """stopping_criteria/density_convergence.jl:1-17.  Restated literally: the loop `for x_idx in nx, y_idx in ny` (:9)
    visits only the node (NX, NY), so norm(rho - rho_old) is the density change of that single node (everything else
    stays at its initial 0).  `whole_field=True` uses the norm over all nodes, which is what the code evidently
    intended.  rho_old lives on the device (lbm_reduce kind LBM_REDUCE_DENSITY_CHANGE)."""

    def __init__(self, eps, problem=None, whole_field=False):
        self.eps = float(eps)
        self.problem = problem
        self.whole_field = whole_field
        self._corner_old = 0.0

    def should_stop_(self, q, state):
        s = state.reduce(_abi.REDUCE_DENSITY_CHANGE)
        if self.whole_field:
            d = float(np.sqrt(np.float64(s[0])))
        else:
            d = abs(float(s[1]) - self._corner_old)
            self._corner_old = float(s[1])
        return d < self.eps or d > 100.0


def StopCriteria(problem):
    """:8-15."""
    if isinstance(problem, PoiseuilleFlow):
        return MeanVelocityStoppingCriteria(0.0, 1e-12, problem)
    if isinstance(problem, CouetteFlow):
        return MeanVelocityStoppingCriteria(0.0, 1e-7, problem)
    if isinstance(problem, LidDrivenCavityFlow):
        return MeanVelocityStoppingCriteria(0.0, 1e-5, problem)
    if isinstance(problem, DecayingShearFlow) and problem.static:
        return MeanVelocityStoppingCriteria(0.0, 1e-8, problem)
    return NoStoppingCriteria()


def should_stop_(sc, q, state):
    return sc.should_stop_(q, state)


# ---------------------------------------------------------------------------------------------
# processing methods
# ---------------------------------------------------------------------------------------------
class ProcessingMethodBase:
    problem = None

    def noop(self, t):
        """True when next_(pm, q, f, t) neither reads f nor has side effects -- lets `simulate`
        keep the device running without a host round trip."""
        return False


def _sdiv(a, b):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(np.float64(a) / np.float64(b))


class TrackHydrodynamicErrors(ProcessingMethodBase):
    """track_hydrodynamic_errors.jl (visualisation omitted).  df rows are dicts with the
    reference's field names (σ spelled `s`: error_σ_xx -> error_sxx ...)."""

    def __init__(self, problem, should_process, n_steps, stop_criteria=None, device_norms=True):
        self.problem = problem
        self.should_process = should_process
        self.n_steps = n_steps
        self.stop_criteria = StopCriteria(problem) if stop_criteria is None else stop_criteria
        self.df = []
        # True: error sums are reduced on the device when the problem provides its analytic fields in
        # separable form (expected_separable); False: per-node fields are downloaded and compared on the host
        self.device_norms = device_norms

    def noop(self, t):
        if t % 100 == 0 and self.stop_criteria.active:
            return False
        return (not self.should_process) and t != self.n_steps

    def next_(self, q, state, t):
        should_stop = False
        if t % 100 == 0:
            if self.stop_criteria.should_stop_(q, state):
                should_stop = True
        if (not should_stop) and t != self.n_steps:
            if not self.should_process:
                return False
        pr = self.problem
        nx, ny = pr.NX, pr.NY
        xstep, ystep = pr.range_steps()
        time = t * pr.delta_t()
        Delta = ystep * xstep
        if nx == 1:
            Delta = ystep
            if ny == 1:
                Delta = 1.0
        elif ny == 1:
            Delta = xstep
        tau = q.speed_of_sound_squared * pr.lattice_viscosity()
        sep = pr.expected_separable(q, time, state.y0, state.ny_local) if self.device_norms else None
        if sep is None and self.device_norms:
            # no analytic separable form: recover one from O(NX + NY) evaluations of the pointwise fields (exact for
            # rank <= 2), falling back to the fields sampled on the whole grid
            xs, ys = pr._xy(state.y0, state.ny_local)
            dev = lambda a, b: (lambda X, Y: pr.deviatoric_tensor(q, X, Y, time)[a][b])  # noqa: E731
            sep = [decompose_function(fn, xs, ys) for fn in (
                lambda X, Y: pr.density(q, X, Y, time), lambda X, Y: pr.velocity(X, Y, time)[0],
                lambda X, Y: pr.velocity(X, Y, time)[1], lambda X, Y: pr.pressure(q, X, Y, time),
                dev(0, 0), dev(0, 1), dev(1, 0), dev(1, 1))]
            if any(d is None for d in sep):
                X, Y = pr.grid(state.y0, state.ny_local)
                e_ux, e_uy = pr.velocity(X, Y, time)
                (e_sxx, e_sxy), (e_syx, e_syy) = pr.deviatoric_tensor(q, X, Y, time)
                sep = separable_from_fields([np.asarray(a, dtype=np.float64) * np.ones_like(X) for a in (
                    pr.density(q, X, Y, time), e_ux, e_uy, pr.pressure(q, X, Y, time), e_sxx, e_sxy, e_syx, e_syy)])
        if sep is not None:
            # all 16 sums on the device (lbm_reduce_errors); nothing but scalars crosses PCIe
            s = state.allreduce(state.ctx.reduce_errors(tau, pr.u_max, sep))
            self.df.append(self._row(t, s, Delta))
            return should_stop
        h = state.moments(tau, ("rho", "ux", "uy", "p_track", "sxx", "sxy", "syy"))
        X, Y = pr.grid(state.y0, state.ny_local)
        e_rho = pr.density(q, X, Y, time)
        e_ux, e_uy = pr.velocity(X, Y, time)
        e_p = pr.pressure(q, X, Y, time)
        (e_sxx, e_sxy), (e_syx, e_syy) = pr.deviatoric_tensor(q, X, Y, time)
        rho = h["rho"]
        ux, uy = h["ux"] / pr.u_max, h["uy"] / pr.u_max
        fac = 1 / pr.u_max ** 2
        sxx, sxy, syy = h["sxx"] * fac, h["sxy"] * fac, h["syy"] * fac
        sums = np.array([
            np.sum((rho - e_rho) ** 2), np.sum((ux - e_ux) ** 2 + (uy - e_uy) ** 2), np.sum(e_ux ** 2 + e_uy ** 2),
            np.sum((h["p_track"] - e_p) ** 2), np.sum(e_p ** 2),
            np.sum((e_sxx - sxx) ** 2), np.sum(e_sxx ** 2), np.sum((e_sxy - sxy) ** 2), np.sum(e_sxy ** 2),
            np.sum((e_syy - syy) ** 2), np.sum(e_syy ** 2), np.sum((e_syx - sxy) ** 2), np.sum(e_syx ** 2),
            np.sum(rho), np.sum(rho * (ux + uy)), np.sum(rho * (ux ** 2 + uy ** 2))])
        s = state.allreduce(sums)
        self.df.append(self._row(t, s, Delta))
        return should_stop

    @staticmethod
    def _row(t, s, Delta):
        # track_hydrodynamic_errors.jl:205-221
        return dict(
            timestep=t, error_rho=float(np.sqrt(s[0])), error_u=float(_sdiv(s[1], s[2])),
            error_p=float(_sdiv(s[3], s[4])), error_sxx=float(_sdiv(s[5], s[6])), error_sxy=float(_sdiv(s[7], s[8])),
            error_syy=float(_sdiv(s[9], s[10])), error_syx=float(_sdiv(s[11], s[12])),
            mass=Delta * s[13], momentum=Delta * s[14], energy=Delta * s[15])


class CompareWithAnalyticalSolution(ProcessingMethodBase):
    """processing_methods.jl:31-269 (visualisation omitted)."""

    def __init__(self, problem, should_process, n_steps, stop_criteria=None):
        self.problem = problem
        self.should_process = should_process
        self.n_steps = n_steps
        self.stop_criteria = StopCriteria(problem) if stop_criteria is None else stop_criteria
        self.df = []

    def noop(self, t):
        if t % 100 == 0 and self.stop_criteria.active:
            return False
        return (not self.should_process) and t != self.n_steps

    def next_(self, q, state, t):
        pr = self.problem
        if t % 100 == 0:
            if self.stop_criteria.should_stop_(q, state):
                process_(pr, q, state, t * pr.delta_t(), self.df)
                return True
        if not self.should_process:
            if t != self.n_steps:
                return False
        process_(pr, q, state, t * pr.delta_t(), self.df)
        return False


def process_(problem, q, state, time, stats, should_visualize=False, device_sums=True):
    """process!(problem, q, f_in, time, stats) (processing_methods.jl:142-269)."""
    pr = problem
    xstep, ystep = pr.range_steps()
    opp = ystep * xstep
    sep = _process_separable(pr, q, time, state) if device_sums else None
    if sep is not None:
        # all 12 sums on the device (lbm_reduce_process): only scalars cross PCIe
        s = state.allreduce(state.ctx.reduce_process(pr.u_max, sep))
        s[10] *= opp
        s[11] *= opp
        stats.append(_process_row(s))
        return False
    h = state.moments(1.0, ("rho", "ux", "uy", "p"))
    rho, p = h["rho"], h["p"]
    T = p / rho
    ux, uy = h["ux"] / pr.u_max, h["uy"] / pr.u_max
    kin = (ux ** 2 + uy ** 2) * rho
    X, Y = pr.grid(state.y0, state.ny_local)
    e_rho = pr.density(q, X, Y, time)
    e_p = pr.pressure(q, X, Y, time)
    e_ux, e_uy = pr.velocity(X, Y, time)
    e_T = e_p / e_rho
    e_kin = e_ux ** 2 + e_uy ** 2
    s = state.allreduce(np.array([
        np.sum(rho), np.sum((ux + uy) * rho), np.sum(kin + T), np.sum(kin), np.sum(T),
        np.sum(e_rho), np.sum(e_rho * (e_ux + e_uy)), np.sum(e_kin + e_T), np.sum(e_kin), np.sum(e_T),
        np.sum(opp * ((ux - e_ux) ** 2 + (uy - e_uy) ** 2)), np.sum(opp * (p - e_p) ** 2)]))
    stats.append(_process_row(s))
    return False


def _process_row(s):
    return dict(
        density=s[0], momentum=s[1], total_energy=s[2], kinetic_energy=s[3], internal_energy=s[4],
        density_a=s[5], momentum_a=s[6], total_energy_a=s[7], kinetic_energy_a=s[8], internal_energy_a=s[9],
        error_u=float(np.sqrt(s[10])), error_p=float(np.sqrt(s[11])),
        error_sxx=0.0, error_sxy=0.0, error_syy=0.0, error_syx=0.0)


def _process_separable(pr, q, time, state):
    """rho, ux, uy, p of the problem at `time` in separable form for the local slab: the problem's own closed form when it
    has one, else a cross approximation from O(NX + NY) evaluations of its pointwise functions; None if not of rank <= 2."""
    if not hasattr(state.ctx, "reduce_process"):
        return None
    sep = pr.expected_separable(q, time, state.y0, state.ny_local)
    if sep is not None:
        return sep[:4]
    xs, ys = pr._xy(state.y0, state.ny_local)
    out = []
    for fn in (lambda X, Y: pr.density(q, X, Y, time), lambda X, Y: pr.velocity(X, Y, time)[0],
               lambda X, Y: pr.velocity(X, Y, time)[1], lambda X, Y: pr.pressure(q, X, Y, time)):
        d = decompose_function(fn, xs, ys)
        if d is None:
            return None
        out.append(d)
    return out


class TakeSnapshots(ProcessingMethodBase):
    """take_snapshots.jl:3-29: snapshot = copy(f_in).  On the device that is one kernel writing f_stream of the current
    state into a compact buffer plus a device-to-host copy on a copy stream (lbm_snapshot_begin); the copy of snapshot k
    is only waited for when snapshot k+1 is taken or `snapshots` is read, so the step loop never stalls on PCIe."""

    def __init__(self, problem, every_t):
        self.problem = problem
        self.every_t = every_t
        self._snapshots = []
        self.timesteps = []
        self._pending = None

    def noop(self, t):
        if isinstance(self.every_t, int):
            return t % self.every_t != 0
        return t not in self.every_t

    def flush(self):
        if self._pending is not None:
            self._pending.snapshot_end()
            self._pending = None

    @property
    def snapshots(self):
        self.flush()
        return self._snapshots

    def next_(self, q, state, t):
        if self.noop(t):
            return False
        self.flush()
        ctx = state.ctx
        if hasattr(ctx, "snapshot_begin"):
            self._snapshots.append(ctx.snapshot_begin())
            self._pending = ctx
        else:
            self._snapshots.append(state.download_f())
        self.timesteps.append(t)
        return False


class ProcessIterativeInitialization(ProcessingMethodBase):
    """processing_methods/process_iterative_initialization.jl:1-26: only the stop criterion runs (the inner
    process method call is commented out in the reference, :18)."""

    def __init__(self, eps, problem, process_method=None, whole_field=False):
        self.stop_criteria = DensityConvergence(eps, problem, whole_field)
        self.internal_process_method = process_method
        self.n_steps = 100
        self.calls = 0

    def next_(self, q, state, t):
        self.calls += 1
        return should_stop_(self.stop_criteria, q, state)


def ProcessingMethod(problem, should_process, n_steps, stop_criteria=None):
    """processing_methods.jl:10-29."""
    if isinstance(problem, (TaylorGreenVortex, DecayingShearFlow, TGV)):
        return TrackHydrodynamicErrors(problem, should_process, n_steps, stop_criteria)
    return CompareWithAnalyticalSolution(problem, should_process, n_steps, stop_criteria)
