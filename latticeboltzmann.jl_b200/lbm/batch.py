"""Batched `simulate`: the reference's parameter studies as ONE device batch.

The reference sweeps parameters with a host loop around `simulate(problem, q; ...)`
(examples/notebooks/trt_magic_parameter.ipynb:30-103: 902 500 solves of a 3 x 5 Poiseuille flow;
poiseuille.ipynb cell 6; notebook_examples.jl).  `simulate_many` is that loop for solves that share one shape
-- grid, lattice, collision-model type, boundary conditions, number of steps -- and differ in relaxation times
and problem parameters (viscosity, hence force and expected fields).  All of them advance in a single launch of the
library's on-chip batch kernel (lbm_batch_*), each with its own stop criterion state, and the
`TrackHydrodynamicErrors` row the reference would have recorded last (`result.processing_method.df[end]`) is
evaluated for every solve on the device.
"""
import numpy as np

from . import _abi
from .boundary_conditions import BounceBack, MovingWall
from .collision_models import MRT, SRT, TRT, LatticeForce
from .initial_conditions import ZeroVelocityInitialCondition, default_strategy, initialize
from .processing_methods import MeanVelocityStoppingCriteria, NoStoppingCriteria, VelocityConvergenceStoppingCriteria

_CM = {SRT: _abi.SRT, TRT: _abi.TRT, MRT: _abi.MRT}
_DTYPES = {"f64": _abi.F64, "f32": _abi.F32}
_ARITH = {"exact": _abi.ARITH_EXACT, "fast": _abi.ARITH_FAST}


def _bc_key(bcs):
    return tuple((type(b).__name__, type(b.direction).__name__, tuple(b.xs), tuple(b.ys),
                  tuple(np.ravel(getattr(b, "u", ()))), getattr(b, "rho", None)) for b in bcs)


def _stop_spec(stop_criteria):
    """(kind, tolerance, old_mean) of a stop-criterion prototype (its state lives per solve on the device)"""
    if stop_criteria is None or isinstance(stop_criteria, NoStoppingCriteria):
        return _abi.BATCH_STOP_OFF, 0.0
    if isinstance(stop_criteria, VelocityConvergenceStoppingCriteria):
        return _abi.BATCH_STOP_VELOCITY_CONVERGENCE, float(stop_criteria.tolerance)
    if isinstance(stop_criteria, MeanVelocityStoppingCriteria):
        if stop_criteria.old_mean_velocity != 0.0:
            raise ValueError("batched MeanVelocityStoppingCriteria starts from old_mean_velocity = 0 (StopCriteria(problem))")
        return _abi.BATCH_STOP_MEAN_VELOCITY, float(stop_criteria.tolerance)
    raise TypeError(f"stop criterion {stop_criteria!r} is not available in batches")


class BatchResult:
    """Per solve k: `timestep[k]` (the t of the recorded row), `stopped[k]`, and the columns of the
    TrackHydrodynamicErrors row (`error_u[k]`, `error_p[k]`, `error_sxx[k]`, ...); `row(k)` returns the reference's
    NamedTuple as a dict, `f_stream(k)` the populations the row was computed from."""

    COLUMNS = ("error_rho", "error_u", "error_p", "error_sxx", "error_sxy", "error_syy", "error_syx", "mass", "momentum",
               "energy")

    def __init__(self, batch, timestep, stopped, columns, device_ms):
        self._batch = batch
        self.timestep, self.stopped, self.device_ms = timestep, stopped, device_ms
        for k, v in columns.items():
            setattr(self, k, v)

    def __len__(self):
        return len(self.timestep)

    def row(self, k):
        d = dict(timestep=int(self.timestep[k]))
        d.update({c: float(getattr(self, c)[k]) for c in self.COLUMNS})
        return d

    def f_stream(self, k, count=1):
        f = self._batch.download_f(k, count)
        return f[..., 0] if count == 1 else f

    def close(self):
        self._batch.close()


def simulate_many(problems, q, taus, collision_model=TRT, *, problem_index=None, t_end=1.0, stop_criteria=None,
                  initialization_strategy=None, forced=True, check_every=100, dtype="f64", arith="exact", device=0,
                  keep_open=False):
    """For k in range(B):  simulate(problems[problem_index[k]], q; t_end, should_process = false,
                                    collision_model = collision_model(taus[k]..., force of the problem),
                                    process_method = TrackHydrodynamicErrors(problem, false, n_steps, stop_criteria),
                                    initialization_strategy)
    and collect `result.processing_method.df[end]` (trt_magic_parameter.ipynb:30-103) -- as one device batch.

    problems        : list of P problems of ONE shape (type, NX, NY, boundary conditions, delta_t)
    taus            : (B, ntau) relaxation times; SRT: tau; TRT: (tau_symmetric, tau_asymmetric); MRT: tau_n
    problem_index   : (B,) index into `problems` (default: B == P, one problem per solve)
    stop_criteria   : prototype (VelocityConvergenceStoppingCriteria(tol, problem), MeanVelocityStoppingCriteria(0, tol,
                      problem), NoStoppingCriteria()) -- every solve gets its own state; None: StopCriteria(problem)
    forced          : the collision model carries `lattice_force(problem, ...)` when the problem has an external force
    Returns a BatchResult."""
    from .processing_methods import StopCriteria
    if collision_model not in _CM:
        raise TypeError("collision_model must be one of the types SRT, TRT, MRT")
    problems = list(problems)
    taus = np.ascontiguousarray(np.atleast_2d(np.asarray(taus, dtype=np.float64)))
    B = taus.shape[0]
    if problem_index is None:
        if len(problems) != B:
            raise ValueError("without problem_index there must be one problem per row of taus")
        problem_index = np.arange(B)
    pidx = np.asarray(problem_index, dtype=np.int64)
    if pidx.shape != (B,) or pidx.min() < 0 or pidx.max() >= len(problems):
        raise ValueError("problem_index must map every solve to a problem")
    p0 = problems[0]
    bcs = p0.boundary_conditions()
    n_steps = round(t_end / p0.delta_t())
    for pr in problems:
        if (type(pr) is not type(p0) or (pr.NX, pr.NY) != (p0.NX, p0.NY) or _bc_key(pr.boundary_conditions()) != _bc_key(bcs)
                or round(t_end / pr.delta_t()) != n_steps):
            raise ValueError("simulate_many needs problems of one shape: same type, grid, boundary conditions and step count")
    for b in bcs:
        if not isinstance(b, (BounceBack, MovingWall)):
            raise TypeError(f"unsupported boundary condition {b!r}")
    q.check_against_library()
    proto = StopCriteria(p0) if stop_criteria is None else stop_criteria
    stop_kind, tol = _stop_spec(proto)
    strategy = default_strategy(p0) if initialization_strategy is None else initialization_strategy

    batch = _abi.Batch(B, p0.NX, p0.NY, q.name, _CM[collision_model], taus[0], [b.to_abi() for b in bcs],
                       dtype=_DTYPES[dtype], arith=_ARITH[arith], device=device)
    try:
        batch.set_tau(taus)
        has_force = forced and p0.has_external_force()
        if has_force:
            kinds = {LatticeForce(pr).kind() for pr in problems}
            if kinds != {"uniform"}:
                raise ValueError("batched solves support problems with a uniform force (Poiseuille) or none")
            F = np.array([LatticeForce(pr).uniform() for pr in problems], dtype=np.float64)
            batch.set_force_uniform(F[pidx])
        else:
            batch.set_force_uniform(None)
        if isinstance(strategy, ZeroVelocityInitialCondition):
            batch.broadcast_f(initialize(strategy, q, p0, collision_model))  # the same rest state for every problem
        else:
            f_by_problem = [initialize(strategy, q, pr, collision_model) for pr in problems]
            chunk = max(1, (64 << 20) // (f_by_problem[0].size * 8))
            for c0 in range(0, B, chunk):
                idx = pidx[c0:c0 + chunk]
                batch.upload_f(np.stack([f_by_problem[i] for i in idx], axis=3), c0)

        # the loop of simulate(model, 0:n_steps) up to the last row it records: next!(t) evaluates the criterion when
        # mod(t, 100) == 0 and records the row when it fires or t == n_steps (track_hydrodynamic_errors.jl:52-65); the
        # step and the two no-op next! calls after t == n_steps do not touch df
        batch.run(n_steps, stop_kind, check_every, tol)
        timestep, stopped = batch.status()
        device_ms = batch.last_run_ms()

        # TrackHydrodynamicErrors row of every solve at its own t (expected fields of its own problem at t * delta_t)
        cs = q.speed_of_sound_squared
        tau_visc = np.array([cs * pr.lattice_viscosity() for pr in problems])[pidx]
        u_max = np.array([pr.u_max for pr in problems])[pidx]
        coef = np.zeros((B, 8, 3))
        tables = None
        uniq, inv = np.unique(np.stack([pidx, timestep]), axis=1, return_inverse=True)
        ucoef = np.zeros((uniq.shape[1], 8, 3))
        for j in range(uniq.shape[1]):
            pr = problems[int(uniq[0, j])]
            sep = pr.expected_separable(q, int(uniq[1, j]) * pr.delta_t(), 0, pr.NY)
            if sep is None:
                raise ValueError(f"{type(pr).__name__} does not provide its analytic fields in separable form")
            tabs = []
            for f, (c0, terms) in enumerate(sep):
                ucoef[j, f, 0] = c0
                for k, (a, X, Y) in enumerate(terms):
                    ucoef[j, f, 1 + k] = a
                    tabs.append((f, k, None if X is None else np.asarray(X, dtype=np.float64),
                                 None if Y is None else np.asarray(Y, dtype=np.float64)))
            if tables is None:
                tables, sep0 = tabs, sep
            elif not _same_tables(tables, tabs):
                raise ValueError("the problems' expected fields do not share their separable tables")
        coef[:] = ucoef[np.ravel(inv)]
        sums = batch.reduce_errors(tau_visc, u_max, sep0, coef)
    except Exception:
        batch.close()
        raise
    Delta = _delta(p0)
    cols = _rows(sums, Delta)
    res = BatchResult(batch, timestep, stopped, cols, device_ms)
    if not keep_open:
        batch.close()
    return res


def _same_tables(a, b):
    if len(a) != len(b):
        return False
    for (f1, k1, x1, y1), (f2, k2, x2, y2) in zip(a, b):
        if (f1, k1) != (f2, k2):
            return False
        for u, v in ((x1, x2), (y1, y2)):
            if (u is None) != (v is None) or (u is not None and not np.array_equal(u, v)):
                return False
    return True


def _delta(pr):
    # track_hydrodynamic_errors.jl:79-90
    xstep, ystep = pr.range_steps()
    Delta = ystep * xstep
    if pr.NX == 1:
        Delta = ystep
        if pr.NY == 1:
            Delta = 1.0
    elif pr.NY == 1:
        Delta = xstep
    return Delta


def _rows(s, Delta):
    """TrackHydrodynamicErrors._row for (B, 16) sums (track_hydrodynamic_errors.jl:205-221)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        sd = lambda a, b: np.sqrt(a / b)  # noqa: E731
        return dict(error_rho=np.sqrt(s[:, 0]), error_u=sd(s[:, 1], s[:, 2]), error_p=sd(s[:, 3], s[:, 4]),
                    error_sxx=sd(s[:, 5], s[:, 6]), error_sxy=sd(s[:, 7], s[:, 8]), error_syy=sd(s[:, 9], s[:, 10]),
                    error_syx=sd(s[:, 11], s[:, 12]), mass=Delta * s[:, 13], momentum=Delta * s[:, 14],
                    energy=Delta * s[:, 15])


__all__ = ["simulate_many", "BatchResult"]
