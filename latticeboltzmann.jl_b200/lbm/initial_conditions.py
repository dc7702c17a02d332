"""Initialisation strategies of the Julia API (src/initial_conditions{.jl,/*.jl}), host side.

`initialize(strategy, q, problem, cm)` returns `f` of shape (NX, NY, Q), Fortran order, Float64
(== Julia's `Array{Float64,3}(NX, NY, Q)` in memory).  `rows=(y0, ny)` builds only a y-slab.
"""
import numpy as np

from . import vdf
from .problems import TGV


class InitializationStrategy:
    pass


class AnalyticalEquilibrium(InitializationStrategy):
    """analytical_equilibrium.jl:8-17."""


class ConstantDensity(InitializationStrategy):
    """constant_density.jl:8-20."""


class AnalyticalVelocityAndStress(InitializationStrategy):
    """analytical_velocity_stress.jl:4-31."""


class AnalyticalEquilibriumAndOffEquilibrium(InitializationStrategy):
    """analytical_offequilibrium.jl:9-87."""


class ZeroVelocityInitialCondition(InitializationStrategy):
    """initial_conditions.jl:37-53."""


class AnalyticalVelocity(InitializationStrategy):
    """analytical_velocity.jl -- broken in the reference (calls an undefined
    pressure(problem, x, y), :21,39); kept so that the name resolves."""


class IterativeInitializationMeiEtAl(InitializationStrategy):
    """IterativeInitializationMeiEtAl(tau, eps) (mei_et_al.jl:4-10; defaults 1.0, 1e-7).  Runs on the device:
    f = w, then up to 10000 collide-stream steps with the constant-velocity operator
    (IterativeInitializationCollisionModel) until DensityConvergence fires.  `whole_field=True` replaces the reference's
    single-node criterion (see DensityConvergence) by the norm over all nodes; `check_every` > 1 evaluates the
    criterion only every so many steps (the reference checks after every step)."""

    def __init__(self, tau=1.0, eps=1e-7, whole_field=False, check_every=1, max_steps=10000):
        self.tau, self.eps = tau, eps
        self.whole_field, self.check_every, self.max_steps = whole_field, int(check_every), int(max_steps)
        self.steps_taken = None


def IterativeInitialization():
    """mei_et_al.jl:9-10."""
    return IterativeInitializationMeiEtAl(1.0, 1e-7)


def default_strategy(problem):
    """InitializationStrategy(problem) (initial_conditions.jl:5)."""
    return AnalyticalEquilibrium()


def _stack_u(ux, uy):
    return np.stack([ux, uy], axis=-1)


def _equilibrium_problem(q, problem, X, Y):
    # equilibrium(q, problem, x, y): problems.jl:121-128
    rho = problem.lattice_density(q, X, Y)
    ux, uy = problem.lattice_velocity(q, X, Y)
    T = problem.lattice_temperature(q, X, Y)
    return vdf.hermite_based_equilibrium(q, rho, _stack_u(ux, uy), T)


def _dot_H2_sym_grad(q, grad):
    """dot(hermite(Val{2}, c_i, q), grad + grad') for every population -> (..., Q)."""
    (a11, a12), (a21, a22) = grad
    S = np.stack([np.stack([a11 + a11, a12 + a21], -1), np.stack([a21 + a12, a22 + a22], -1)], -2)
    out = np.empty(S.shape[:-2] + (q.Q,))
    for i in range(q.Q):
        H = vdf.hermite(2, (int(q.abscissae[0, i]), int(q.abscissae[1, i])), q)
        out[..., i] = np.tensordot(S, H, axes=([-2, -1], [0, 1]))
    return out


def initialize(strategy, q, problem, cm=None, rows=None):
    """initialize(strategy, q, problem, cm = SRT) (initial_conditions.jl:7-22)."""
    y0, ny = (0, problem.NY) if rows is None else rows
    X, Y = problem.grid(y0, ny)
    one = np.ones_like(X)
    cs = q.speed_of_sound_squared
    if isinstance(strategy, ZeroVelocityInitialCondition):
        f = one[..., None] * q.weights
    elif isinstance(strategy, AnalyticalEquilibrium):
        f = _equilibrium_problem(q, problem, X, Y)
    elif isinstance(strategy, ConstantDensity):
        ux, uy = problem.lattice_velocity(q, X, Y)
        f = vdf.hermite_based_equilibrium(q, one, _stack_u(ux, uy), 1.0)
    elif isinstance(strategy, AnalyticalVelocityAndStress):
        ux, uy = problem.lattice_velocity(q, X, Y)
        f = vdf.hermite_based_equilibrium(q, one, _stack_u(ux, uy), 1.0)
        g = problem.velocity_gradient(X, Y, 0.0)
        g = tuple(tuple(problem.u_max ** 2 * c for c in row) for row in g)
        tau_eff = cs * problem.lattice_viscosity() + 0.5
        f = f - q.weights * (cs * tau_eff * 1.0 * 1.0) / 2 * _dot_H2_sym_grad(q, g)
    elif isinstance(strategy, AnalyticalEquilibriumAndOffEquilibrium):
        f = _equilibrium_problem(q, problem, X, Y)
        tau = cs * problem.lattice_viscosity()
        if isinstance(problem, TGV):  # analytical_offequilibrium.jl:50-87
            rho = problem.lattice_density(q, X, Y)
            g = problem.velocity_gradient(X, Y, 0.0)
            f = f - q.weights * (cs * (tau + 0.5) * rho[..., None] * 1.0) / 2 * _dot_H2_sym_grad(q, g)
        else:  # :10-49
            g = problem.velocity_gradient(X, Y, 0.0)
            g = tuple(tuple(problem.u_max ** 2 * c for c in row) for row in g)
            factor = problem.domain_size[0] * problem.domain_size[1]
            f = f + (-factor * q.weights * 0.5 * ((tau + 0.5) * cs)) * _dot_H2_sym_grad(q, g)
    elif isinstance(strategy, AnalyticalVelocity):
        raise NotImplementedError("AnalyticalVelocity is broken in the reference (analytical_velocity.jl:21,39)")
    elif isinstance(strategy, IterativeInitializationMeiEtAl):
        if rows is not None and rows != (0, problem.NY):
            raise ValueError("the iterative initialisation runs on the whole domain (use initialize_mei_et_al with a comm)")
        return initialize_mei_et_al(strategy, q, problem)
    else:
        raise TypeError(f"unknown initialisation strategy {strategy!r}")
    return np.asfortranarray(f, dtype=np.float64)


def analytic_init_spec(strategy, q, problem, y0=0, ny=None):
    """The arguments of Context.init_analytic for `strategy`: the problem's fields in separable form, found from
    O(NX + NY) evaluations of its pointwise functions (separable.decompose_function), plus the off-equilibrium
    coefficient of the strategy.  None when the strategy has no closed form or some field is not of rank <= 2."""
    from .separable import decompose_function
    xs, ys = problem._xy(y0, ny)
    cs = q.speed_of_sound_squared
    zero = (0.0, [])
    kw = dict(unit_density=False, unit_temperature=False, offeq=0, offeq_coef=0.0)
    grad_scale = None
    if isinstance(strategy, ZeroVelocityInitialCondition):
        return [(1.0, []), zero, zero, (1.0, [])] + [zero] * 4, dict(kw, unit_density=True, unit_temperature=True)
    if isinstance(strategy, AnalyticalEquilibrium):
        pass
    elif isinstance(strategy, ConstantDensity):
        kw.update(unit_density=True, unit_temperature=True)
    elif isinstance(strategy, AnalyticalVelocityAndStress):  # analytical_velocity_stress.jl:5-31
        tau_eff = cs * problem.lattice_viscosity() + 0.5
        kw.update(unit_density=True, unit_temperature=True, offeq=1, offeq_coef=-(cs * tau_eff * 1.0 * 1.0) / 2)
        grad_scale = problem.u_max ** 2
    elif isinstance(strategy, AnalyticalEquilibriumAndOffEquilibrium):  # analytical_offequilibrium.jl:10-87
        tau = cs * problem.lattice_viscosity()
        if isinstance(problem, TGV):
            kw.update(offeq=2, offeq_coef=-(cs * (tau + 0.5) * 1.0) / 2)
            grad_scale = 1.0
        else:
            factor = problem.domain_size[0] * problem.domain_size[1]
            kw.update(offeq=1, offeq_coef=-factor * 0.5 * ((tau + 0.5) * cs))
            grad_scale = problem.u_max ** 2
    else:
        return None
    fns = [lambda X, Y: problem.lattice_density(q, X, Y),
           lambda X, Y: problem.lattice_velocity(q, X, Y)[0],
           lambda X, Y: problem.lattice_velocity(q, X, Y)[1],
           lambda X, Y: problem.pressure(q, X, Y)]
    if kw["unit_density"] and kw["unit_temperature"]:
        fns[0] = fns[3] = None
    if grad_scale is not None:
        for a, b in ((0, 0), (0, 1), (1, 0), (1, 1)):  # du_x/dx, du_x/dy, du_y/dx, du_y/dy
            fns.append(lambda X, Y, a=a, b=b: grad_scale * np.asarray(problem.velocity_gradient(X, Y, 0.0)[a][b]))
    else:
        fns += [None] * 4
    fields = []
    for fn in fns:
        sep = (1.0, []) if fn is None and len(fields) in (0, 3) else zero if fn is None else decompose_function(fn, xs, ys)
        if sep is None:
            return None
        fields.append(sep)
    return fields, kw


def initialize_on_device(strategy, q, problem, ctx, chunk_nodes=1 << 24, analytic=True):
    """initialize(strategy, q, problem) evaluated on the device.  Every closed-form strategy (equilibrium, constant
    density, velocity + stress, equilibrium + off-equilibrium, zero velocity) goes through lbm_init_analytic: the host
    passes the problem's rho, u, p and grad u as separable tables (O(NX + NY) numbers) and the kernel evaluates the
    Hermite-series equilibrium and the off-equilibrium part per node.  Problems whose fields are not of rank <= 2 fall
    back, for the equilibrium-only strategies, to (rho, u, T) row blocks from the host (lbm_init_equilibrium_rows).
    Returns False when the strategy cannot be evaluated on the device."""
    spec = analytic_init_spec(strategy, q, problem, ctx.y0, ctx.ny_local) if analytic else None
    if spec is not None:
        ctx.init_analytic(spec[0], **spec[1])
        return True
    if not isinstance(strategy, (ZeroVelocityInitialCondition, AnalyticalEquilibrium, ConstantDensity)):
        return False
    rows = max(1, min(ctx.ny_local, chunk_nodes // max(problem.NX, 1)))
    for off in range(0, ctx.ny_local, rows):
        n = min(rows, ctx.ny_local - off)
        X, Y = problem.grid(ctx.y0 + off, n)
        one = np.ones_like(X)
        if isinstance(strategy, ZeroVelocityInitialCondition):
            rho, ux, uy, T = one, 0 * one, 0 * one, one
        elif isinstance(strategy, ConstantDensity):
            ux, uy = problem.lattice_velocity(q, X, Y)
            rho, T = one, one
        else:
            rho = problem.lattice_density(q, X, Y)
            ux, uy = problem.lattice_velocity(q, X, Y)
            T = problem.lattice_temperature(q, X, Y)
        ctx.init_equilibrium_rows(off, rho, ux, uy, T)
    return True


def initialize_mei_et_al(strategy, q, problem, dtype="f64", arith="exact", comm=None, device=None, model_out=None):
    """initialize(::IterativeInitializationMeiEtAl, q, problem) (mei_et_al.jl:11-40) on the device.  Returns the
    local slab of f_stream; `strategy.steps_taken` records how many collide-stream steps ran."""
    from .collision_models import IterativeInitializationCollisionModel
    from .model import LatticeBoltzmannModel, simulate_model
    from .processing_methods import ProcessIterativeInitialization

    class _EveryK(ProcessIterativeInitialization):
        def noop(self, t, k=strategy.check_every):
            return k > 1 and (t - 1) % k != 0

    cm = IterativeInitializationCollisionModel(q, strategy.tau, problem)
    pm = _EveryK(strategy.eps, problem, None, strategy.whole_field)
    model = LatticeBoltzmannModel(problem, q, collision_model=cm, initialization_strategy=ZeroVelocityInitialCondition(),
                                  process_method=pm, dtype=dtype, arith=arith, comm=comm, device=device)
    try:
        simulate_model(model, range(1, strategy.max_steps + 1))
        strategy.steps_taken = model.state.steps_done
        return model.f_stream
    finally:
        model.close()
