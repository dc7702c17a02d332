"""Collision-model types of the Julia API (src/collision_models{.jl,/*.jl}).

The reference stores the force as a Julia closure `(x_idx, y_idx, t) -> [Fx, Fy]`.  Here
`force` may be: None; such a callable (1-based indices; evaluated on the host for every node
and step -- slow path); or a `LatticeForce(problem)` built by the `CollisionModel` factory,
which is also callable but additionally tells the device path whether the force is uniform,
a static field, or a separable time-dependent table.
"""
import numpy as np

from .problems import DecayingShearFlow, lattice_force
from .quadratures import order


class LatticeForce:
    """`(x_idx, y_idx, t) -> lattice_force(problem, x_idx, y_idx, t)` (srt.jl:11, trt.jl:17-18, mrt.jl:41)."""

    def __init__(self, problem):
        self.problem = problem

    def __call__(self, x_idx, y_idx, t=0.0):
        return lattice_force(self.problem, x_idx, y_idx, t)

    def kind(self):
        p = self.problem
        if hasattr(p, "force_uniform"):
            return "uniform"
        if isinstance(p, DecayingShearFlow):
            return "separable"
        return "field"  # static in every shipped problem (taylor_green_vortex.jl:113-116)

    def uniform(self):
        return self.problem.force_uniform()

    def field(self, t, y0, ny):
        Fx, Fy = self.problem.force_on_grid(t, y0, ny)
        s = self.problem.u_max * self.problem.delta_t()
        return s * Fx, s * Fy

    def separable(self, t0, nsteps, y0, ny):
        return self.problem.force_separable(t0, nsteps, y0, ny)


class CollisionModelBase:
    force = None


class SRT(CollisionModelBase):
    """SRT(tau[, force])  (srt.jl:1-5)."""

    def __init__(self, tau, force=None):
        self.tau = float(tau)
        self.force = force

    def taus(self):
        return [self.tau]


class IterativeInitializationCollisionModel(CollisionModelBase):
    """IterativeInitializationCollisionModel(q, tau, problem) (collision_models/iterative_initialization.jl:1-41):
    SRT towards an equilibrium whose velocity is pinned to `lattice_velocity(q, problem, x, y)`; only the density comes
    from f.  The reference stores `nonlinear_term[x, y, i]`; here the device evaluates it from the velocity field."""

    def __init__(self, q, tau, problem):
        self.tau = float(tau)
        self.q, self.problem = q, problem

    def taus(self):
        return [self.tau]

    def velocity_field(self, y0=0, ny=None):
        X, Y = self.problem.grid(y0, self.problem.NY if ny is None else ny)
        return self.problem.lattice_velocity(self.q, X, Y)


class TRT(CollisionModelBase):
    """TRT(tau_symmetric, tau_asymmetric, force) -- and the 2-argument convenience
    constructor TRT(tau_a, tau_s) with SWAPPED order (trt.jl:1-6)."""

    def __init__(self, a, b, *force):
        if force:
            self.tau_symmetric, self.tau_asymmetric, self.force = float(a), float(b), force[0]
        else:
            self.tau_symmetric, self.tau_asymmetric, self.force = float(b), float(a), None

    def taus(self):
        return [self.tau_symmetric, self.tau_asymmetric]


class TRT_Lambda:
    """TRT_Λ(Λ): TRT factory with magic parameter (trt.jl:35-40)."""

    def __init__(self, Lambda):
        self.Lambda = Lambda


class MRT(CollisionModelBase):
    """MRT(q, tau | (tau_s, tau_a) | taus[, force])  (mrt.jl:1-34).  As in the reference, the
    scalar forms DROP `force` (mrt.jl:19-27)."""

    def __init__(self, q, *args):
        N = round(order(q) / 2)
        force = None
        if len(args) >= 1 and np.ndim(args[0]) == 1:
            taus = [float(t) for t in args[0]]
            force = args[1] if len(args) > 1 else None
        elif len(args) >= 2 and np.isscalar(args[1]) and not callable(args[1]) and args[1] is not None:
            taus = ([float(args[0]), float(args[1])] * N)[:N]
        else:
            taus = [float(args[0])] * N
        self.tau_list = taus
        self.force = force
        self.q = q

    def taus(self):
        return list(self.tau_list)


def CollisionModel(cm, q, problem, Lambda=1 / 4):
    """CollisionModel(cm, q, problem): type -> factory, instance -> itself
    (collision_models.jl:11-18, srt.jl:7-16, trt.jl:8-21, mrt.jl:36-47)."""
    if isinstance(cm, TRT_Lambda):
        return CollisionModel(TRT, q, problem, Lambda=cm.Lambda)
    if isinstance(cm, CollisionModelBase):
        return cm
    tau = q.speed_of_sound_squared * problem.lattice_viscosity() + 0.5
    force = LatticeForce(problem) if problem.has_external_force() else None
    if cm is TRT:
        return TRT(tau, 0.5 + Lambda / (tau - 0.5), force)
    if cm is MRT:
        return MRT(q, [tau] * order(q), force)
    return SRT(tau, force)  # default (collision_models.jl:11-17)
