"""Driver of the Julia API (src/lattice_boltzmann_model.jl) on top of liblbm_b200.so.

`LatticeBoltzmannModel` owns one device context; `simulate(model, time)` keeps the reference's
loop semantics (`collide!(time = t*dt)`, `stream!`, `apply!`, `next!(model, t + 1)`, trailing
`next!`), but consecutive steps whose `next!` is a no-op are issued as one fused device batch.
"""
import atexit
import collections

import numpy as np

from . import _abi
from .boundary_conditions import BounceBack, MovingWall
from .collision_models import (MRT, SRT, TRT, CollisionModel, IterativeInitializationCollisionModel, LatticeForce)
from .initial_conditions import default_strategy, initialize, initialize_on_device
from .processing_methods import ProcessingMethod

_DTYPES = {"f64": _abi.F64, "float64": _abi.F64, np.float64: _abi.F64, "f32": _abi.F32, "float32": _abi.F32,
           np.float32: _abi.F32, _abi.F64: _abi.F64, _abi.F32: _abi.F32}
_ARITH = {"exact": _abi.ARITH_EXACT, "fast": _abi.ARITH_FAST, 0: 0, 1: 1}


def _cm_code(cm):
    if isinstance(cm, SRT):
        return _abi.SRT
    if isinstance(cm, TRT):
        return _abi.TRT
    if isinstance(cm, MRT):
        return _abi.MRT
    if isinstance(cm, IterativeInitializationCollisionModel):
        return _abi.ITERATIVE_INIT
    raise TypeError(f"not a collision model: {cm!r}")


def make_context(q, cm, bcs, nx, ny, dtype="f64", arith="exact", comm=None, device=None):
    q.check_against_library()
    for bc in bcs:
        if not isinstance(bc, (BounceBack, MovingWall)):
            raise TypeError(f"unsupported boundary condition {bc!r}")
    kw = {}
    if comm is not None and comm.world > 1:
        kw = dict(rank=comm.rank, world=comm.world, nccl_id=comm.nccl_id())
    dev = device if device is not None else (comm.device if comm is not None else 0)
    return _abi.Context(nx, ny, q.name, _cm_code(cm), cm.taus(), [b.to_abi() for b in bcs], dtype=_DTYPES[dtype],
                        arith=_ARITH[arith], device=dev, **kw)


class DeviceState:
    """Device-resident populations + the collision model's force data."""

    SEPARABLE_WINDOW = 2048  # lattice steps per separable force table (DecayingShearFlow(static), decaying_shear_flow.jl:131-147)

    def __init__(self, ctx, q, cm, comm=None):
        self.ctx, self.q, self.cm, self.comm = ctx, q, cm, comm
        self.y0, self.ny_local, self.nx, self.ny = ctx.y0, ctx.ny_local, ctx.nx, ctx.ny
        self._force_window = None
        self._static_force_set = False
        self.steps_done = 0

    # -- force -------------------------------------------------------------------------------
    def max_batch(self):
        f = self.cm.force
        if f is None:
            return 1 << 30
        if isinstance(f, LatticeForce):
            # a separable time-dependent table holds (NX + NY_local) values per step on host and device: bounded window
            return self.SEPARABLE_WINDOW if f.kind() == "separable" else 1 << 30
        return 1  # opaque host closure: re-evaluated every step

    def prepare_force(self, t0, n, dt):
        if isinstance(self.cm, IterativeInitializationCollisionModel):
            if not self._static_force_set:
                self.ctx.set_velocity_field(*self.cm.velocity_field(self.y0, self.ny_local))
                self._static_force_set = True
            return
        f = self.cm.force
        if f is None:
            if not self._static_force_set:
                self.ctx.set_force_none()
                self._static_force_set = True
            return
        if isinstance(f, LatticeForce):
            kind = f.kind()
            if kind == "uniform":
                if not self._static_force_set:
                    self.ctx.set_force_uniform(*f.uniform())
                    self._static_force_set = True
            elif kind == "field":
                if not self._static_force_set:
                    self.ctx.set_force_field(*f.field(0.0, self.y0, self.ny_local))
                    self._static_force_set = True
            else:
                w = self._force_window
                if w is None or t0 < w[0] or t0 + n > w[1]:
                    m = max(n, 1)
                    fx, fy = f.separable(t0, m, self.y0, self.ny_local)
                    self.ctx.set_force_separable(t0, fx, fy)
                    self._force_window = (t0, t0 + m)
            return
        self._closure_force(f, t0 * dt)

    def _closure_force(self, f, time):
        """generic closure (x_idx, y_idx, t) -> [Fx, Fy], 1-based indices, evaluated on the host"""
        Fx = np.empty((self.nx, self.ny_local))
        Fy = np.empty((self.nx, self.ny_local))
        for x in range(self.nx):
            for y in range(self.ny_local):
                F = f(x + 1, self.y0 + y + 1, time)
                Fx[x, y], Fy[x, y] = F[0], F[1]
        self.ctx.set_force_field(Fx, Fy)

    def prepare_force_time(self, time):
        """Force data for ONE collide at an arbitrary `time` (step index 0)."""
        f = self.cm.force
        if isinstance(f, LatticeForce) and f.kind() == "separable":
            fx, fy = f.problem.force_separable_times([time], self.y0, self.ny_local)
            self.ctx.set_force_separable(0, fx, fy)
            self._force_window = None
        elif f is None or isinstance(f, LatticeForce):
            self.prepare_force(0, 1, 0.0)
        else:
            self._closure_force(f, time)

    # -- stepping ----------------------------------------------------------------------------
    def step(self, t0, n, dt):
        done = 0
        mb = self.max_batch()
        while done < n:
            m = min(mb, n - done)
            self.prepare_force(t0 + done, m, dt)
            self.ctx.step(t0 + done, m, dt)
            done += m
        self.steps_done = getattr(self, "steps_done", 0) + n

    # -- diagnostics -------------------------------------------------------------------------
    def moments(self, tau_visc, fields):
        return self.ctx.moments(tau_visc, fields)

    def allreduce(self, values):
        if self.comm is None or self.comm.world == 1:
            return np.asarray(values, dtype=np.float64)
        return self.comm.allreduce_sum(values)

    def reduce(self, kind):
        return self.allreduce(self.ctx.reduce(kind))

    def download_f(self):
        return self.ctx.download_f()


class LatticeBoltzmannModel:
    """LatticeBoltzmannModel(problem, quadrature; collision_model = SRT, initialization_strategy,
    process_method) (lattice_boltzmann_model.jl:15-33).  Extra, B200-only keywords: `dtype`
    ("f64" | "f32"), `arith` ("exact" | "fast"), `comm` (SlabComm for y-slab multi-GPU)."""

    def __init__(self, problem, quadrature, collision_model=SRT, initialization_strategy=None, process_method=None,
                 dtype="f64", arith="exact", comm=None, device=None, device_init=False):
        strategy = default_strategy(problem) if initialization_strategy is None else initialization_strategy
        cm = CollisionModel(collision_model, quadrature, problem)
        bcs = problem.boundary_conditions()
        ctx = make_context(quadrature, cm, bcs, problem.NX, problem.NY, dtype, arith, comm, device)
        self._init(ctx, quadrature, cm, bcs, process_method, comm)
        if not (device_init and initialize_on_device(strategy, quadrature, problem, ctx)):
            ctx.upload_f(initialize(strategy, quadrature, problem, collision_model, rows=(ctx.y0, ctx.ny_local)))

    @classmethod
    def from_fields(cls, f_stream, f_collision, quadrature, collision_model, boundary_conditions, processing_method,
                    dtype="f64", arith="exact", device=None):
        """Positional constructor (lattice_boltzmann_model.jl:1-14)."""
        f_stream = np.asfortranarray(f_stream, dtype=np.float64)
        nx, ny, _ = f_stream.shape
        m = cls.__new__(cls)
        ctx = make_context(quadrature, collision_model, boundary_conditions, nx, ny, dtype, arith, None, device)
        m._init(ctx, quadrature, collision_model, list(boundary_conditions), processing_method, None)
        ctx.upload_f(f_stream)
        return m

    def _init(self, ctx, q, cm, bcs, pm, comm):
        self.ctx = ctx
        self.quadrature = q
        self.collision_model = cm
        self.boundary_conditions = bcs
        self.processing_method = pm
        self.state = DeviceState(ctx, q, cm, comm)

    @property
    def f_stream(self):
        return self.ctx.download_f()

    @f_stream.setter
    def f_stream(self, f):
        self.ctx.upload_f(f)

    @property
    def f_collision(self):
        return self.ctx.download_f_collision()

    def close(self):
        self.ctx.close()


def _dt_of(model):
    pm = model.processing_method
    pr = getattr(pm, "problem", None)
    return pr.delta_t() if pr is not None else 0.0


def collide_model_(model, time=0.0):
    """collide!(model; time) (lattice_boltzmann_model.jl:84-92)."""
    model.state.prepare_force_time(time)
    model.ctx.collide(0, time)


def stream_model_(model):
    """stream!(model) (:94-96)."""
    model.ctx.stream()


def apply_boundary_conditions_(model, time=0.0):
    """apply_boundary_conditions!(model; time) (:98-106)."""
    model.ctx.apply_bcs(time)


def next_model_(model, t):
    """next!(model, t) (:108-110)."""
    pm = model.processing_method
    if pm is None:
        return False
    return bool(pm.next_(model.quadrature, model.state, t))


def simulate_model(model, time):
    """simulate(model, time) (lattice_boltzmann_model.jl:60-77)."""
    dt = _dt_of(model)
    pm = model.processing_method
    time = list(time)
    batch_t0, batch_n = None, 0

    def flush():
        nonlocal batch_t0, batch_n
        if batch_n:
            model.state.step(batch_t0, batch_n, dt)
        batch_t0, batch_n = None, 0

    for t in time:
        if batch_n and t != batch_t0 + batch_n:
            flush()
        if batch_n == 0:
            batch_t0 = t
        batch_n += 1
        if pm is None or pm.noop(t + 1):
            continue
        flush()
        if next_model_(model, t + 1):
            return model
    flush()
    if time:
        next_model_(model, time[-1] + 1)
    return model


def simulate(problem_or_model, q_or_time=None, *, process_method=None, should_process=True,
             initialization_strategy=None, t_end=1.0, collision_model=SRT, **device_kw):
    """simulate(problem, q; ...) (lattice_boltzmann_model.jl:34-59) or simulate(model, time) (:60-77)."""
    if isinstance(problem_or_model, LatticeBoltzmannModel):
        return simulate_model(problem_or_model, q_or_time)
    problem, q = problem_or_model, q_or_time
    dt = problem.delta_t()
    n_steps = round(t_end / dt)  # Julia round(Int, x): half to even, like Python's round
    if process_method is None:
        process_method = ProcessingMethod(problem, should_process, n_steps)
    model = LatticeBoltzmannModel(problem, q, collision_model=collision_model,
                                  initialization_strategy=initialization_strategy, process_method=process_method,
                                  **device_kw)
    return simulate_model(model, range(0, n_steps + 1))


# ---------------------------------------------------------------------------------------------
# array-level operators (the generic functions the reference's tests and benchmarks call)
# ---------------------------------------------------------------------------------------------
# The reference's tests and benchmarks call collide!/stream!/apply! on plain arrays, often in a loop.  A device context
# per call would cost more than the operator (allocation, streams, constant upload), so the scratch contexts of the most
# recent (lattice, model, relaxation times, boundary conditions, shape, dtype) combinations are kept alive.
_SCRATCH_CACHE = collections.OrderedDict()
_SCRATCH_CACHE_SIZE = 8


def _scratch_key(q, cm, bcs, nx, ny, dtype, arith):
    if not isinstance(cm, (SRT, TRT, MRT)):
        return None  # (operators tied to a problem instance are not shared)
    bkey = tuple((type(b).__name__, type(b.direction).__name__, tuple(b.xs), tuple(b.ys),
                  tuple(np.ravel(getattr(b, "u", ()))), getattr(b, "rho", None)) for b in bcs)
    return (q.name, type(cm).__name__, tuple(cm.taus()), bkey, nx, ny, str(dtype), str(arith))


def _scratch(q, cm, bcs, f, dtype="f64", arith="exact"):
    f = np.asfortranarray(f, dtype=np.float64)
    nx, ny, Q = f.shape
    if Q != q.Q:
        raise ValueError(f"{q.name} has {q.Q} populations, array has {Q}")
    try:
        key = _scratch_key(q, cm, bcs, nx, ny, dtype, arith)
    except Exception:
        key = None
    if key is not None and key in _SCRATCH_CACHE:
        _SCRATCH_CACHE.move_to_end(key)
        return _SCRATCH_CACHE[key], f
    ctx = make_context(q, cm, bcs, nx, ny, dtype, arith)
    if key is not None:
        _SCRATCH_CACHE[key] = ctx
        while len(_SCRATCH_CACHE) > _SCRATCH_CACHE_SIZE:
            _, old = _SCRATCH_CACHE.popitem(last=False)
            old.close()
    return ctx, f


def _direct_out(ctx, f_out):
    """f_out if the library can write the result straight into it (Fortran-ordered float64 of the context's shape), else
    None: the result is then downloaded into a new array and assigned -- one more pass over host memory"""
    if (isinstance(f_out, np.ndarray) and f_out.dtype == np.float64 and f_out.shape == ctx.shape and f_out.flags.f_contiguous
            and f_out.flags.writeable):
        return f_out
    return None


def _release(ctx):
    """end of an array-level call: cached scratch contexts stay alive, the others are destroyed"""
    if not any(c is ctx for c in _SCRATCH_CACHE.values()):
        ctx.close()


def clear_scratch_contexts():
    while _SCRATCH_CACHE:
        _, ctx = _SCRATCH_CACHE.popitem()
        ctx.close()


atexit.register(clear_scratch_contexts)


def collide_(collision_model, q, f_in=None, f_out=None, *, time=0.0, f_old=None, f_new=None, dtype="f64",
             arith="exact"):
    """collide!(cm, q, f_in, f_out; time) and the keyword form collide!(cm, q; time, f_new, f_old)
    (collision_models.jl:19; srt.jl:18, trt.jl:42, mrt.jl:56)."""
    f_in = f_old if f_in is None else f_in
    f_out = f_new if f_out is None else f_out
    ctx, f = _scratch(q, collision_model, [], f_in, dtype, arith)
    try:
        st = DeviceState(ctx, q, collision_model)
        ctx.upload_f(f)
        st.prepare_force_time(time)
        ctx.collide(0, time)
        direct = _direct_out(ctx, f_out)
        out = ctx.download_f_collision(direct)
    finally:
        _release(ctx)
    if f_out is not None and direct is None:
        f_out[...] = out
    return out


def stream_(q, f=None, f_new=None, *, f_old=None, dtype="f64"):
    """stream!(q, f, f_new) / stream!(q; f_new, f_old) (stream.jl:18-30): periodic pull."""
    f = f_old if f is None else f
    ctx, f = _scratch(q, SRT(1.0), [], f, dtype)
    try:
        ctx.upload_f_collision(f)
        ctx.stream()
        direct = _direct_out(ctx, f_new)
        out = ctx.download_f(direct)
    finally:
        _release(ctx)
    if f_new is not None and direct is None:
        f_new[...] = out
    return out


def stream(q, f, f_new=None, *, dtype="f64"):
    """stream(q, f, f_new = copy(f)) -- the scatter ("push") variant (stream.jl:6-16, 44-61).  It wraps an index at most
    once (`if next_x > lx ... elseif next_x < 1`), so it is only defined for grids at least as large as the widest
    lattice velocity (smaller ones index out of bounds under the reference's `@inbounds`); there it is the same
    permutation as the periodic pull and runs on the same kernel.  Returns f_new."""
    f = np.asarray(f)
    h = int(np.abs(q.abscissae).max())
    if f.shape[0] < h or f.shape[1] < h:
        raise ValueError(f"push streaming needs a grid of at least {h} x {h} nodes for {q.name} (single wrap, stream.jl:44-61)")
    out = stream_(q, f, dtype=dtype)
    if f_new is not None:
        f_new[...] = out
        return f_new
    return out


def apply_(bcs, q, f_new, f_old, *, time=0.0, dtype="f64"):
    """apply!(bcs, q, f_new, f_old; time) (boundary_conditions.jl:6-16): in place on f_new."""
    if not isinstance(bcs, (list, tuple)):
        bcs = [bcs]
    ctx, fo = _scratch(q, SRT(1.0), list(bcs), f_old, dtype)
    try:
        ctx.upload_f(np.asfortranarray(f_new, dtype=np.float64))
        ctx.upload_f_collision(fo)
        ctx.apply_bcs(time)
        direct = _direct_out(ctx, f_new)
        out = ctx.download_f(direct)
    finally:
        _release(ctx)
    if direct is None:
        f_new[...] = out
    return f_new
