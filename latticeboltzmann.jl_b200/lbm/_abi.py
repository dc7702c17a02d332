"""ctypes binding of liblbm_b200.so (C ABI: include/lbm_b200.h).

This is the exact surface a Julia `ccall` shim binds (see INTEGRATION.md and
latticeboltzmann.jl_b200/julia/LatticeBoltzmannB200.jl); Python is the host language
here only because Julia is not installed in the build image.

There is NO CPU fallback: importing works without a GPU (so that the ABI can be
inspected), but every compute entry point needs the CUDA library and a device, and
raises `LbmError` otherwise.
"""
import ctypes as C
import os

import numpy as np

LBM_ABI_VERSION = 1
LBM_MAX_Q = 37
LBM_MAX_TAU = 16
LBM_MAX_BCS = 8
LBM_NCCL_ID_BYTES = 128

# enums (include/lbm_b200.h)
LATTICE_IDS = {"D2Q4": 0, "D2Q5": 1, "D2Q9": 2, "D2Q13": 3, "D2Q17": 4, "D2Q21": 5, "D2Q37": 6}
F64, F32 = 0, 1
SRT, TRT, MRT, ITERATIVE_INIT = 0, 1, 2, 3
ARITH_EXACT, ARITH_FAST = 0, 1
BC_BOUNCE_BACK, BC_MOVING_WALL = 0, 1
NORTH, EAST, SOUTH, WEST = 0, 1, 2, 3
REDUCE_MEAN_UX, REDUCE_VELOCITY_CHANGE, REDUCE_CONSERVED, REDUCE_DENSITY_CHANGE = 0, 1, 2, 3

EXPORTS = [
    "lbm_abi_version", "lbm_last_error", "lbm_lattice_info", "lbm_nccl_unique_id", "lbm_create",
    "lbm_destroy", "lbm_local_rows", "lbm_upload_f", "lbm_upload_f_collision", "lbm_download_f",
    "lbm_upload_f_rows", "lbm_download_f_rows", "lbm_init_equilibrium_rows",
    "lbm_download_f_collision",
    "lbm_set_force_none", "lbm_set_force_uniform", "lbm_set_force_field", "lbm_set_velocity_field",
    "lbm_set_force_separable",
    "lbm_collide", "lbm_stream", "lbm_apply_bcs", "lbm_step", "lbm_sync", "lbm_moments", "lbm_reduce",
    "lbm_reduce_errors", "lbm_reduce_process",
    "lbm_kernel_launches", "lbm_halo_path", "lbm_last_step_ms", "lbm_timer_start", "lbm_timer_stop", "lbm_set_option",
    "lbm_init_analytic", "lbm_host_alloc", "lbm_host_free", "lbm_upload_f_async", "lbm_download_f_async", "lbm_snapshot_begin", "lbm_snapshot_end",
    "lbm_batch_create", "lbm_batch_destroy", "lbm_batch_set_tau", "lbm_batch_set_force_uniform", "lbm_batch_upload_f",
    "lbm_batch_broadcast_f", "lbm_batch_download_f", "lbm_batch_run", "lbm_batch_status", "lbm_batch_reduce_errors",
    "lbm_batch_last_run_ms", "lbm_batch_kernel_launches",
]
BATCH_STOP_OFF, BATCH_STOP_MEAN_VELOCITY, BATCH_STOP_VELOCITY_CONVERGENCE = 0, 1, 2


class LbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblbm_b200 error {code}: {msg}")
        self.code = code


class lbm_bc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("direction", C.c_int32),
                ("x0", C.c_int32), ("x1", C.c_int32), ("y0", C.c_int32), ("y1", C.c_int32),
                ("u", C.c_double * 2), ("rho", C.c_double), ("T", C.c_double)]


class lbm_sep_field(C.Structure):
    _fields_ = [("c0", C.c_double), ("a", C.c_double * 2), ("x", C.c_void_p * 2), ("y", C.c_void_p * 2)]


class lbm_init_spec(C.Structure):
    _fields_ = [("rho", lbm_sep_field), ("ux", lbm_sep_field), ("uy", lbm_sep_field), ("p", lbm_sep_field),
                ("grad", lbm_sep_field * 4), ("unit_density", C.c_int32), ("unit_temperature", C.c_int32),
                ("offeq", C.c_int32), ("offeq_coef", C.c_double)]


class lbm_batch_stop(C.Structure):
    _fields_ = [("kind", C.c_int32), ("check_every", C.c_int32), ("tolerance", C.c_double)]


class lbm_desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
                ("lattice", C.c_int32), ("dtype", C.c_int32), ("collision", C.c_int32),
                ("arith", C.c_int32), ("ntau", C.c_int32), ("tau", C.c_double * LBM_MAX_TAU),
                ("n_bcs", C.c_int32), ("bcs", lbm_bc * LBM_MAX_BCS),
                ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("nccl_id", C.c_uint8 * LBM_NCCL_ID_BYTES)]


_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("LBM_B200_LIB", os.path.join(_PKG, "liblbm_b200.so"))
_lib = None


def lib():
    """Loads liblbm_b200.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LbmError(-1, f"{LIB_PATH} not found: build it with `make -C latticeboltzmann.jl_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    l = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    l.lbm_abi_version.restype = C.c_int
    l.lbm_last_error.restype = C.c_char_p
    l.lbm_lattice_info.argtypes = [C.c_int32, ip, ip, ip, dp, dp, ip, ip, ip, ip]
    l.lbm_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    l.lbm_create.argtypes = [C.POINTER(lbm_desc), C.POINTER(vp)]
    l.lbm_destroy.argtypes = [vp]
    l.lbm_destroy.restype = None
    l.lbm_local_rows.argtypes = [vp, ip, ip]
    l.lbm_upload_f.argtypes = [vp, vp]
    l.lbm_upload_f_collision.argtypes = [vp, vp]
    l.lbm_download_f.argtypes = [vp, vp]
    l.lbm_upload_f_rows.argtypes = [vp, C.c_int32, C.c_int32, vp]
    l.lbm_download_f_rows.argtypes = [vp, C.c_int32, C.c_int32, vp]
    l.lbm_init_equilibrium_rows.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
    l.lbm_download_f_collision.argtypes = [vp, vp]
    l.lbm_set_force_none.argtypes = [vp]
    l.lbm_set_force_uniform.argtypes = [vp, C.c_double, C.c_double]
    l.lbm_set_force_field.argtypes = [vp, vp]
    l.lbm_set_velocity_field.argtypes = [vp, vp]
    l.lbm_set_force_separable.argtypes = [vp, C.c_int64, C.c_int32, vp, vp]
    l.lbm_collide.argtypes = [vp, C.c_int64, C.c_double]
    l.lbm_stream.argtypes = [vp]
    l.lbm_apply_bcs.argtypes = [vp, C.c_double]
    l.lbm_step.argtypes = [vp, C.c_int64, C.c_int64, C.c_double]
    l.lbm_sync.argtypes = [vp]
    l.lbm_moments.argtypes = [vp, C.c_double] + [vp] * 8
    l.lbm_reduce.argtypes = [vp, C.c_int32, dp, C.c_int32]
    l.lbm_reduce_errors.argtypes = [vp, C.c_double, C.c_double, C.POINTER(lbm_sep_field), dp]
    l.lbm_reduce_process.argtypes = [vp, C.c_double, C.POINTER(lbm_sep_field), dp]
    l.lbm_kernel_launches.argtypes = [vp]
    l.lbm_kernel_launches.restype = C.c_int64
    l.lbm_halo_path.argtypes = [vp]
    l.lbm_last_step_ms.argtypes = [vp, C.POINTER(C.c_float)]
    l.lbm_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    l.lbm_timer_start.argtypes = [vp]
    l.lbm_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    l.lbm_init_analytic.argtypes = [vp, C.POINTER(lbm_init_spec)]
    l.lbm_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    l.lbm_host_free.argtypes = [vp]
    l.lbm_upload_f_async.argtypes = [vp, vp]
    l.lbm_download_f_async.argtypes = [vp, vp]
    l.lbm_snapshot_begin.argtypes = [vp, vp]
    l.lbm_snapshot_end.argtypes = [vp]
    l.lbm_batch_create.argtypes = [C.POINTER(lbm_desc), C.c_int32, C.POINTER(vp)]
    l.lbm_batch_destroy.argtypes = [vp]
    l.lbm_batch_destroy.restype = None
    l.lbm_batch_set_tau.argtypes = [vp, vp]
    l.lbm_batch_set_force_uniform.argtypes = [vp, vp]
    l.lbm_batch_upload_f.argtypes = [vp, C.c_int32, C.c_int32, vp]
    l.lbm_batch_broadcast_f.argtypes = [vp, vp]
    l.lbm_batch_download_f.argtypes = [vp, C.c_int32, C.c_int32, vp]
    l.lbm_batch_run.argtypes = [vp, C.c_int64, C.POINTER(lbm_batch_stop)]
    l.lbm_batch_status.argtypes = [vp, C.c_int32, C.c_int32, vp, vp]
    l.lbm_batch_reduce_errors.argtypes = [vp, vp, vp, C.POINTER(lbm_sep_field), vp, vp]
    l.lbm_batch_last_run_ms.argtypes = [vp, C.POINTER(C.c_float)]
    l.lbm_batch_kernel_launches.argtypes = [vp]
    l.lbm_batch_kernel_launches.restype = C.c_int64
    if l.lbm_abi_version() != LBM_ABI_VERSION:
        raise LbmError(-1, f"ABI version mismatch: library {l.lbm_abi_version()} != binding {LBM_ABI_VERSION}")
    _lib = l
    return l


def check(rc):
    if rc != 0:
        raise LbmError(rc, lib().lbm_last_error().decode(errors="replace"))


def lattice_info(lattice_id):
    """Built-in tables of one quadrature (lbm_lattice_info)."""
    q = C.c_int32()
    cx = (C.c_int32 * LBM_MAX_Q)()
    cy = (C.c_int32 * LBM_MAX_Q)()
    opp = (C.c_int32 * LBM_MAX_Q)()
    w = (C.c_double * LBM_MAX_Q)()
    css = C.c_double()
    eq_order, n, halo = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib().lbm_lattice_info(lattice_id, C.byref(q), cx, cy, w, C.byref(css), opp, C.byref(eq_order),
                                 C.byref(n), C.byref(halo)))
    Q = q.value
    return dict(Q=Q, cx=np.array(cx[:Q]), cy=np.array(cy[:Q]), w=np.array(w[:Q]), css=css.value,
                opposite=np.array(opp[:Q]), eq_order=eq_order.value, hermite_order=n.value, halo=halo.value)


def nccl_unique_id():
    buf = (C.c_uint8 * LBM_NCCL_ID_BYTES)()
    check(lib().lbm_nccl_unique_id(buf))
    return bytes(buf)


def _as_f64(a, shape=None):
    a = np.asarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def make_desc(nx, ny, lattice, collision, tau, bcs=(), dtype=F64, arith=ARITH_EXACT, device=0, rank=0, world=1,
              nccl_id=None):
    """lbm_desc: everything `LatticeBoltzmannModel(problem, q; collision_model, ...)` fixes."""
    d = lbm_desc()
    d.abi_version = LBM_ABI_VERSION
    d.nx, d.ny = int(nx), int(ny)
    d.lattice = LATTICE_IDS[lattice] if isinstance(lattice, str) else int(lattice)
    d.dtype, d.collision, d.arith = int(dtype), int(collision), int(arith)
    tau = [float(t) for t in np.atleast_1d(tau)]
    if len(tau) > LBM_MAX_TAU:
        raise ValueError("too many relaxation times")
    d.ntau = len(tau)
    for i, t in enumerate(tau):
        d.tau[i] = t
    bcs = list(bcs)
    if len(bcs) > LBM_MAX_BCS:
        raise ValueError(f"at most {LBM_MAX_BCS} boundary conditions")
    d.n_bcs = len(bcs)
    for i, b in enumerate(bcs):
        d.bcs[i] = b
    d.device, d.rank, d.world = int(device), int(rank), int(world)
    if world > 1:
        if nccl_id is None or len(nccl_id) != LBM_NCCL_ID_BYTES:
            raise ValueError("world > 1 needs the 128-byte NCCL id from rank 0")
        C.memmove(d.nccl_id, nccl_id, LBM_NCCL_ID_BYTES)
    return d


class Context:
    """One lbm_ctx.  Population arrays are numpy Float64 of shape (NX, NY_local, Q) in Fortran
    order -- the memory order of Julia's `f[x, y, i]`."""

    def __init__(self, nx, ny, lattice, collision, tau, bcs=(), dtype=F64, arith=ARITH_EXACT, device=0,
                 rank=0, world=1, nccl_id=None):
        d = make_desc(nx, ny, lattice, collision, tau, bcs, dtype, arith, device, rank, world, nccl_id)
        self._h = C.c_void_p()
        self.desc = d
        check(lib().lbm_create(C.byref(d), C.byref(self._h)))
        y0, nyl = C.c_int32(), C.c_int32()
        check(lib().lbm_local_rows(self._h, C.byref(y0), C.byref(nyl)))
        self.nx, self.ny, self.y0, self.ny_local = d.nx, d.ny, y0.value, nyl.value
        self.Q = lattice_info(d.lattice)["Q"]
        self.shape = (self.nx, self.ny_local, self.Q)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().lbm_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- populations ---------------------------------------------------------------------
    def _check_f(self, f, writable=False):
        if not isinstance(f, np.ndarray) or f.dtype != np.float64 or f.shape != self.shape or not f.flags.f_contiguous:
            raise ValueError(f"populations must be a Fortran-ordered float64 array of shape {self.shape}")
        if writable and not f.flags.writeable:
            raise ValueError("read-only array")
        return f

    def new_f(self):
        return np.empty(self.shape, dtype=np.float64, order="F")

    def upload_f(self, f):
        f = np.asfortranarray(f, dtype=np.float64)
        check(lib().lbm_upload_f(self._h, self._check_f(f).ctypes.data))

    def upload_f_collision(self, f):
        f = np.asfortranarray(f, dtype=np.float64)
        check(lib().lbm_upload_f_collision(self._h, self._check_f(f).ctypes.data))

    def download_f(self, out=None):
        out = self.new_f() if out is None else self._check_f(out, True)
        check(lib().lbm_download_f(self._h, out.ctypes.data))
        return out

    def upload_f_async(self, f):
        """upload_f for a page-locked array (pinned_empty) without waiting: returns once the copies are enqueued"""
        check(lib().lbm_upload_f_async(self._h, self._check_f(f).ctypes.data))

    def download_f_async(self, out):
        """download_f into a page-locked array without waiting; valid after sync()"""
        check(lib().lbm_download_f_async(self._h, self._check_f(out, True).ctypes.data))
        return out

    def snapshot_begin(self, out=None):
        """TakeSnapshots: start an asynchronous copy of f_stream of the current state (take_snapshots.jl:12-29) and return
        the array it will land in -- page-locked unless `out` is given.  Valid after snapshot_end(); steps enqueued in
        between run concurrently with the copy and the fused state machine is not disturbed."""
        out = pinned_empty(self.shape) if out is None else self._check_f(out, True)
        check(lib().lbm_snapshot_begin(self._h, out.ctypes.data))
        self._snapshot_out = out
        return out

    def snapshot_end(self):
        if getattr(self, "_h", None):  # lbm_destroy completes a pending snapshot: nothing left to wait for afterwards
            check(lib().lbm_snapshot_end(self._h))
        self._snapshot_out = None

    def upload_f_rows(self, y0, f_rows):
        """f_rows: (NX, ny, Q) Fortran-ordered block for local rows y0 .. y0+ny-1."""
        f_rows = np.asfortranarray(f_rows, dtype=np.float64)
        if f_rows.ndim != 3 or f_rows.shape[0] != self.nx or f_rows.shape[2] != self.Q:
            raise ValueError(f"expected (NX={self.nx}, ny, Q={self.Q})")
        check(lib().lbm_upload_f_rows(self._h, int(y0), f_rows.shape[1], f_rows.ctypes.data))

    def init_equilibrium_rows(self, y0, rho, ux, uy, T):
        """rows y0.. of f_stream := hermite_based_equilibrium!(q, rho, u, T); fields are (NX, ny) arrays."""
        ny = np.shape(rho)[1]
        arrs = [np.asfortranarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (self.nx, ny))) for a in (rho, ux, uy, T)]
        check(lib().lbm_init_equilibrium_rows(self._h, int(y0), int(ny), *[a.ctypes.data for a in arrs]))

    def init_analytic(self, fields, unit_density=False, unit_temperature=False, offeq=0, offeq_coef=0.0):
        """f_stream := equilibrium (+ off-equilibrium part) of the analytic fields, evaluated on the device.  fields: 8
        separable fields (c0, [(a, X or None, Y or None), ...]) for rho, ux, uy, p, du_x/dx, du_x/dy, du_y/dx, du_y/dy
        (lattice units; X: NX entries, Y: NY_local entries)."""
        arr, keep = _sep_fields(fields, self.nx, self.ny_local)
        spec = lbm_init_spec()
        spec.rho, spec.ux, spec.uy, spec.p = arr[0], arr[1], arr[2], arr[3]
        for k in range(4):
            spec.grad[k] = arr[4 + k]
        spec.unit_density, spec.unit_temperature = int(bool(unit_density)), int(bool(unit_temperature))
        spec.offeq, spec.offeq_coef = int(offeq), float(offeq_coef)
        check(lib().lbm_init_analytic(self._h, C.byref(spec)))
        del keep

    def download_f_rows(self, y0, ny):
        out = np.empty((self.nx, int(ny), self.Q), dtype=np.float64, order="F")
        check(lib().lbm_download_f_rows(self._h, int(y0), int(ny), out.ctypes.data))
        return out

    def download_f_collision(self, out=None):
        out = self.new_f() if out is None else self._check_f(out, True)
        check(lib().lbm_download_f_collision(self._h, out.ctypes.data))
        return out

    # ---- force ---------------------------------------------------------------------------
    def set_force_none(self):
        check(lib().lbm_set_force_none(self._h))

    def set_force_uniform(self, fx, fy):
        check(lib().lbm_set_force_uniform(self._h, float(fx), float(fy)))

    def set_force_field(self, Fx, Fy):
        """Fx, Fy: arrays (NX, NY_local)."""
        F = np.empty((self.nx, self.ny_local, 2), dtype=np.float64, order="F")
        F[:, :, 0] = _as_f64(Fx, (self.nx, self.ny_local))
        F[:, :, 1] = _as_f64(Fy, (self.nx, self.ny_local))
        check(lib().lbm_set_force_field(self._h, F.ctypes.data))

    def set_velocity_field(self, ux, uy):
        """ITERATIVE_INIT contexts: the lattice velocity (NX, NY_local) every node is held at."""
        U = np.empty((self.nx, self.ny_local, 2), dtype=np.float64, order="F")
        U[:, :, 0] = _as_f64(ux, (self.nx, self.ny_local))
        U[:, :, 1] = _as_f64(uy, (self.nx, self.ny_local))
        check(lib().lbm_set_velocity_field(self._h, U.ctypes.data))

    def set_force_separable(self, t0, fx_of_y, fy_of_x):
        """fx_of_y: (nsteps, NY_local), fy_of_x: (nsteps, NX), C order."""
        fx = np.ascontiguousarray(fx_of_y, dtype=np.float64)
        fy = np.ascontiguousarray(fy_of_x, dtype=np.float64)
        n = fx.shape[0]
        if fx.shape != (n, self.ny_local) or fy.shape != (n, self.nx):
            raise ValueError("separable force tables must be (nsteps, NY_local) and (nsteps, NX)")
        check(lib().lbm_set_force_separable(self._h, int(t0), n, fx.ctypes.data, fy.ctypes.data))

    # ---- operators -----------------------------------------------------------------------
    def collide(self, step=0, time=0.0):
        check(lib().lbm_collide(self._h, int(step), float(time)))

    def stream(self):
        check(lib().lbm_stream(self._h))

    def apply_bcs(self, time=0.0):
        check(lib().lbm_apply_bcs(self._h, float(time)))

    def step(self, t0, nsteps, dt=1.0):
        check(lib().lbm_step(self._h, int(t0), int(nsteps), float(dt)))

    def sync(self):
        check(lib().lbm_sync(self._h))

    # ---- diagnostics ---------------------------------------------------------------------
    FIELDS = ("rho", "ux", "uy", "p", "p_track", "sxx", "sxy", "syy")

    def moments(self, tau_visc=1.0, fields=("rho", "ux", "uy")):
        """dict of (NX, NY_local) Fortran arrays for the requested fields (lbm_moments)."""
        out = {k: np.empty((self.nx, self.ny_local), dtype=np.float64, order="F") for k in fields}
        ptrs = [out[k].ctypes.data if k in out else None for k in self.FIELDS]
        check(lib().lbm_moments(self._h, float(tau_visc), *ptrs))
        return out

    def reduce(self, kind):
        buf = (C.c_double * 4)()
        check(lib().lbm_reduce(self._h, int(kind), buf, 4))
        return np.array(buf[:])

    def reduce_errors(self, tau_visc, u_max, expected):
        """expected: 8 tuples (c0, [(a, X or None, Y or None), ...up to 2 terms]) for
        rho, ux, uy, p, sxx, sxy, syx, syy; X has NX entries, Y has NY_local.  Returns the 16 local sums."""
        arr, keep = _sep_fields(expected, self.nx, self.ny_local)
        out = (C.c_double * 16)()
        check(lib().lbm_reduce_errors(self._h, float(tau_visc), float(u_max), arr, out))
        del keep
        return np.array(out[:])

    def reduce_process(self, u_max, expected):
        """The 12 local sums of process! (CompareWithAnalyticalSolution; see lbm_reduce_process).  expected: the first
        four separable fields of reduce_errors (rho, ux, uy, p); further entries are ignored."""
        expected = list(expected[:4]) + [(0.0, [])] * 4
        arr, keep = _sep_fields(expected, self.nx, self.ny_local)
        out = (C.c_double * 16)()
        check(lib().lbm_reduce_process(self._h, float(u_max), arr, out))
        del keep
        return np.array(out[:12])

    # ---- introspection -------------------------------------------------------------------
    @property
    def kernel_launches(self):
        return int(lib().lbm_kernel_launches(self._h))

    @property
    def halo_path(self):
        """0 single GPU, 1 NCCL send/recv, 2 peer-memory stores from the boundary-row launch."""
        return int(lib().lbm_halo_path(self._h))

    def last_step_ms(self):
        ms = C.c_float()
        check(lib().lbm_last_step_ms(self._h, C.byref(ms)))
        return ms.value

    def timer_start(self):
        check(lib().lbm_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().lbm_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_option(self, key, value):
        check(lib().lbm_set_option(self._h, key.encode(), int(value)))


def pinned_empty(shape):
    """np.empty(shape, float64, order="F") in page-locked host memory (lbm_host_alloc); freed with the array."""
    import weakref
    n = int(np.prod(shape))
    p = C.c_void_p()
    check(lib().lbm_host_alloc(C.byref(p), n * 8))
    buf = (C.c_double * max(n, 1)).from_address(p.value)
    weakref.finalize(buf, lib().lbm_host_free, p.value)
    return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape, order="F")


def _sep_fields(expected, nx, ny):
    """expected: 8 tuples (c0, [(a, X or None, Y or None), ...up to 2 terms]) -> (lbm_sep_field * 8, keep-alive list)"""
    arr = (lbm_sep_field * 8)()
    keep = []
    for f, (c0, terms) in enumerate(expected):
        arr[f].c0 = float(c0)
        if len(terms) > 2:
            raise ValueError("at most two separable terms per field")
        for k, (a, X, Y) in enumerate(terms):
            arr[f].a[k] = float(a)
            for name, tab, n in (("x", X, nx), ("y", Y, ny)):
                if tab is not None:
                    t = np.ascontiguousarray(tab, dtype=np.float64)
                    if t.shape != (n,):
                        raise ValueError(f"separable table '{name}' must have {n} entries")
                    keep.append(t)
                    getattr(arr[f], name)[k] = t.ctypes.data
    return arr, keep


class Batch:
    """One lbm_batch: `nbatch` independent problems of one shape, advanced by a single launch that keeps every problem
    on chip.  Population arrays are numpy Float64 of shape (NX, NY, Q, nbatch) in Fortran order (problem after problem,
    each in the memory order of Julia's `f[x, y, i]`)."""

    def __init__(self, nbatch, nx, ny, lattice, collision, tau, bcs=(), dtype=F64, arith=ARITH_EXACT, device=0):
        d = make_desc(nx, ny, lattice, collision, tau, bcs, dtype, arith, device)
        self._h = C.c_void_p()
        self.desc = d
        check(lib().lbm_batch_create(C.byref(d), int(nbatch), C.byref(self._h)))
        self.nbatch, self.nx, self.ny, self.ntau = int(nbatch), d.nx, d.ny, d.ntau
        self.Q = lattice_info(d.lattice)["Q"]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().lbm_batch_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_tau(self, tau):
        """tau: (nbatch, ntau)"""
        t = np.ascontiguousarray(tau, dtype=np.float64)
        if t.shape != (self.nbatch, self.ntau):
            raise ValueError(f"relaxation times must have shape ({self.nbatch}, {self.ntau})")
        check(lib().lbm_batch_set_tau(self._h, t.ctypes.data))

    def set_force_uniform(self, fxy):
        """fxy: (nbatch, 2) lattice force per problem, or None"""
        if fxy is None:
            check(lib().lbm_batch_set_force_uniform(self._h, None))
            return
        f = np.ascontiguousarray(fxy, dtype=np.float64)
        if f.shape != (self.nbatch, 2):
            raise ValueError(f"forces must have shape ({self.nbatch}, 2)")
        check(lib().lbm_batch_set_force_uniform(self._h, f.ctypes.data))

    def upload_f(self, f, first=0):
        """f: (NX, NY, Q, count) Fortran-ordered"""
        f = np.asfortranarray(f, dtype=np.float64)
        if f.ndim != 4 or f.shape[:3] != (self.nx, self.ny, self.Q):
            raise ValueError(f"expected (NX={self.nx}, NY={self.ny}, Q={self.Q}, count)")
        check(lib().lbm_batch_upload_f(self._h, int(first), f.shape[3], f.ctypes.data))

    def broadcast_f(self, f):
        """f: (NX, NY, Q): the same initial f_stream for every problem"""
        f = np.asfortranarray(f, dtype=np.float64)
        if f.shape != (self.nx, self.ny, self.Q):
            raise ValueError(f"expected (NX={self.nx}, NY={self.ny}, Q={self.Q})")
        check(lib().lbm_batch_broadcast_f(self._h, f.ctypes.data))

    def download_f(self, first=0, count=None):
        count = self.nbatch - first if count is None else int(count)
        out = np.empty((self.nx, self.ny, self.Q, count), dtype=np.float64, order="F")
        check(lib().lbm_batch_download_f(self._h, int(first), count, out.ctypes.data))
        return out

    def run(self, nsteps, stop_kind=BATCH_STOP_OFF, check_every=100, tolerance=0.0):
        st = lbm_batch_stop(int(stop_kind), int(check_every), float(tolerance))
        check(lib().lbm_batch_run(self._h, int(nsteps), C.byref(st) if stop_kind else None))

    def status(self, first=0, count=None):
        """-> (steps_done int64[count], stopped bool[count]); synchronises"""
        count = self.nbatch - first if count is None else int(count)
        steps = np.empty(count, dtype=np.int64)
        stopped = np.empty(count, dtype=np.int32)
        check(lib().lbm_batch_status(self._h, int(first), count, steps.ctypes.data, stopped.ctypes.data))
        return steps, stopped.astype(bool)

    def reduce_errors(self, tau_visc, u_max, expected, coef=None):
        """-> (nbatch, 16) sums of TrackHydrodynamicErrors.next! per problem.  expected as Context.reduce_errors (tables
        shared by all problems); coef (nbatch, 8, 3) = per-problem (c0, a0, a1) of every field, or None."""
        tv = np.ascontiguousarray(np.broadcast_to(np.asarray(tau_visc, dtype=np.float64), (self.nbatch,)))
        um = np.ascontiguousarray(np.broadcast_to(np.asarray(u_max, dtype=np.float64), (self.nbatch,)))
        arr, keep = _sep_fields(expected, self.nx, self.ny)
        cf = None
        if coef is not None:
            cf = np.ascontiguousarray(coef, dtype=np.float64)
            if cf.shape != (self.nbatch, 8, 3):
                raise ValueError(f"coef must have shape ({self.nbatch}, 8, 3)")
        out = np.empty((self.nbatch, 16), dtype=np.float64)
        check(lib().lbm_batch_reduce_errors(self._h, tv.ctypes.data, um.ctypes.data, arr, cf.ctypes.data if cf is not None else None,
                                            out.ctypes.data))
        del keep
        return out

    def last_run_ms(self):
        ms = C.c_float()
        check(lib().lbm_batch_last_run_ms(self._h, C.byref(ms)))
        return ms.value

    @property
    def kernel_launches(self):
        return int(lib().lbm_batch_kernel_launches(self._h))
