"""ORACLE -- ctypes loader for the C restatement (oracle/lbm_oracle.c).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblbm_oracle.so")
_lib = None

QMAX = 37


class CollisionT(C.Structure):
    _fields_ = [("model", C.c_int), ("tau", C.c_double * 16), ("force_mode", C.c_int),
                ("fx", C.c_double), ("fy", C.c_double), ("field", C.c_void_p)]


class BcT(C.Structure):
    _fields_ = [("kind", C.c_int), ("dir", C.c_int), ("x0", C.c_int), ("x1", C.c_int),
                ("y0", C.c_int), ("y1", C.c_int), ("ux", C.c_double), ("uy", C.c_double), ("rho", C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "lbm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_lattice_size.restype = C.c_size_t
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


_DIR = {"N": 0, "E": 1, "S": 2, "W": 3}


class COracle:
    """Steps f[Q,NY,NX] (float64) with the C restatement."""

    def __init__(self, q, cm, bcs=(), threads=None):
        from . import lbm_oracle as O
        l = lib()
        if threads:
            l.oracle_set_threads(int(threads))
        self.q = q
        self._lat = C.create_string_buffer(l.oracle_lattice_size())
        cx = np.ascontiguousarray(q.cx, dtype=np.int32)
        cy = np.ascontiguousarray(q.cy, dtype=np.int32)
        opp = np.ascontiguousarray(q.opp, dtype=np.int32)
        w = np.ascontiguousarray(q.w, dtype=np.float64)
        l.oracle_lattice_init(self._lat, C.c_int(q.Q), cx.ctypes, cy.ctypes, w.ctypes, C.c_double(q.css),
                              opp.ctypes, C.c_int(q.eq_order), C.c_int(q.N))
        c = CollisionT()
        if isinstance(cm, O.SRT):
            c.model = 0
            c.tau[0] = cm.tau
        elif isinstance(cm, O.TRT):
            c.model = 1
            c.tau[0], c.tau[1] = cm.tau_s, cm.tau_a
        else:
            c.model = 2
            for i, t in enumerate(cm.taus):
                c.tau[i] = t
        self._field = None
        F = cm.force
        if F is None:
            c.force_mode = 0
        elif callable(F):
            raise ValueError("time-dependent force: drive the numpy oracle instead")
        elif np.isscalar(F[0]):
            c.force_mode, c.fx, c.fy = 1, float(F[0]), float(F[1])
        else:
            self._field = np.ascontiguousarray(np.stack([F[0], F[1]]), dtype=np.float64)
            c.force_mode = 2
            c.field = self._field.ctypes.data
        self._cm = c
        arr = (BcT * max(1, len(bcs)))()
        for k, bc in enumerate(bcs):
            arr[k].kind = 1 if isinstance(bc, O.MovingWall) else 0
            arr[k].dir = _DIR[bc.direction]
            arr[k].x0, arr[k].x1 = bc.xs
            arr[k].y0, arr[k].y1 = bc.ys
            if isinstance(bc, O.MovingWall):
                if bc.direction != "N":
                    raise NotImplementedError
                arr[k].ux, arr[k].uy, arr[k].rho = bc.u[0], bc.u[1], bc.rho
        self._bcs, self._nbc = arr, len(bcs)

    def collide(self, f):
        out = np.empty_like(f)
        Q, ny, nx = f.shape
        lib().oracle_collide(self._lat, C.byref(self._cm), nx, ny, f.ctypes, out.ctypes)
        return out

    def stream(self, f):
        out = np.empty_like(f)
        Q, ny, nx = f.shape
        lib().oracle_stream(self._lat, nx, ny, f.ctypes, out.ctypes)
        return out

    def apply_bcs(self, f_new, f_old):
        Q, ny, nx = f_new.shape
        lib().oracle_apply_bcs(self._lat, self._nbc, self._bcs, nx, ny, f_new.ctypes, f_old.ctypes)
        return f_new

    def steps_inplace(self, fs, fc, nsteps):
        """Advance fs (f_stream, C-contiguous float64 [Q,NY,NX]) by nsteps in place; fc receives f_collision."""
        Q, ny, nx = fs.shape
        assert fs.flags.c_contiguous and fc.flags.c_contiguous and fs.dtype == np.float64 == fc.dtype and fc.shape == fs.shape
        lib().oracle_steps(self._lat, C.byref(self._cm), self._nbc, self._bcs, nx, ny, fs.ctypes, fc.ctypes, int(nsteps))

    def steps(self, f_stream, nsteps):
        """Returns (f_stream, f_collision) after nsteps; inputs untouched."""
        fs = np.ascontiguousarray(f_stream, dtype=np.float64).copy()
        fc = np.empty_like(fs)
        Q, ny, nx = fs.shape
        lib().oracle_steps(self._lat, C.byref(self._cm), self._nbc, self._bcs, nx, ny, fs.ctypes, fc.ctypes,
                           int(nsteps))
        return fs, fc


def num_threads():
    return lib().oracle_num_threads()
