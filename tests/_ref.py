"""ctypes binding of oracle/_ref/libpbf_ref.so — the reference's OWN Simulator.cu compiled headless
(oracle/ref_glue.cu, oracle/Makefile target `ref`). Test infrastructure only; needs a GPU to run.
Used by tests/golden/make_golden.py (to pin the oracle), the GPU parity tests and
`bench.py --impl reference`."""
import ctypes as C
import os

import numpy as np

from _oracle import Params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libpbf_ref.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, f3 = C.c_void_p, C.POINTER(C.c_float)
        L.ref_create.restype = vp
        L.ref_create.argtypes = [C.POINTER(Params), f3, f3, C.c_int64]
        L.ref_destroy.argtypes = [vp]
        L.ref_set_params.argtypes = [vp, C.POINTER(Params)]
        L.ref_set_lim.argtypes = [vp, f3, f3]
        L.ref_bind.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64]
        L.ref_stage.argtypes = [vp, C.c_int]
        L.ref_step.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64]
        L.ref_read.argtypes = [vp, C.c_int, vp, C.c_int64]
        L.ref_grid_dim.argtypes = [vp, C.POINTER(C.c_int32)]
        L.ref_coef_corr.restype = C.c_float
        L.ref_coef_corr.argtypes = [vp]
        _lib = L
    return _lib


def _f3(v):
    a = np.ascontiguousarray(np.asarray(v, np.float32).reshape(3))
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _ptr(x):
    return x if isinstance(x, int) else x.data_ptr()


class RefSimulator:
    """The reference's Simulator driven headless on raw device pointers (torch CUDA tensors)."""
    ADVECT, GRID, DENSITY, VELOCITY_UPDATE, VELOCITY_CORRECT = range(5)

    def __init__(self, params, ulim, llim, max_particles):
        p = Params()
        C.memmove(C.byref(p), C.byref(params), C.sizeof(Params))
        u, up = _f3(ulim)
        l, lp = _f3(llim)
        self.h = lib().ref_create(C.byref(p), up, lp, int(max_particles))
        if not self.h:
            raise RuntimeError("ref_create failed")
        self.n = 0

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def set_lim(self, ulim, llim):
        u, up = _f3(ulim)
        l, lp = _f3(llim)
        lib().ref_set_lim(self.h, up, lp)

    def bind(self, pos, npos, vel, nvel, iid, n):
        self.n = int(n)
        lib().ref_bind(self.h, _ptr(pos), _ptr(npos), _ptr(vel), _ptr(nvel), _ptr(iid), int(n))

    def stage(self, k):
        rc = lib().ref_stage(self.h, k)
        if rc:
            raise RuntimeError("ref_stage %d failed: %d" % (k, rc))

    def step(self, pos, npos, vel, nvel, iid, n):
        self.n = int(n)
        rc = lib().ref_step(self.h, _ptr(pos), _ptr(npos), _ptr(vel), _ptr(nvel), _ptr(iid), int(n))
        if rc:
            raise RuntimeError("ref_step failed: %d" % rc)

    def grid_dim(self):
        d = (C.c_int32 * 3)()
        lib().ref_grid_dim(self.h, d)
        return tuple(d)

    def _read(self, what, count, dtype, shape):
        out = np.empty(shape, dtype)
        rc = lib().ref_read(self.h, what, out.ctypes.data, count)
        if rc:
            raise RuntimeError("ref_read failed")
        return out

    def grid_id(self): return self._read(0, self.n, np.uint32, self.n)

    def grid_start(self):
        d = self.grid_dim(); c = d[0] * d[1] * d[2]
        return self._read(1, c, np.uint32, c)

    def grid_end(self):
        d = self.grid_dim(); c = d[0] * d[1] * d[2]
        return self._read(2, c, np.uint32, c)

    def lam(self): return self._read(3, self.n, np.float32, self.n)
    def pho(self): return self._read(4, self.n, np.float32, self.n)
    def tpos(self): return self._read(5, self.n, np.float32, (self.n, 3))
    def coef_corr(self): return float(lib().ref_coef_corr(self.h))
