"""ctypes binding of oracle/libpbf_oracle.so (the CPU checker). Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libpbf_oracle.so")


class Params(C.Structure):
    """Same layout as pbf_params / orc_params / ref_params (reference GUIParams.h:7-17)."""
    _fields_ = [("niter", C.c_int32), ("pho0", C.c_float), ("g", C.c_float), ("h", C.c_float),
                ("dt", C.c_float), ("lambda_eps", C.c_float), ("delta_q", C.c_float),
                ("k_corr", C.c_float), ("n_corr", C.c_float), ("k_boundaryDensity", C.c_float),
                ("c_XSPH", C.c_float)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


def build():
    src = os.path.join(ORACLE_DIR, "pbf_oracle.c")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        f3 = C.POINTER(C.c_float)
        u32p = C.POINTER(C.c_uint32)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(Params), f3, f3, C.c_int64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.orc_set_lim.argtypes = [C.c_void_p, f3, f3]
        L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_step.argtypes = [C.c_void_p, f3, f3, f3, f3, u32p, C.c_int64]
        L.orc_bind.argtypes = [C.c_void_p, f3, f3, f3, f3, u32p, C.c_int64]
        for name in ("orc_advect", "orc_build_grid", "orc_correct_density", "orc_update_velocity",
                     "orc_correct_velocity"):
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("orc_grid_id", "orc_grid_start", "orc_grid_end"):
            getattr(L, name).restype = u32p
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("orc_lambda", "orc_pho", "orc_tpos"):
            getattr(L, name).restype = f3
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_grid_dim.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        for name in ("orc_coef_corr", "orc_poly6_coef", "orc_spiky_coef"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_poly6.restype = C.c_float
        L.orc_poly6.argtypes = [C.c_void_p, C.c_float]
        L.orc_neighbor_count.argtypes = [C.c_void_p, u32p]
        L.orc_candidate_count.argtypes = [C.c_void_p, u32p]
        L.orc_lambda_allpairs.argtypes = [C.c_void_p, f3, f3, u32p]
        L.orc_scene_cube.restype = C.c_int64
        L.orc_scene_cube.argtypes = [f3, f3, C.POINTER(C.c_int32), u32p, C.c_uint32, f3, f3, u32p]
        L.orc_scene_double_dam_reference.restype = C.c_int64
        L.orc_scene_double_dam_reference.argtypes = [f3, f3, u32p, f3, f3]
        L.orc_scene_block.argtypes = [f3, C.POINTER(C.c_int32), C.c_float, C.c_uint32, C.c_uint32, f3, f3, u32p]
        L.orc_wall_lim.argtypes = [f3, f3, f3, f3, C.c_float, C.c_int, C.c_int, f3, f3]
        L.orc_stats.argtypes = [f3, f3, f3, C.c_int64, C.c_float, C.POINTER(C.c_double)]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def fptr(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def uptr(a):
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def iptr(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def default_params():
    p = Params()
    lib().orc_default_params(C.byref(p))
    return p


def scene_double_dam_reference():
    """The reference's shipped 32 000-particle scene (FluidSystem.cpp:55-61)."""
    n = 32000
    pos = np.zeros((n, 3), np.float32)
    vel = np.zeros((n, 3), np.float32)
    iid = np.zeros(n, np.uint32)
    ulim = np.zeros(3, np.float32)
    llim = np.zeros(3, np.float32)
    cnt = lib().orc_scene_double_dam_reference(fptr(pos), fptr(vel), uptr(iid), fptr(ulim), fptr(llim))
    assert cnt == n
    return pos, vel, iid, ulim, llim


def scene_cube(ulim, llim, ns, seed=27, first_iid=0):
    ulim = np.asarray(ulim, np.float32)
    llim = np.asarray(llim, np.float32)
    ns = np.asarray(ns, np.int32)
    n = int(ns.prod())
    pos = np.zeros((n, 3), np.float32)
    vel = np.zeros((n, 3), np.float32)
    iid = np.zeros(n, np.uint32)
    st = C.c_uint32(seed)
    cnt = lib().orc_scene_cube(fptr(ulim), fptr(llim), iptr(ns), C.byref(st), first_iid, fptr(pos), fptr(vel), uptr(iid))
    assert cnt == n
    return pos, vel, iid


def scene_block(origin, n3, spacing=0.05, seed=27, first_iid=0):
    origin = np.asarray(origin, np.float32)
    n3 = np.asarray(n3, np.int32)
    n = int(n3.astype(np.int64).prod())
    pos = np.zeros((n, 3), np.float32)
    vel = np.zeros((n, 3), np.float32)
    iid = np.zeros(n, np.uint32)
    lib().orc_scene_block(fptr(origin), iptr(n3), spacing, seed, first_iid, fptr(pos), fptr(vel), uptr(iid))
    return pos, vel, iid


class Oracle:
    """Host mirror of the reference's Simulator on numpy arrays."""

    def __init__(self, params, ulim, llim, max_particles, threads=1):
        self.params = params
        self.ulim = np.asarray(ulim, np.float32).copy()
        self.llim = np.asarray(llim, np.float32).copy()
        self.h = lib().orc_create(C.byref(params), fptr(self.ulim), fptr(self.llim), int(max_particles))
        lib().orc_set_threads(self.h, threads)
        self.n = 0

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_lim(self, ulim, llim):
        self.ulim = np.asarray(ulim, np.float32).copy()
        self.llim = np.asarray(llim, np.float32).copy()
        lib().orc_set_lim(self.h, fptr(self.ulim), fptr(self.llim))

    def set_params(self, params):
        self.params = params
        lib().orc_set_params(self.h, C.byref(params))

    def step(self, pos, npos, vel, nvel, iid):
        self.n = len(iid)
        self._keep = (pos, npos, vel, nvel, iid)
        lib().orc_step(self.h, fptr(pos), fptr(npos), fptr(vel), fptr(nvel), uptr(iid), self.n)

    def bind(self, pos, npos, vel, nvel, iid):
        self.n = len(iid)
        self._keep = (pos, npos, vel, nvel, iid)
        lib().orc_bind(self.h, fptr(pos), fptr(npos), fptr(vel), fptr(nvel), uptr(iid), self.n)

    def advect(self): lib().orc_advect(self.h)
    def build_grid(self): lib().orc_build_grid(self.h)
    def correct_density(self): lib().orc_correct_density(self.h)
    def update_velocity(self): lib().orc_update_velocity(self.h)
    def correct_velocity(self): lib().orc_correct_velocity(self.h)

    def grid_dim(self):
        d = (C.c_int32 * 3)()
        lib().orc_grid_dim(self.h, d)
        return tuple(d)

    def _u32(self, fn, n):
        return np.ctypeslib.as_array(fn(self.h), shape=(n,)).copy()

    def _f32(self, fn, n):
        return np.ctypeslib.as_array(fn(self.h), shape=(n,)).copy()

    def grid_id(self): return self._u32(lib().orc_grid_id, self.n)

    def grid_start(self):
        d = self.grid_dim()
        return self._u32(lib().orc_grid_start, d[0] * d[1] * d[2])

    def grid_end(self):
        d = self.grid_dim()
        return self._u32(lib().orc_grid_end, d[0] * d[1] * d[2])

    def lam(self): return self._f32(lib().orc_lambda, self.n)
    def pho(self): return self._f32(lib().orc_pho, self.n)
    def tpos(self): return self._f32(lib().orc_tpos, 3 * self.n).reshape(-1, 3)
    def coef_corr(self): return float(lib().orc_coef_corr(self.h))

    def neighbor_count(self):
        out = np.zeros(self.n, np.uint32)
        lib().orc_neighbor_count(self.h, uptr(out))
        return out

    def candidate_count(self):
        out = np.zeros(self.n, np.uint32)
        lib().orc_candidate_count(self.h, uptr(out))
        return out

    def lambda_allpairs(self):
        lam = np.zeros(self.n, np.float32)
        pho = np.zeros(self.n, np.float32)
        cnt = np.zeros(self.n, np.uint32)
        lib().orc_lambda_allpairs(self.h, fptr(lam), fptr(pho), uptr(cnt))
        return lam, pho, cnt


def stats(pho, npos, nvel, pho0):
    out = (C.c_double * 5)()
    lib().orc_stats(fptr(pho), fptr(npos), fptr(nvel), len(pho), pho0, out)
    return dict(density_err_mean=out[0], density_err_max=out[1], kinetic_energy=out[2],
                max_speed=out[3], mean_z=out[4])


def wall_lim(ulim0, llim0, a_ulim, a_llim, w, frame, start_frame=0):
    ulim0 = np.asarray(ulim0, np.float32); llim0 = np.asarray(llim0, np.float32)
    a_ulim = np.asarray(a_ulim, np.float32); a_llim = np.asarray(a_llim, np.float32)
    u = np.zeros(3, np.float32); l = np.zeros(3, np.float32)
    lib().orc_wall_lim(fptr(ulim0), fptr(llim0), fptr(a_ulim), fptr(a_llim), w, frame, start_frame, fptr(u), fptr(l))
    return u, l
