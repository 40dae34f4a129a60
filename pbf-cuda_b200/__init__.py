"""pbf-cuda_b200 — Python host-side mirror of the reference's Simulator interface over the C-ABI.

The product is `libpbf_b200.so` (hand-written sm_100a kernels behind include/pbf.h). This module
is a thin ctypes binding used by tests/, bench.py and __graft_entry__.py; it mirrors the
reference's `Simulator` (fluids/Simulator.h:10,37-42), `GUIParams` (fluids/GUIParams.h:7-17) and
`ParticleSource` scenes (fluids/ParticleSource.h:11-13). There is no CPU fallback: if the shared
library is missing, importing this module fails; if no sm_100 GPU is present, every compute call
raises PbfError.

The directory name has a hyphen (the project's name), so import it with
    pbf = importlib.import_module("pbf-cuda_b200")
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBF_LIB selects a tuning-experiment build (pbf-cuda_b200/Makefile VARIANT=...); default is the product
LIB_PATH = os.environ.get("PBF_LIB") or os.path.join(_HERE, "libpbf_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_STATE = 0, 1, 2, 3, 4

READ_KEY, READ_SRC_INDEX, READ_IID, READ_CELL_START, READ_CELL_END, READ_NPOS, READ_LAMBDA, READ_RHO, \
    READ_POS0, READ_VEL, READ_NEIGHBOR_COUNT = range(11)

KERNEL_NAMES = ("advect_key", "sort", "reorder", "lambda", "delta_p", "update_velocity", "xsph")
STAGE_NAMES = ("ADVECT", "GRID", "DENSITY", "VELOCITY_UPDATE", "VELOCITY_CORRECT")  # reference Logger.h:7-23

# every symbol include/pbf.h declares (tests check the library exports all of them)
EXPORTS = (
    "pbf_default_params", "pbf_create", "pbf_destroy", "pbf_set_params", "pbf_get_params", "pbf_set_lim",
    "pbf_get_lim", "pbf_set_option_exact_pow", "pbf_get_grid_dim", "pbf_step", "pbf_step_host",
    "pbf_stage_begin", "pbf_stage_advect", "pbf_stage_build_grid", "pbf_stage_correct_density",
    "pbf_stage_update_velocity", "pbf_stage_correct_velocity", "pbf_stage_end", "pbf_read", "pbf_get_stats",
    "pbf_enable_stage_timing", "pbf_get_stage_ms", "pbf_get_kernel_ms", "pbf_launch_count", "pbf_device_alloc", "pbf_device_free",
    "pbf_copy_h2d", "pbf_copy_d2h", "pbf_device_sync", "pbf_scene_cube", "pbf_scene_double_dam_reference",
    "pbf_scene_block_device", "pbf_scene_block_host", "pbf_last_error", "pbf_version",
    "pbf_slab_begin", "pbf_slab_get_layout", "pbf_slab_plane_counts", "pbf_stage_lambda", "pbf_stage_delta_p",
    "pbf_slab_halo", "pbf_slab_flags", "pbf_slab_sort_state", "pbf_scene_block_slice_device",
    "pbf_scene_block_slice_host", "pbf_slab_peer_export", "pbf_slab_peer_attach",
    "pbf_slab_halo_sync", "pbf_slab_push_state", "pbf_slab_register_state", "pbf_slab_adopt_state", "pbf_stream_create", "pbf_stream_destroy",
    "pbf_stream_sync", "pbf_copy_d2h_async", "pbf_device_count", "pbf_get_const_div_interval",
    "pbf_get_fast_spiky", "pbf_get_pair_list", "pbf_get_trim_pow", "pbf_set_option", "pbf_get_option", "pbf_state_digest_device",
    "pbf_state_digest_host", "pbf_state_write", "pbf_state_read_info", "pbf_state_read", "pbf_checkpoint_save", "pbf_checkpoint_load",
)

HALO_LAMBDA, HALO_POSITION, HALO_VELOCITY = 0, 1, 2
SLAB_STATE_PUSHED = -1   # pbf_slab_step.pull_left_first: the neighbours stored the raw state themselves (pbf_slab_push_state)
OPT_TEAM, OPT_REBIN, OPT_PDL, OPT_GRAPH, OPT_HALO_INKERNEL, OPT_STAGED, OPT_PAIRED, OPT_MORTON, OPT_COOP = 0, 1, 2, 3, 4, 5, 6, 7, 8
SLAB_FLAG_MIGRATION, SLAB_FLAG_GHOST, SLAB_FLAG_TIMEOUT = 1, 2, 4


class PbfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pbf error %d: %s" % (code, msg))
        self.code = code


class GUIParams(C.Structure):
    """Fluid half of the reference's GUIParams (fluids/GUIParams.h:7-17); layout == pbf_params."""
    _fields_ = [("niter", C.c_int32), ("pho0", C.c_float), ("g", C.c_float), ("h", C.c_float),
                ("dt", C.c_float), ("lambda_eps", C.c_float), ("delta_q", C.c_float),
                ("k_corr", C.c_float), ("n_corr", C.c_float), ("k_boundaryDensity", C.c_float),
                ("c_XSPH", C.c_float)]

    def copy(self):
        q = GUIParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(GUIParams))
        return q


class SlabStep(C.Structure):
    """pbf_slab_step (include/pbf.h): one rank's view of one step of the x-slab decomposition."""
    _fields_ = [("x_begin", C.c_int32), ("x_end", C.c_int32), ("ghost", C.c_int32), ("has_left", C.c_int32),
                ("has_right", C.c_int32), ("n_own", C.c_int64), ("m_left", C.c_int64), ("m_right", C.c_int64),
                ("send_left_end", C.c_int64), ("send_right_begin", C.c_int64), ("pull_left_first", C.c_int64)]


class SlabLayout(C.Structure):
    """pbf_slab_layout (include/pbf.h)."""
    _fields_ = [("n_local", C.c_int64), ("own_first", C.c_int64), ("own_count", C.c_int64),
                ("send_left_count", C.c_int64), ("send_right_count", C.c_int64),
                ("recv_left_count", C.c_int64), ("recv_right_count", C.c_int64), ("flags", C.c_uint32)]


class SlabPeerInfo(C.Structure):
    """pbf_slab_peer_info (include/pbf.h): a rank's solver arrays as CUDA IPC handles / raw pointers."""
    _fields_ = [("ipc", (C.c_ubyte * 64) * 9), ("ptr", C.c_uint64 * 9), ("pid", C.c_int64), ("device", C.c_int32),
                ("has_state", C.c_int32), ("state_capacity", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("density_err_mean", C.c_double), ("density_err_max", C.c_double),
                ("kinetic_energy", C.c_double), ("max_speed", C.c_double), ("mean_z", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class StateInfo(C.Structure):
    """pbf_state_info (include/pbf.h): header of a state file (checkpoint)."""
    _fields_ = [("n", C.c_int64), ("frame", C.c_int64), ("params", GUIParams), ("ulim", C.c_float * 3),
                ("llim", C.c_float * 3), ("exact_pow", C.c_int32), ("reserved", C.c_uint32), ("checksum", C.c_uint64)]


if not os.path.exists(LIB_PATH):
    raise ImportError("%s is missing: run `make -C pbf-cuda_b200` (or __graft_entry__.build()); "
                      "there is no CPU fallback" % LIB_PATH)

_lib = C.CDLL(LIB_PATH)
_vp, _i64, _f3 = C.c_void_p, C.c_int64, C.POINTER(C.c_float)
_lib.pbf_default_params.argtypes = [C.POINTER(GUIParams)]
_lib.pbf_create.argtypes = [C.POINTER(GUIParams), _f3, _f3, _i64, C.c_int, C.POINTER(_vp)]
_lib.pbf_destroy.argtypes = [_vp]
_lib.pbf_set_params.argtypes = [_vp, C.POINTER(GUIParams)]
_lib.pbf_get_params.argtypes = [_vp, C.POINTER(GUIParams)]
_lib.pbf_set_lim.argtypes = [_vp, _f3, _f3]
_lib.pbf_get_lim.argtypes = [_vp, _f3, _f3]
_lib.pbf_set_option_exact_pow.argtypes = [_vp, C.c_int]
_lib.pbf_get_grid_dim.argtypes = [_vp, C.POINTER(C.c_int32)]
_lib.pbf_get_fast_spiky.argtypes = [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
_lib.pbf_get_pair_list.argtypes = [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
_lib.pbf_get_trim_pow.argtypes = [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
_lib.pbf_set_option.argtypes = [_vp, C.c_int, C.c_int]
_lib.pbf_get_option.argtypes = [_vp, C.c_int, C.POINTER(C.c_int)]
_lib.pbf_state_digest_device.argtypes = [C.c_int, _vp, _vp, _vp, _i64, _vp, C.POINTER(C.c_uint64)]
_lib.pbf_state_digest_host.argtypes = [_vp, _vp, _vp, _i64, C.POINTER(C.c_uint64)]
_lib.pbf_step.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]
_lib.pbf_step_host.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64]
_lib.pbf_stage_begin.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]
for _n in ("advect", "build_grid", "correct_density", "update_velocity", "correct_velocity", "end"):
    getattr(_lib, "pbf_stage_" + _n).argtypes = [_vp]
_lib.pbf_read.argtypes = [_vp, C.c_int, _vp, _i64]
_lib.pbf_get_stats.argtypes = [_vp, _vp, _vp, _i64, C.POINTER(Stats)]
_lib.pbf_enable_stage_timing.argtypes = [_vp, C.c_int]
_lib.pbf_get_stage_ms.argtypes = [_vp, _f3]
_lib.pbf_get_kernel_ms.argtypes = [_vp, _f3]
_lib.pbf_launch_count.argtypes = [_vp]
_lib.pbf_launch_count.restype = _i64
_lib.pbf_device_alloc.argtypes = [C.c_int, _i64, C.POINTER(_vp)]
_lib.pbf_device_free.argtypes = [C.c_int, _vp]
_lib.pbf_copy_h2d.argtypes = [_vp, _vp, _i64]
_lib.pbf_copy_d2h.argtypes = [_vp, _vp, _i64]
_lib.pbf_device_sync.argtypes = [C.c_int]
_lib.pbf_scene_cube.argtypes = [_f3, _f3, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_uint32, _vp, _vp, _vp,
                                _i64, C.POINTER(_i64)]
_lib.pbf_scene_double_dam_reference.argtypes = [_vp, _vp, _vp, _i64, C.POINTER(_i64), _f3, _f3]
_lib.pbf_scene_block_device.argtypes = [_f3, C.POINTER(C.c_int32), C.c_float, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp]
_lib.pbf_scene_block_host.argtypes = [_f3, C.POINTER(C.c_int32), C.c_float, C.c_uint32, C.c_uint32, _vp, _vp, _vp]
_lib.pbf_slab_begin.argtypes = [_vp, C.POINTER(SlabStep), _vp, _vp, _vp, _vp, _vp, _vp]
_lib.pbf_slab_get_layout.argtypes = [_vp, C.POINTER(SlabLayout)]
_lib.pbf_slab_plane_counts.argtypes = [_vp, C.c_int32, C.c_int32, C.POINTER(_i64)]
_lib.pbf_stage_lambda.argtypes = [_vp]
_lib.pbf_stage_delta_p.argtypes = [_vp]
_lib.pbf_slab_halo.argtypes = [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]
_lib.pbf_slab_flags.argtypes = [_vp, C.POINTER(C.c_uint32)]
_lib.pbf_slab_sort_state.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _i64, _vp]
_lib.pbf_scene_block_slice_device.argtypes = [_f3, C.POINTER(C.c_int32), C.c_float, C.c_uint32, C.c_uint32, C.c_int32,
                                              C.c_int32, _vp, _vp, _vp, _vp]
_lib.pbf_scene_block_slice_host.argtypes = [_f3, C.POINTER(C.c_int32), C.c_float, C.c_uint32, C.c_uint32, C.c_int32,
                                            C.c_int32, _vp, _vp, _vp]
_lib.pbf_slab_adopt_state.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _i64,
                                      C.POINTER(_i64), _vp]
_lib.pbf_get_const_div_interval.argtypes = [_vp, _f3, _f3]
_lib.pbf_stream_create.argtypes = [C.c_int, C.POINTER(_vp)]
_lib.pbf_stream_destroy.argtypes = [C.c_int, _vp]
_lib.pbf_stream_sync.argtypes = [C.c_int, _vp]
_lib.pbf_copy_d2h_async.argtypes = [_vp, _vp, _i64, _vp]
_lib.pbf_device_count.argtypes = [C.POINTER(C.c_int)]
_lib.pbf_slab_register_state.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
_lib.pbf_slab_peer_export.argtypes = [_vp, C.POINTER(SlabPeerInfo)]
_lib.pbf_slab_peer_attach.argtypes = [_vp, C.c_int, C.POINTER(SlabPeerInfo)]
_lib.pbf_slab_halo_sync.argtypes = [_vp]
_lib.pbf_slab_push_state.argtypes = [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
_lib.pbf_last_error.restype = C.c_char_p
_lib.pbf_version.restype = C.c_char_p


def lib():
    return _lib


def _check(rc):
    if rc != OK:
        raise PbfError(rc, _lib.pbf_last_error().decode())


def _f3arr(v):
    a = np.ascontiguousarray(np.asarray(v, np.float32).reshape(3))
    return a, a.ctypes.data_as(_f3)


def _ptr(x):
    """Device pointer of a torch tensor / DeviceBuffer / int; host pointer of a numpy array."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags.c_contiguous
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ptr"):
        return x.ptr
    raise TypeError(type(x))


def version():
    return _lib.pbf_version().decode()


_lib.pbf_state_write.argtypes = [C.c_char_p, C.POINTER(StateInfo), _vp, _vp, _vp]
_lib.pbf_state_read_info.argtypes = [C.c_char_p, C.POINTER(StateInfo)]
_lib.pbf_state_read.argtypes = [C.c_char_p, C.POINTER(StateInfo), _vp, _vp, _vp, _i64]
_lib.pbf_checkpoint_save.argtypes = [_vp, C.c_char_p, _vp, _vp, _vp, _i64, _i64]
_lib.pbf_checkpoint_load.argtypes = [_vp, C.c_char_p, _vp, _vp, _vp, _i64, C.POINTER(_i64), C.POINTER(_i64)]


def state_write(path, pos, vel, iid, params, ulim, llim, frame=0, exact_pow=1):
    """Host arrays -> state file (include/pbf.h pbf_state_write). No device involved."""
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 3)
    iid = np.ascontiguousarray(iid, np.uint32).reshape(-1)
    assert len(pos) == len(vel) == len(iid)
    info = StateInfo()
    info.n, info.frame, info.exact_pow = len(iid), int(frame), int(exact_pow)
    C.memmove(C.byref(info.params), C.byref(params), C.sizeof(GUIParams))
    info.ulim[:] = [float(v) for v in ulim]
    info.llim[:] = [float(v) for v in llim]
    _check(_lib.pbf_state_write(os.fsencode(path), C.byref(info), pos.ctypes.data, vel.ctypes.data, iid.ctypes.data))


def state_info(path):
    info = StateInfo()
    _check(_lib.pbf_state_read_info(os.fsencode(path), C.byref(info)))
    return info


def state_read(path):
    """State file -> (info, pos, vel, iid) as host arrays; verifies the checksum."""
    info = state_info(path)
    n = int(info.n)
    pos, vel, iid = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty(n, np.uint32)
    _check(_lib.pbf_state_read(os.fsencode(path), C.byref(info), pos.ctypes.data, vel.ctypes.data, iid.ctypes.data, n))
    return info, pos, vel, iid


def state_digest(pos, vel, iid, n=None, device=0, stream=None):
    """Order-independent 128-bit digest (sum, xor) of a particle state (include/pbf.h pbf_state_digest_*): numpy
    arrays -> the host routine, torch CUDA tensors / device pointers -> the kernel. Returns two Python ints."""
    d = (C.c_uint64 * 2)()
    if isinstance(iid, np.ndarray):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        iid = np.ascontiguousarray(iid).view(np.uint32)
        n = len(iid) if n is None else int(n)
        _check(_lib.pbf_state_digest_host(pos.ctypes.data, vel.ctypes.data, iid.ctypes.data, n, d))
    else:
        n = int(iid.shape[0]) if n is None else int(n)
        _check(_lib.pbf_state_digest_device(int(device), _ptr(pos), _ptr(vel), _ptr(iid), n, stream, d))
    return int(d[0]), int(d[1])


def combine_digests(digests):
    """Digest of the union of disjoint particle sets (e.g. the slabs of G ranks)."""
    s = x = 0
    for a, b in digests:
        s = (s + int(a)) & 0xFFFFFFFFFFFFFFFF
        x ^= int(b)
    return s, x


def default_params():
    """Defaults the reference writes at FluidSystem.cpp:15-25."""
    p = GUIParams()
    _check(_lib.pbf_default_params(C.byref(p)))
    return p


class DeviceBuffer:
    """cudaMalloc'ed buffer through the C-ABI helpers (for callers that do not use torch)."""

    def __init__(self, nbytes, device=0):
        self.device, self.nbytes = device, int(nbytes)
        p = _vp()
        _check(_lib.pbf_device_alloc(device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _check(_lib.pbf_copy_h2d(self.ptr, arr.ctypes.data, arr.nbytes))
        return self

    def download(self, dtype, shape):
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        _check(_lib.pbf_copy_d2h(out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            _lib.pbf_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Simulator:
    """Mirror of the reference's `Simulator` (fluids/Simulator.h:7-61).

    Simulator(params, ulim, llim)   <- Simulator(const GUIParams&, float3 ulim, float3 llim)
    step(pos, npos, vel, nvel, iid, n) <- step(uint d_pos, ..., int nparticle), on device buffers
    loadParams(params) / saveParams()  <- loadParams() / saveParams() without the singleton
    setLim(ulim, llim)                 <- setLim(const float3&, const float3&)
    """

    def __init__(self, params, ulim, llim, max_particles, device=0):
        self._h = None
        u, up = _f3arr(ulim)
        l, lp = _f3arr(llim)
        h = _vp()
        _check(_lib.pbf_create(C.byref(params), up, lp, int(max_particles), device, C.byref(h)))
        self._h = h.value
        self.device = device
        self.max_particles = int(max_particles)
        self.n = 0

    def close(self):
        if self._h:
            _lib.pbf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters / box
    def loadParams(self, params):
        _check(_lib.pbf_set_params(self._h, C.byref(params)))

    def saveParams(self):
        p = GUIParams()
        _check(_lib.pbf_get_params(self._h, C.byref(p)))
        return p

    def setLim(self, ulim, llim):
        u, up = _f3arr(ulim)
        l, lp = _f3arr(llim)
        _check(_lib.pbf_set_lim(self._h, up, lp))

    def getLim(self):
        u = np.zeros(3, np.float32)
        l = np.zeros(3, np.float32)
        _check(_lib.pbf_get_lim(self._h, u.ctypes.data_as(_f3), l.ctypes.data_as(_f3)))
        return u, l

    def set_option(self, option, value):
        _check(_lib.pbf_set_option(self._h, int(option), int(value)))

    def get_option(self, option):
        v = C.c_int(0)
        _check(_lib.pbf_get_option(self._h, int(option), C.byref(v)))
        return int(v.value)

    def set_exact_pow(self, on):
        _check(_lib.pbf_set_option_exact_pow(self._h, int(bool(on))))

    def const_div_interval(self):
        lo, hi = C.c_float(), C.c_float()
        _check(_lib.pbf_get_const_div_interval(self._h, C.byref(lo), C.byref(hi)))
        return float(lo.value), float(hi.value)

    def pair_list(self):
        """(in use, bytes) of the lambda -> delta-p neighbour list (pbf_get_pair_list): False = the handle was too large
        for it (or PBF_NO_PAIR_REUSE=1) and the delta-p pass repeats the full gather."""
        on, nbytes = C.c_int32(0), C.c_uint64(0)
        _check(_lib.pbf_get_pair_list(self._h, C.byref(on), C.byref(nbytes)))
        return bool(on.value), int(nbytes.value)

    def fast_spiky(self):
        """(in use, mismatches) of the exhaustively verified branch-free spiky scale (pbf_get_fast_spiky)."""
        on, bad = C.c_int32(0), C.c_uint64(0)
        _check(_lib.pbf_get_fast_spiky(self._h, C.byref(on), C.byref(bad)))
        return int(on.value), int(bad.value)

    def trim_pow(self):
        """(in use, mismatches) of the exhaustively verified trimmed powf(w, 4.0f) (pbf_get_trim_pow)."""
        on, bad = C.c_int32(0), C.c_uint64(0)
        _check(_lib.pbf_get_trim_pow(self._h, C.byref(on), C.byref(bad)))
        return int(on.value), int(bad.value)

    def grid_dim(self):
        d = (C.c_int32 * 3)()
        _check(_lib.pbf_get_grid_dim(self._h, d))
        return tuple(d)

    # -- the hot path
    def step(self, pos, npos, vel, nvel, iid, n, stream=None):
        self.n = int(n)
        _check(_lib.pbf_step(self._h, _ptr(pos), _ptr(npos), _ptr(vel), _ptr(nvel), _ptr(iid), int(n), stream))

    def step_host(self, pos, npos, vel, nvel, iid):
        n = len(iid)
        self.n = n
        for a in (pos, npos, vel, nvel):
            assert a.dtype == np.float32 and a.flags.c_contiguous and a.size == 3 * n
        assert iid.dtype == np.uint32 and iid.flags.c_contiguous
        _check(_lib.pbf_step_host(self._h, _ptr(pos), _ptr(npos), _ptr(vel), _ptr(nvel), _ptr(iid), n))

    # -- stages (reference Simulator.h:44-48)
    def begin(self, pos, npos, vel, nvel, iid, n, stream=None):
        self.n = int(n)
        _check(_lib.pbf_stage_begin(self._h, _ptr(pos), _ptr(npos), _ptr(vel), _ptr(nvel), _ptr(iid), int(n), stream))

    def advect(self): _check(_lib.pbf_stage_advect(self._h))
    def buildGridHash(self): _check(_lib.pbf_stage_build_grid(self._h))
    def correctDensity(self): _check(_lib.pbf_stage_correct_density(self._h))
    def computeLambda(self): _check(_lib.pbf_stage_lambda(self._h))     # first half of correctDensity
    def computeDeltaP(self): _check(_lib.pbf_stage_delta_p(self._h))    # second half (+ the Jacobi commit)
    def updateVelocity(self): _check(_lib.pbf_stage_update_velocity(self._h))
    def correctVelocity(self): _check(_lib.pbf_stage_correct_velocity(self._h))
    def end(self): _check(_lib.pbf_stage_end(self._h))

    # -- multi-GPU x-slab decomposition (include/pbf.h "multi-GPU"); the transport lives in slab.py
    def slab_begin(self, step, pos, npos, vel, nvel, iid, stream=None):
        _check(_lib.pbf_slab_begin(self._h, C.byref(step), _ptr(pos), _ptr(npos), _ptr(vel), _ptr(nvel), _ptr(iid), stream))

    def slab_layout(self):
        lay = SlabLayout()
        _check(_lib.pbf_slab_get_layout(self._h, C.byref(lay)))
        self.n = int(lay.n_local)
        return lay

    def slab_plane_counts(self, x_first, count):
        out = np.zeros(count, np.int64)
        _check(_lib.pbf_slab_plane_counts(self._h, int(x_first), int(count), out.ctypes.data_as(C.POINTER(_i64))))
        return out

    def slab_halo(self, what):
        """Device pointers (send_left, recv_left, send_right, recv_right) of one halo refresh."""
        p = [_vp() for _ in range(4)]
        _check(_lib.pbf_slab_halo(self._h, int(what), *[C.byref(q) for q in p]))
        return tuple(q.value or 0 for q in p)

    def slab_adopt_state(self, x_begin, x_end, has_left, has_right, pos, npos, vel, nvel, iid, n, stream=None):
        """Like slab_sort_state, but particles outside planes [x_begin, x_end) are dropped; returns the count kept."""
        kept = _i64()
        _check(_lib.pbf_slab_adopt_state(self._h, int(x_begin), int(x_end), int(bool(has_left)), int(bool(has_right)),
                                         _ptr(pos), _ptr(npos), _ptr(vel), _ptr(nvel), _ptr(iid), int(n), C.byref(kept), stream))
        return int(kept.value)

    def slab_register_state(self, pos_a, pos_b, vel_a, vel_b, iid):
        _check(_lib.pbf_slab_register_state(self._h, _ptr(pos_a), _ptr(pos_b), _ptr(vel_a), _ptr(vel_b), _ptr(iid)))

    def slab_peer_export(self):
        info = SlabPeerInfo()
        _check(_lib.pbf_slab_peer_export(self._h, C.byref(info)))
        return bytes(info)

    def slab_peer_attach(self, side, info_bytes):
        if info_bytes is None:
            _check(_lib.pbf_slab_peer_attach(self._h, int(side), None))
        else:
            info = SlabPeerInfo.from_buffer_copy(info_bytes)
            _check(_lib.pbf_slab_peer_attach(self._h, int(side), C.byref(info)))

    def slab_halo_sync(self):
        _check(_lib.pbf_slab_halo_sync(self._h))

    def slab_push_state(self, left_count, left_dst, right_first, right_dst):
        _check(_lib.pbf_slab_push_state(self._h, int(left_count), int(left_dst), int(right_first), int(right_dst)))

    def slab_flags(self):
        f = C.c_uint32()
        _check(_lib.pbf_slab_flags(self._h, C.byref(f)))
        return int(f.value)

    def slab_sort_state(self, x_begin, x_end, has_left, has_right, pos, npos, vel, nvel, iid, n, stream=None):
        _check(_lib.pbf_slab_sort_state(self._h, int(x_begin), int(x_end), int(bool(has_left)), int(bool(has_right)),
                                        _ptr(pos), _ptr(npos), _ptr(vel), _ptr(nvel), _ptr(iid), int(n), stream))

    # -- read-backs
    def read(self, what, count=None):
        if what in (READ_CELL_START, READ_CELL_END):
            d = self.grid_dim()
            count = d[0] * d[1] * d[2] if count is None else count
            out = np.empty(count, np.uint32)
        elif what in (READ_NPOS, READ_POS0, READ_VEL):
            count = self.n if count is None else count
            out = np.empty((count, 3), np.float32)
        elif what in (READ_LAMBDA, READ_RHO):
            count = self.n if count is None else count
            out = np.empty(count, np.float32)
        else:
            count = self.n if count is None else count
            out = np.empty(count, np.uint32)
        _check(_lib.pbf_read(self._h, what, out.ctypes.data, count))
        return out

    def stats(self, npos, nvel, n):
        st = Stats()
        _check(_lib.pbf_get_stats(self._h, _ptr(npos), _ptr(nvel), int(n), C.byref(st)))
        return st.as_dict()

    def checkpoint_save(self, path, pos, vel, iid, n, frame=0):
        """Device state (what the next step would consume) + parameters + box -> state file."""
        _check(_lib.pbf_checkpoint_save(self._h, os.fsencode(path), _ptr(pos), _ptr(vel), _ptr(iid), int(n), int(frame)))

    def checkpoint_load(self, path, pos, vel, iid, capacity):
        """State file -> device buffers; applies the file's parameters, box and pow option. Returns (n, frame)."""
        n, frame = _i64(0), _i64(0)
        _check(_lib.pbf_checkpoint_load(self._h, os.fsencode(path), _ptr(pos), _ptr(vel), _ptr(iid), int(capacity),
                                        C.byref(n), C.byref(frame)))
        return int(n.value), int(frame.value)

    def enable_stage_timing(self, on=True):
        _check(_lib.pbf_enable_stage_timing(self._h, int(on)))

    def stage_ms(self):
        ms = (C.c_float * 5)()
        _check(_lib.pbf_get_stage_ms(self._h, ms))
        return dict(zip(STAGE_NAMES, list(ms)))

    def kernel_ms(self):
        ms = (C.c_float * len(KERNEL_NAMES))()
        _check(_lib.pbf_get_kernel_ms(self._h, ms))
        return dict(zip(KERNEL_NAMES, list(ms)))

    def launch_count(self):
        return int(_lib.pbf_launch_count(self._h))


# ---- ParticleSource scenes (fluids/ParticleSource.h, DoubleDamSource.*, FixedCubeSource.*) ----

class ParticleSource:
    """initialize() returns host arrays (pos[n,3], vel[n,3], iid[n]); update()/reset() follow the
    reference: update returns the count unchanged, reset == initialize (DoubleDamSource.cpp:43-49)."""

    def initialize(self):
        raise NotImplementedError

    def update(self):
        return self.count

    def reset(self):
        return self.initialize()


class FixedCubeSource(ParticleSource):
    """FixedCubeSource(ulim, llim, ns) (fluids/FixedCubeSource.h:12-20)."""

    def __init__(self, ulim, llim, ns, seed=27):
        self.ulim, self.llim, self.ns, self.seed = ulim, llim, ns, seed
        self.count = 0

    def initialize(self):
        pos, vel, iid, _ = _scene_cubes([(self.ulim, self.llim, self.ns)], self.seed)
        self.count = len(iid)
        return pos, vel, iid


class DoubleDamSource(ParticleSource):
    """DoubleDamSource(ulim1, llim1, ns1, ulim2, llim2, ns2) (fluids/DoubleDamSource.h:11-23)."""

    def __init__(self, ulim1, llim1, ns1, ulim2, llim2, ns2, seed=27):
        self.blocks = [(ulim1, llim1, ns1), (ulim2, llim2, ns2)]
        self.seed = seed
        self.count = 0

    def initialize(self):
        pos, vel, iid, _ = _scene_cubes(self.blocks, self.seed)
        self.count = len(iid)
        return pos, vel, iid


class EmitterSource(ParticleSource):
    """Mirror of host/ParticleSource.h EmitterSource: one ny x nz lattice layer at `origin` (spacing d,
    velocity v0) every `period` calls, appended at the end of the caller's buffers, until `total` particles
    exist. update(pos, vel, iid, capacity) writes into torch / numpy buffers in place and returns the count."""

    def __init__(self, origin, ny, nz, d, v0, period, total):
        self.origin, self.ny, self.nz, self.d = np.asarray(origin, np.float32), int(ny), int(nz), np.float32(d)
        self.v0, self.period, self.total = np.asarray(v0, np.float32), max(int(period), 1), int(total)
        self.count = self.calls = 0

    def layer(self):
        j, k = np.meshgrid(np.arange(self.ny, dtype=np.float32), np.arange(self.nz, dtype=np.float32), indexing="ij")
        pos = np.empty((self.ny * self.nz, 3), np.float32)
        pos[:, 0] = self.origin[0]
        pos[:, 1] = (self.origin[1] + self.d * j).reshape(-1)   # float32 throughout, like the C++ loop
        pos[:, 2] = (self.origin[2] + self.d * k).reshape(-1)
        vel = np.broadcast_to(self.v0, pos.shape).copy()
        iid = np.arange(self.count, self.count + len(pos), dtype=np.uint32)
        return pos, vel, iid

    def initialize(self, pos, vel, iid, capacity):
        self.count = self.calls = 0
        return self.update(pos, vel, iid, capacity)

    def update(self, pos, vel, iid, capacity):
        m = self.ny * self.nz
        if self.calls % self.period == 0 and self.count + m <= min(self.total, int(capacity)):
            p, v, i = self.layer()
            a, b = self.count, self.count + m
            if isinstance(pos, np.ndarray):
                pos[a:b], vel[a:b], iid[a:b] = p, v, i
            else:   # torch tensors on the device
                import torch
                pos[a:b] = torch.from_numpy(p).to(pos.device)
                vel[a:b] = torch.from_numpy(v).to(vel.device)
                iid[a:b] = torch.from_numpy(i.astype(np.int64)).to(iid.device).to(iid.dtype)
            self.count = b
        self.calls += 1
        return self.count

    def reset(self, pos, vel, iid, capacity):
        return self.initialize(pos, vel, iid, capacity)


def _scene_cubes(blocks, seed):
    total = sum(int(np.prod(b[2])) for b in blocks)
    pos = np.zeros((total, 3), np.float32)
    vel = np.zeros((total, 3), np.float32)
    iid = np.zeros(total, np.uint32)
    rng = C.c_uint32(seed)
    off = 0
    for ulim, llim, ns in blocks:
        u, up = _f3arr(ulim)
        l, lp = _f3arr(llim)
        nsa = (C.c_int32 * 3)(*[int(v) for v in ns])
        cnt = _i64()
        _check(_lib.pbf_scene_cube(up, lp, nsa, C.byref(rng), off, pos[off:].ctypes.data, vel[off:].ctypes.data,
                                   iid[off:].ctypes.data, total - off, C.byref(cnt)))
        off += cnt.value
    return pos, vel, iid, rng.value


def scene_double_dam_reference():
    """The reference's shipped scene (FluidSystem.cpp:34-35,55-61): returns pos, vel, iid, ulim, llim."""
    pos = np.zeros((32000, 3), np.float32)
    vel = np.zeros((32000, 3), np.float32)
    iid = np.zeros(32000, np.uint32)
    ulim = np.zeros(3, np.float32)
    llim = np.zeros(3, np.float32)
    cnt = _i64()
    _check(_lib.pbf_scene_double_dam_reference(pos.ctypes.data, vel.ctypes.data, iid.ctypes.data, 32000, C.byref(cnt),
                                               ulim.ctypes.data_as(_f3), llim.ctypes.data_as(_f3)))
    assert cnt.value == 32000
    return pos, vel, iid, ulim, llim


def scene_block_host(origin, n3, spacing=0.05, seed=27, first_iid=0):
    o, op = _f3arr(origin)
    n3a = (C.c_int32 * 3)(*[int(v) for v in n3])
    total = int(n3[0]) * int(n3[1]) * int(n3[2])
    pos = np.zeros((total, 3), np.float32)
    vel = np.zeros((total, 3), np.float32)
    iid = np.zeros(total, np.uint32)
    _check(_lib.pbf_scene_block_host(op, n3a, spacing, seed, first_iid, pos.ctypes.data, vel.ctypes.data, iid.ctypes.data))
    return pos, vel, iid


def scene_block_device(origin, n3, d_pos, d_vel, d_iid, spacing=0.05, seed=27, first_iid=0, stream=None):
    o, op = _f3arr(origin)
    n3a = (C.c_int32 * 3)(*[int(v) for v in n3])
    _check(_lib.pbf_scene_block_device(op, n3a, spacing, seed, first_iid, _ptr(d_pos), _ptr(d_vel), _ptr(d_iid), stream))
    return int(n3[0]) * int(n3[1]) * int(n3[2])


def scene_block_slice_device(origin, n3, ix_begin, ix_end, d_pos, d_vel, d_iid, spacing=0.05, seed=27, first_iid=0,
                             stream=None):
    """Lattice layers [ix_begin, ix_end) of the block, bit-identical to the full block's particles."""
    o, op = _f3arr(origin)
    n3a = (C.c_int32 * 3)(*[int(v) for v in n3])
    _check(_lib.pbf_scene_block_slice_device(op, n3a, spacing, seed, first_iid, int(ix_begin), int(ix_end),
                                             _ptr(d_pos), _ptr(d_vel), _ptr(d_iid), stream))
    return (int(ix_end) - int(ix_begin)) * int(n3[1]) * int(n3[2])


def scene_block_slice_host(origin, n3, ix_begin, ix_end, spacing=0.05, seed=27, first_iid=0):
    o, op = _f3arr(origin)
    n3a = (C.c_int32 * 3)(*[int(v) for v in n3])
    total = (int(ix_end) - int(ix_begin)) * int(n3[1]) * int(n3[2])
    pos = np.zeros((total, 3), np.float32)
    vel = np.zeros((total, 3), np.float32)
    iid = np.zeros(total, np.uint32)
    _check(_lib.pbf_scene_block_slice_host(op, n3a, spacing, seed, first_iid, int(ix_begin), int(ix_end),
                                           pos.ctypes.data, vel.ctypes.data, iid.ctypes.data))
    return pos, vel, iid


from .scenes import SCENES, wall_lim, scene_dims, scene_particles  # noqa: E402,F401  (pure numpy: bench.py's reference arm loads that file alone)
