"""CPU engine for the slab protocol: the ORACLE behind the same engine interface as
pbf-cuda_b200/slab.py::GpuEngine. TEST INFRASTRUCTURE ONLY — it exists so that the multi-rank host logic
(planning, message sizing, sort order, ghost refreshes) can be checked at world_size 2 over gloo on a
machine without a GPU, against the single-domain oracle step. The product never imports this file.

Differences from the GPU engine that do not matter to the protocol: the oracle keeps the global cell
table (no local keys), computes the passes for ghost slots too (the values are overwritten by the halo
refresh from the owner) and moves lambda / position / rho as separate tight arrays instead of float4.
"""
import ctypes as C
import importlib

import numpy as np
import torch

import _oracle as O

slab = importlib.import_module("pbf-cuda_b200.slab")
HALO_LAMBDA, HALO_POSITION, HALO_VELOCITY = slab.HALO_LAMBDA, slab.HALO_POSITION, slab.HALO_VELOCITY


class Layout:
    pass


class OracleEngine:
    def __init__(self, params, ulim, llim, capacity, threads=2):
        self.params = params
        self.ulim, self.llim = np.asarray(ulim, np.float32), np.asarray(llim, np.float32)
        self.capacity = int(capacity)
        self.o = O.Oracle(params, ulim, llim, self.capacity, threads=threads)
        h = np.float32(params.h)
        self.dims = [int(np.ceil(np.float32(self.ulim[a] - self.llim[a]) / h)) for a in range(3)]
        self.planes = self.dims[0]
        self.pos = np.zeros((self.capacity, 3), np.float32)
        self.vel = np.zeros((self.capacity, 3), np.float32)
        self.iid = np.zeros(self.capacity, np.uint32)
        self.n_own = 0
        self._own_planes = np.zeros(0, np.int64)
        self.layout = None

    def _cells(self, p):
        c = [slab.plane_of(p[:, a], self.llim[a], self.params.h, self.dims[a]) for a in range(3)]
        return c

    def load_state(self, pos, vel, iid, x_begin, x_end, has_left, has_right):
        pos, vel, iid = np.asarray(pos, np.float32), np.asarray(vel, np.float32), np.asarray(iid, np.uint32)
        n = len(iid)
        cx, cy, cz = self._cells(pos)
        assert ((cx >= x_begin) & (cx < x_end)).all()
        key = (cx * self.dims[1] + cy) * self.dims[2] + cz
        order = np.argsort(key, kind="stable")
        self.pos[:n], self.vel[:n], self.iid[:n] = pos[order], vel[order], iid[order]
        self._own_planes = cx[order]
        self.n_own = n

    def plane_counts(self):
        return np.bincount(self._own_planes, minlength=self.planes).astype(np.int64)

    def raw_views(self, lo, hi):
        return [torch.from_numpy(self.pos[lo:hi]), torch.from_numpy(self.vel[lo:hi]),
                torch.from_numpy(self.iid[lo:hi].view(np.int32))]

    def begin(self, step):
        self.step = step
        assert step.n_own + step.m_left + step.m_right <= self.capacity

    def grid(self):
        st = self.step
        n, ml, mr = st.n_own, st.m_left, st.m_right
        # the sort's logical order: [from left | own | from right]
        order = np.concatenate([np.arange(n, n + ml), np.arange(0, n), np.arange(n + ml, n + ml + mr)])
        pos, vel, iid = self.pos[order].copy(), self.vel[order].copy(), self.iid[order].copy()
        npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
        self.o.bind(pos, npos, vel, nvel, iid)
        self.o.advect()
        lo = st.x_begin - st.ghost if st.has_left else 0
        hi = st.x_end + st.ghost if st.has_right else self.planes
        lo, hi = max(lo, 0), min(hi, self.planes)
        cx = slab.plane_of(npos[:, 0], self.llim[0], self.params.h, self.planes)
        keep = (cx >= lo) & (cx < hi)
        pos, vel, iid, npos, nvel = pos[keep], vel[keep], iid[keep], npos[keep], nvel[keep]
        self.o.bind(pos, npos, vel, nvel, iid)
        self.o.build_grid()
        self.a = dict(pos=pos, vel=vel, iid=iid, npos=npos, nvel=nvel)
        cx = slab.plane_of(npos[:, 0], self.llim[0], self.params.h, self.planes)
        assert (np.diff(cx) >= 0).all()
        self._planes_sorted = cx
        L = Layout()
        L.n_local = len(iid)
        L.own_first = int(np.searchsorted(cx, st.x_begin, side="left"))
        own_end = int(np.searchsorted(cx, st.x_end, side="left"))
        L.own_count = own_end - L.own_first
        L.recv_left_count = L.own_first
        L.recv_right_count = L.n_local - own_end
        gl = min(st.ghost, st.x_end - st.x_begin) if st.has_left else 0
        gr = min(st.ghost, st.x_end - st.x_begin) if st.has_right else 0
        L.send_left_count = int(np.searchsorted(cx, st.x_begin + gl, side="left")) - L.own_first
        L.send_right_count = own_end - int(np.searchsorted(cx, st.x_end - gr, side="left"))
        L.flags = 0
        self.layout = L
        self._own_planes = cx[L.own_first:own_end]
        lib = O.lib()
        self.lam = np.ctypeslib.as_array(C.cast(lib.orc_lambda(self.o.h), C.POINTER(C.c_float)), shape=(L.n_local,))
        self.pho = np.ctypeslib.as_array(C.cast(lib.orc_pho(self.o.h), C.POINTER(C.c_float)), shape=(L.n_local,))
        return L

    def lambda_pass(self): O.lib().orc_lambda_pass(self.o.h)
    def delta_p_pass(self): O.lib().orc_delta_p_pass(self.o.h)
    def update_velocity(self): self.o.update_velocity()
    def xsph(self): self.o.correct_velocity()

    def halo(self, what):
        L = self.layout
        f, e = L.own_first, L.own_first + L.own_count
        if what == HALO_LAMBDA:
            arrs = [self.lam]
        elif what == HALO_POSITION:
            arrs = [self.a["npos"]]
        else:
            arrs = [self.pho]     # a ghost's velocity is recomputed locally from exact copies; rho is not
        T = torch.from_numpy
        return ([T(a[f:f + L.send_left_count]) for a in arrs], [T(a[0:f]) for a in arrs],
                [T(a[e - L.send_right_count:e]) for a in arrs], [T(a[e:L.n_local]) for a in arrs])

    def end(self):
        L = self.layout
        f, e = L.own_first, L.own_first + L.own_count
        n = L.own_count
        self.pos[:n], self.vel[:n], self.iid[:n] = self.a["npos"][f:e], self.a["nvel"][f:e], self.a["iid"][f:e]
        self.n_own = n
        return n

    def flags(self):
        return 0

    def state(self):
        n = self.n_own
        return self.pos[:n], self.vel[:n], self.iid[:n]

    def close(self):
        self.o.close()


def sorted_global_state(pos, vel, iid, llim, h, dims):
    """The global initial state in the stable cell order of its positions — the order every slab run
    starts from, and the order the single-domain comparison run is given."""
    c = [slab.plane_of(pos[:, a], llim[a], h, dims[a]) for a in range(3)]
    key = (c[0] * dims[1] + c[1]) * dims[2] + c[2]
    order = np.argsort(key, kind="stable")
    return pos[order].copy(), vel[order].copy(), iid[order].copy(), c[0][order]


def small_dam(nx=40, ny=8, nz=14, box=(3.2, 0.6, 1.2), origin=(0.35, 0.05, 0.05)):
    """A dam-break block in a small box: 32 planes along x, enough for two or three slabs."""
    pos, vel, iid = O.scene_block(origin, (nx, ny, nz), 0.05, 27, 0)
    ulim = np.asarray(box, np.float32)
    llim = np.zeros(3, np.float32)
    return pos, vel, iid, ulim, llim
