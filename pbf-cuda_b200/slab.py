"""slab.py — one scene on G GPUs: the x-slab decomposition of the PBF step.

The reference is single-GPU (one `Simulator`, fluids/Simulator.h:7-61); this module is what makes G
`Simulator` handles on G devices advance ONE scene so that every particle receives exactly the bits the
single-GPU `pbf_step` gives it (tests/test_slab_*.py). One process per GPU; the transport is
`torch.distributed` (NCCL send/recv over NVLink on the GPU box, gloo in the CPU tests); the per-rank work
is the library's slab entry points (include/pbf.h "multi-GPU"). Three layers:

  Planner        where the slab boundaries are (particle-count quantiles over cell planes), replicated:
                 every rank holds every rank's per-plane particle counts (one small all-gather per step)
                 and computes the same plan and the same message sizes — no size handshake on the wire.
  SlabSimulator  the step protocol: raw-state exchange of the planes around each boundary, one local
                 stable sort, then per Jacobi iteration a lambda halo and a position halo, a velocity
                 halo before XSPH — each halo one contiguous float4 range per side.
  engine         the per-rank compute behind the protocol: `GpuEngine` (libpbf_b200.so through the C-ABI;
                 the product) — the CPU tests plug the oracle in instead (tests/_slab_cpu.py) to check the
                 protocol itself at world_size 2 over gloo.

Why the result is bit-identical to one GPU: keys are x-major (reference Simulator.cu:45-53), so a rank's
planes are contiguous slot ranges; the within-cell order of the single-GPU stable sort is the previous
global order, and a rank sorts [from left | own | from right], which IS the previous global order
restricted to the particles it sees; ghosts are exact copies refreshed from their owner after every pass.
"""
import numpy as np

from . import (HALO_LAMBDA, HALO_POSITION, HALO_VELOCITY, SLAB_FLAG_GHOST, SLAB_FLAG_MIGRATION, SLAB_FLAG_TIMEOUT,
               SLAB_STATE_PUSHED, SlabPeerInfo, SlabStep)
import ctypes as _C


class SlabError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------------
# planning (pure host logic, identical on every rank)
# ---------------------------------------------------------------------------------------------------

def plan_boundaries(plane_totals, world, min_width, old=None, reach=None):
    """Slab boundaries b[0..world] (b[0] = 0, b[world] = planes) cutting `plane_totals` at particle-count
    quantiles, every slab at least `min_width` planes wide. With `old` boundaries and a `reach`, boundary r
    stays inside [old[r-1] + reach, old[r+1] - reach] so that everything a rank needs next step is held by
    itself or an adjacent rank; returns `old` unchanged if that cannot be met."""
    t = np.asarray(plane_totals, np.int64)
    planes = len(t)
    if world * min_width > planes:
        raise SlabError("%d planes cannot hold %d slabs of >= %d planes" % (planes, world, min_width))
    cum = np.concatenate([[0], np.cumsum(t)])
    total = int(cum[-1])
    b = [0]
    for r in range(1, world):
        target = total * r / world
        x = int(np.searchsorted(cum, target, side="left"))
        # the plane boundary closest to the quantile
        if x > 0 and abs(cum[x - 1] - target) <= abs(cum[min(x, planes)] - target):
            x -= 1
        b.append(x)
    b.append(planes)
    lo = [0] + [r * min_width for r in range(1, world)] + [planes]
    hi = [0] + [planes - (world - r) * min_width for r in range(1, world)] + [planes]
    if old is not None and reach is not None:
        for r in range(1, world):
            lo[r] = max(lo[r], old[r - 1] + reach)
            hi[r] = min(hi[r], old[r + 1] - reach)
    for r in range(1, world):            # forward: clamp and keep the minimum width
        b[r] = min(max(b[r], lo[r], b[r - 1] + min_width), hi[r])
    for r in range(world - 1, 0, -1):    # backward
        b[r] = max(min(b[r], b[r + 1] - min_width), lo[r])
    ok = all(b[r + 1] - b[r] >= min_width for r in range(world)) and all(lo[r] <= b[r] <= hi[r] for r in range(1, world))
    if not ok:
        if old is not None:
            return list(old)
        raise SlabError("no valid slab plan for %d ranks over %d planes" % (world, planes))
    return b


def exchange_plan(counts, old, new, rank, reach):
    """What rank `rank` sends and receives in the raw-state exchange. counts[r][x] = particles rank r owns in
    plane x (previous step); old/new = boundaries of the previous / this step; reach = ghost + margin planes.
    Returns dict(send_left_end, send_right_begin, m_left, m_right) in own-slot units. Every rank evaluates
    the same formulas on the same replicated table, so sender and receiver agree without a handshake."""
    world = len(old) - 1
    own = counts[rank]
    off = np.concatenate([[0], np.cumsum(own)])   # off[x] = first own slot of plane x

    def clip(x, lo, hi):
        return min(max(x, lo), hi)

    n_own = int(off[-1])
    send_left_end, send_right_begin, m_left, m_right = 0, n_own, 0, 0
    if rank > 0:
        # the left rank will store planes up to new[rank] + ghost; its particles come from <= margin further
        send_left_end = int(off[clip(new[rank] + reach, old[rank], old[rank + 1])])
        first = clip(new[rank] - reach, old[rank - 1], old[rank])
        m_left = int(np.sum(counts[rank - 1][first:old[rank]]))
        if new[rank] - reach < old[rank - 1]:
            raise SlabError("plan moves boundary %d beyond the left neighbour's slab" % rank)
    if rank < world - 1:
        send_right_begin = int(off[clip(new[rank + 1] - reach, old[rank], old[rank + 1])])
        last = clip(new[rank + 1] + reach, old[rank + 1], old[rank + 2])
        m_right = int(np.sum(counts[rank + 1][old[rank + 1]:last]))
        if new[rank + 1] + reach > old[rank + 2]:
            raise SlabError("plan moves boundary %d beyond the right neighbour's slab" % (rank + 1))
    # (fused mode pulls instead of receiving: the left neighbour's range starts at ITS send_right_begin)
    pull_left_first = 0
    if rank > 0:
        off_l = np.concatenate([[0], np.cumsum(counts[rank - 1])])
        pull_left_first = int(off_l[clip(new[rank] - reach, old[rank - 1], old[rank])])
    return dict(send_left_end=send_left_end, send_right_begin=send_right_begin, m_left=m_left, m_right=m_right,
                pull_left_first=pull_left_first)


# ---------------------------------------------------------------------------------------------------
# transport
# ---------------------------------------------------------------------------------------------------

class TorchComm:
    """Neighbour exchange + small all-gather over torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, dist, device=None, group=None):
        self.dist, self.device, self.group = dist, device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def exchange(self, sends, recvs):
        """sends / recvs: {peer: [tensor, ...]} (empty tensors are skipped on both sides)."""
        ops = []
        for peer in sorted(set(sends) | set(recvs)):
            for t in sends.get(peer, ()):
                if t.numel():
                    ops.append(self.dist.P2POp(self.dist.isend, t, peer, self.group))
            for t in recvs.get(peer, ()):
                if t.numel():
                    ops.append(self.dist.P2POp(self.dist.irecv, t, peer, self.group))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()

    def connect(self, peers):
        """Exchanges one word with every peer in BOTH directions. NCCL sets a send/recv connection up on
        first use per direction (~0.2 s); a dam break may not send anything leftwards for many steps, so
        without this the set-up lands in the middle of a timed run."""
        import torch
        dev = self.device if self.device is not None else "cpu"
        tx = {p: [torch.zeros(1, dtype=torch.float32, device=dev)] for p in peers}
        rx = {p: [torch.zeros(1, dtype=torch.float32, device=dev)] for p in peers}
        self.exchange(tx, rx)

    def allgather_counts(self, counts):
        import torch
        mine = torch.from_numpy(np.ascontiguousarray(counts, np.int64))
        if self.device is None:
            out = torch.empty(self.world * mine.numel(), dtype=torch.int64)
            self.dist.all_gather_into_tensor(out, mine, group=self.group)
            return out.numpy().reshape(self.world, -1)
        # on a stream of its own: the collective and the host's wait for it do not queue behind the pass
        # kernels already enqueued on the compute stream
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._side):
            mine = mine.to(self.device, non_blocking=True)
            out = torch.empty(self.world * mine.numel(), dtype=torch.int64, device=self.device)
            self.dist.all_gather_into_tensor(out, mine, group=self.group)
            host = out.cpu()
        return host.numpy().reshape(self.world, -1)

    _side = None
    _pin = [None, None]

    def allgather_counts_start(self, counts):
        """The same all-gather without waiting for it: returns a handle whose wait() gives the table. On GPUs the
        collective and the download into pinned memory run on the transport's own stream; nothing on the compute
        stream and nothing on the host waits until wait() is called — one step later (SlabSimulator.step)."""
        if self.device is None:
            return _Done(self.allgather_counts(counts))
        import torch
        mine = torch.from_numpy(np.ascontiguousarray(counts, np.int64))
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._pin = [None, None]
        with torch.cuda.stream(self._side):
            dev = mine.to(self.device, non_blocking=True)
            out = torch.empty(self.world * mine.numel(), dtype=torch.int64, device=self.device)
            self.dist.all_gather_into_tensor(out, dev, group=self.group)
            # (two pinned buffers in rotation: pinning memory every step would cost more than the exchange)
            k = self._pin_next = (getattr(self, "_pin_next", 0) + 1) % 2
            if self._pin[k] is None or self._pin[k].numel() != out.numel():
                self._pin[k] = torch.empty(out.shape, dtype=torch.int64).pin_memory()
            host = self._pin[k]
            host.copy_(out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._side)
        return _Pending(ev, host, self.world, (mine, dev, out))


class _Done:
    def __init__(self, table):
        self.table = table

    def wait(self):
        return self.table


class _Pending:
    def __init__(self, ev, host, world, keep):
        self.ev, self.host, self.world, self.keep = ev, host, world, keep

    def wait(self):
        self.ev.synchronize()
        self.keep = None
        return self.host.numpy().reshape(self.world, -1).copy()


class SingleComm:
    """world == 1: nothing to exchange (the slab code path on one rank, for tests and N=1 runs)."""
    rank, world = 0, 1

    def exchange(self, sends, recvs):
        assert not any(t.numel() for ts in sends.values() for t in ts)

    def connect(self, peers):
        pass

    def allgather_counts(self, counts):
        return np.asarray(counts, np.int64)[None, :]

    def allgather_counts_start(self, counts):
        return _Done(self.allgather_counts(counts))


class ThreadComm:
    """`world` ranks as threads of ONE process sharing one device: the same protocol with in-process
    mailboxes. Lets the one-GPU test box run the multi-rank path bit for bit (tests/test_slab_gpu.py)."""

    class Hub:
        def __init__(self, world):
            import queue
            import threading
            self.world = world
            self.box = {(a, b): queue.Queue() for a in range(world) for b in range(world)}
            self.barrier = threading.Barrier(world)
            self.gather = [None] * world
            self.failed = threading.Event()   # a rank died: the others stop waiting for its messages

        def abort(self):
            self.failed.set()
            self.barrier.abort()

    def __init__(self, hub, rank):
        self.hub, self.rank, self.world = hub, rank, hub.world

    def exchange(self, sends, recvs):
        # ranks may run on different CUDA streams of the one device: the copy is ordered behind the
        # sender's clone with an event
        for peer in sorted(sends):
            for t in sends[peer]:
                if t.numel():
                    c, ev = t.clone(), None
                    if c.is_cuda:
                        import torch
                        ev = torch.cuda.Event()
                        ev.record()
                    self.hub.box[(self.rank, peer)].put((c, ev))
        for peer in sorted(recvs):
            for t in recvs[peer]:
                if t.numel():
                    c, ev = self._get(self.hub.box[(peer, self.rank)])
                    if ev is not None:
                        import torch
                        torch.cuda.current_stream().wait_event(ev)
                    t.copy_(c)

    def connect(self, peers):
        pass

    def _get(self, box):
        import queue
        for _ in range(1200):
            try:
                return box.get(timeout=0.1)
            except queue.Empty:
                if self.hub.failed.is_set():
                    raise SlabError("a peer rank failed")
        raise SlabError("timed out waiting for a peer rank")

    def allgather_counts(self, counts):
        self.hub.gather[self.rank] = np.asarray(counts, np.int64).copy()
        self.hub.barrier.wait(timeout=120)
        out = np.stack(self.hub.gather)
        self.hub.barrier.wait(timeout=120)
        return out

    def allgather_counts_start(self, counts):
        return _Done(self.allgather_counts(counts))


# ---------------------------------------------------------------------------------------------------
# the product's engine: libpbf_b200.so through the C-ABI
# ---------------------------------------------------------------------------------------------------

class _DevView:
    """Zero-copy torch view of a library-owned device range (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class GpuEngine:
    """Per-rank compute of the slab protocol on one GPU. Owns the five caller-side particle arrays
    (capacity = max particles this rank may hold incl. received and ghost particles)."""

    def __init__(self, pbf, params, ulim, llim, capacity, device_index=0, stream=None):
        import torch
        self.torch, self.pbf = torch, pbf
        self.dev = torch.device("cuda", device_index)
        self.capacity = int(capacity)
        self.sim = pbf.Simulator(params, ulim, llim, self.capacity, device=device_index)
        self.stream = stream
        # the five state arrays are plain cudaMalloc allocations (not sub-blocks of torch's caching allocator):
        # as allocation bases they can be handed to neighbour processes through CUDA IPC (fused mode)
        self._bufs = [pbf.DeviceBuffer(self.capacity * 12, device_index) for _ in range(4)] + \
                     [pbf.DeviceBuffer(self.capacity * 4, device_index)]
        v3 = lambda b: torch.as_tensor(_DevView(b.ptr, (self.capacity, 3)), device=self.dev)
        self.pos, self.npos, self.vel, self.nvel = (v3(b) for b in self._bufs[:4])
        self.iid = torch.as_tensor(_DevView(self._bufs[4].ptr, (self.capacity,), "<i4"), device=self.dev)
        for t in (self.pos, self.npos, self.vel, self.nvel, self.iid):
            t.zero_()
        self.sim.slab_register_state(self.pos, self.npos, self.vel, self.nvel, self.iid)
        self.n_own = 0
        self.layout = None
        self.planes = self.sim.grid_dim()[0]

    # -- state
    def load_state(self, pos, vel, iid, x_begin, x_end, has_left, has_right):
        """Adopt `n` particles (device tensors, any order, all inside the owned planes) and cell-sort them."""
        n = int(iid.shape[0])
        if n > self.capacity:
            raise SlabError("rank holds %d particles, capacity %d" % (n, self.capacity))
        self.pos[:n].copy_(pos); self.vel[:n].copy_(vel); self.iid[:n].copy_(iid)
        self.sim.slab_sort_state(x_begin, x_end, has_left, has_right, self.pos, self.npos, self.vel, self.nvel,
                                 self.iid, n, self.stream)
        self._swap()
        self.n_own = n

    def _swap(self):
        self.pos, self.npos = self.npos, self.pos
        self.vel, self.nvel = self.nvel, self.vel

    def plane_counts(self):
        return self.sim.slab_plane_counts(0, self.planes)

    def raw_views(self, lo, hi):
        return [self.pos[lo:hi], self.vel[lo:hi], self.iid[lo:hi]]

    # -- the step
    def begin(self, step):
        n_in = step.n_own + step.m_left + step.m_right
        if n_in > self.capacity:
            raise SlabError("rank would hold %d particles, capacity %d" % (n_in, self.capacity))
        self.sim.slab_begin(step, self.pos, self.npos, self.vel, self.nvel, self.iid, self.stream)

    def grid(self):
        self.sim.advect()
        self.sim.buildGridHash()
        self.layout = self.sim.slab_layout()
        return self.layout

    def lambda_pass(self): self.sim.computeLambda()
    def delta_p_pass(self): self.sim.computeDeltaP()
    def update_velocity(self): self.sim.updateVelocity()
    def xsph(self): self.sim.correctVelocity()

    def halo(self, what):
        """([send_left], [recv_left], [send_right], [recv_right]) tensors viewing the solver's arrays."""
        L = self.layout
        ptrs = self.sim.slab_halo(what)
        cnts = (L.send_left_count, L.recv_left_count, L.send_right_count, L.recv_right_count)
        out = []
        for p, c in zip(ptrs, cnts):
            if c > 0:
                out.append([self.torch.as_tensor(_DevView(p, (int(c), 4)), device=self.dev)])
            else:
                out.append([])
        return out

    def end(self):
        try:
            self.sim.end()
        except self.pbf.PbfError as ex:   # pbf_stage_end reports a raised TIMEOUT / GHOST flag as an error
            raise SlabError(str(ex)) from ex
        self._swap()
        self.n_own = int(self.layout.own_count)
        return self.n_own

    # -- fused halo over peer memory (the kernels push their boundary values into the neighbours' ghost slots)
    peer_info_bytes = _C.sizeof(SlabPeerInfo)

    def peer_export(self):
        t = self.torch.frombuffer(bytearray(self.sim.slab_peer_export()), dtype=self.torch.uint8)
        return t.to(self.dev)

    def peer_buffer(self):
        return self.torch.zeros(self.peer_info_bytes, dtype=self.torch.uint8, device=self.dev)

    def peer_attach(self, side, info_tensor):
        self.sim.slab_peer_attach(side, None if info_tensor is None else bytes(info_tensor.cpu().numpy().tobytes()))

    def halo_sync(self):
        self.sim.slab_halo_sync()

    def push_state(self, left_count, left_dst, right_first, right_dst):
        """Arms the fused raw-state hand-over of this step (include/pbf.h pbf_slab_push_state)."""
        self.sim.slab_push_state(left_count, left_dst, right_first, right_dst)

    def flags(self):
        return self.sim.slab_flags()

    def mark(self):
        """A CUDA event on the stream the step runs on (phase timing of SlabSimulator.profile_step)."""
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def state(self):
        """(pos, vel, iid) of the owned particles after the last step (device tensors, views)."""
        n = self.n_own
        return self.pos[:n], self.vel[:n], self.iid[:n]

    def close(self):
        self.sim.close()
        for b in self._bufs:
            b.free()


# ---------------------------------------------------------------------------------------------------
# the protocol
# ---------------------------------------------------------------------------------------------------

class SlabSimulator:
    """One rank of a G-rank run. `engine` does the rank's compute, `comm` moves the bytes.

    ghost   ghost planes either side (2: a particle may drift one cell from its stored cell during the
            Jacobi iterations — the reference re-derives the home cell from the current iterate,
            Simulator_kernel.cuh:70,148,212 — and still find all 27 cells; violations raise)
    margin  planes a particle may travel between two sorts (advection + drift); the raw-state exchange
            covers ghost + margin planes either side of a boundary; violations raise
    replan_every  re-cut the slabs at particle-count quantiles every k steps (0: keep the first plan)
    """

    def __init__(self, engine, comm, niter, planes, ghost=2, margin=4, replan_every=0, fused_halo=False):
        self.e, self.c = engine, comm
        self.rank, self.world = comm.rank, comm.world
        self.niter, self.planes = int(niter), int(planes)
        self.ghost, self.margin, self.reach = int(ghost), int(margin), int(ghost) + int(margin)
        self.min_width = 2 * self.reach
        self.replan_every = int(replan_every)
        # fused_halo: the ghost refreshes are peer-memory stores issued by the pass kernels themselves plus a
        # flag handshake (include/pbf.h "Fused halo refresh"); otherwise one send/recv pair per side through comm
        self.fused = bool(fused_halo) and comm.world > 1
        import os
        self.push_state = os.environ.get("PBF_SLAB_PUSH", "1") != "0"   # (A/B switch: 0 = pbf_slab_begin pulls)
        self.bounds = None
        self.counts = None
        self.steps = 0
        self.messages = 0
        self.bytes_sent = 0

    # -- start-up
    def plan_initial(self, plane_totals):
        self.bounds = plan_boundaries(plane_totals, self.world, self.min_width if self.world > 1 else 1)
        return self.bounds

    def my_planes(self):
        return self.bounds[self.rank], self.bounds[self.rank + 1]

    def load_owned(self, pos, vel, iid):
        """pos/vel/iid: this rank's particles (all inside my_planes()), in the GLOBAL input order."""
        x0, x1 = self.my_planes()
        self.c.connect([p for p in (self.rank - 1, self.rank + 1) if 0 <= p < self.world])
        if self.fused:
            self._attach_peers()
        self.e.load_state(pos, vel, iid, x0, x1, self.rank > 0, self.rank < self.world - 1)
        self._gather_counts()

    def _attach_peers(self):
        """Maps the neighbours' arrays (CUDA IPC / same-process pointers). If ANY rank cannot (no peer access,
        IPC disabled in the container ...), every rank detaches and the run uses the comm transport instead —
        said in `self.fused_note`, never silently."""
        e, r, w = self.e, self.rank, self.world
        mine = e.peer_export()
        peers = [p for p in (r - 1, r + 1) if 0 <= p < w]
        got = {p: e.peer_buffer() for p in peers}
        self.c.exchange({p: [mine] for p in peers}, {p: [got[p]] for p in peers})
        err = ""
        try:
            if r > 0:
                e.peer_attach(0, got[r - 1])
            if r < w - 1:
                e.peer_attach(1, got[r + 1])
        except Exception as ex:   # noqa: BLE001 - reported through fused_note, all ranks fall back together
            err = str(ex)
        ok = self.c.allgather_counts(np.asarray([0 if err else 1], np.int64))
        if int(ok.min()) == 0:
            for side in (0, 1):
                try:
                    e.peer_attach(side, None)
                except Exception:   # noqa: BLE001
                    pass
            self.fused = False
            self.fused_note = "fused halo unavailable (%s): using the comm transport" % (err or "a neighbour rank could not attach")

    fused_note = ""

    def _gather_counts(self):
        """Replicates every rank's per-plane particle counts (sizes of the next step's raw exchange, input of
        the planner). The exchange is only STARTED here; the table is needed at the beginning of the next step
        (`_counts()`), a whole step of device work later, so neither the host nor the compute stream waits for the
        slowest rank in the middle of a step."""
        start = getattr(self.c, "allgather_counts_start", None)
        mine = np.asarray(self.e.plane_counts(), np.int64)
        self._pending_counts = start(mine) if start else _Done(self.c.allgather_counts(mine))

    _pending_counts = None

    def _counts(self):
        if self._pending_counts is not None:
            self.counts = self._pending_counts.wait()
            self._pending_counts = None
        return self.counts

    # -- one step
    def _xchg(self, left_send, left_recv, right_send, right_recv):
        sends, recvs = {}, {}
        if self.rank > 0:
            sends[self.rank - 1], recvs[self.rank - 1] = left_send, left_recv
        if self.rank < self.world - 1:
            sends[self.rank + 1], recvs[self.rank + 1] = right_send, right_recv
        for ts in sends.values():
            for t in ts:
                if t.numel():
                    self.messages += 1
                    self.bytes_sent += t.numel() * t.element_size()
        self.c.exchange(sends, recvs)

    def profile_step(self):
        """One step with device-time stamps at the phase boundaries: returns {phase: ms} for this rank
        (raw exchange, keys+sort+layout, count all-gather, passes, halos). Synchronises; not for timed runs."""
        marks = []
        self._mark = lambda name: marks.append((name, self.e.mark()))
        self._mark("start")
        try:
            self.step()
        finally:
            self._mark = None
        marks[-1][1].synchronize()
        out = {}
        for (_, a), (name, b) in zip(marks, marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    _mark = None

    def _m(self, name):
        if self._mark:
            self._mark(name)

    def step(self, after_velocity=None):
        """One step of this rank. after_velocity: optional callable run once the velocity update is enqueued —
        the step's positions (engine.npos[:layout.own_count]) are final from that point of the stream on, so a
        caller that wants them on the host can start the download under the XSPH sweep (bench.py's e2e leg)."""
        e, r, w = self.e, self.rank, self.world
        old = self.bounds
        pushed = False
        if self._planned is not None:    # planned during the last step; the neighbours' kernels delivered the raw state
            new, xp = self._planned
            self._planned = None
            pushed = True
            self.pushed_steps += 1
        else:
            new = old
            if self.replan_every and self.steps and self.steps % self.replan_every == 0 and w > 1:
                new = plan_boundaries(self._counts().sum(axis=0), w, self.min_width, old=old, reach=self.reach)
            xp = exchange_plan(self._counts(), old, new, r, self.reach)
        n_own = e.n_own
        assert n_own == int(self.counts[r].sum())
        m_l, m_r = xp["m_left"], xp["m_right"]
        # 1. raw state of the planes around each boundary: sent / received through comm — or, fused, pulled by
        #    pbf_slab_begin straight out of the neighbours' state arrays (peer-memory copies, no collective)
        if not self.fused:
            self._xchg(e.raw_views(0, xp["send_left_end"]), e.raw_views(n_own, n_own + m_l),
                       e.raw_views(xp["send_right_begin"], n_own), e.raw_views(n_own + m_l, n_own + m_l + m_r))
            self._m("raw_exchange")
        st = SlabStep(x_begin=new[r], x_end=new[r + 1], ghost=self.ghost, has_left=int(r > 0), has_right=int(r < w - 1),
                      n_own=n_own, m_left=m_l, m_right=m_r, send_left_end=xp["send_left_end"],
                      send_right_begin=xp["send_right_begin"],
                      pull_left_first=SLAB_STATE_PUSHED if pushed else xp["pull_left_first"])
        # 2. keys, one stable sort, layout (the step's one host synchronisation on the compute stream)
        e.begin(st)
        if self.fused:
            self._m("raw_exchange")
        lay = e.grid()
        self._m("keys_sort_layout")
        self._raise_flags(lay.flags)
        # 3. Jacobi iterations with ghost refreshes. Behind the FIRST lambda pass — so that the device has
        #    work while the host waits — the new per-plane counts are replicated (on the transport's own
        #    stream): they size the next step's raw exchange and feed the planner. Fetching them here keeps
        #    the END of the step free of any synchronisation: the host runs ahead into the next step.
        for it in range(self.niter):
            e.lambda_pass()
            self._m("lambda")
            if it == 0:
                self._gather_counts()
                self._m("count_allgather")
            self._halo(HALO_LAMBDA)
            self._m("halo")
            e.delta_p_pass()
            self._m("delta_p")
            self._halo(HALO_POSITION)
            self._m("halo")
        if self.niter == 0:
            self._gather_counts()
        if self.fused and w > 1 and self.push_state and hasattr(e, "push_state"):
            self._plan_next(new)
        e.update_velocity()
        self._m("update_velocity")
        if after_velocity is not None:
            after_velocity()
        self._halo(HALO_VELOCITY)
        self._m("halo")
        e.xsph()
        self._m("xsph")
        e.end()
        self.bounds = new
        self.steps += 1

    # fused raw-state hand-over: on by default with the fused halo (off: pbf_slab_begin pulls, six peer copies per step)
    push_state = True
    pushed_steps = 0     # steps whose raw state came by the neighbours' stores (all but the first in fused mode)
    _planned = None

    def _plan_next(self, cur):
        """The NEXT step's plan, made before this step's velocity update is enqueued: the per-plane counts of this
        step's sort are replicated by now (the all-gather was started behind the first lambda pass, a millisecond
        of device work ago), and the plan is a function of that table alone. Knowing it here lets this step's
        velocity / XSPH kernels store the raw state of the boundary planes straight into the neighbours' next
        input arrays (pbf_slab_push_state) — where each range lands follows from the same table: behind the
        neighbour's own particles, and behind what ITS left neighbour delivers on its right-hand side."""
        e, r, w = self.e, self.rank, self.world
        counts = self._counts()
        nxt = cur
        if self.replan_every and (self.steps + 1) % self.replan_every == 0:
            nxt = plan_boundaries(counts.sum(axis=0), w, self.min_width, old=cur, reach=self.reach)
        xp = exchange_plan(counts, cur, nxt, r, self.reach)
        n_next = int(counts[r].sum())
        left_dst = right_dst = 0
        if r > 0:
            xl = exchange_plan(counts, cur, nxt, r - 1, self.reach)
            assert xl["m_right"] == xp["send_left_end"], (xl, xp)
            left_dst = int(counts[r - 1].sum()) + xl["m_left"]
        if r < w - 1:
            xr = exchange_plan(counts, cur, nxt, r + 1, self.reach)
            assert xr["m_left"] == n_next - xp["send_right_begin"] and xr["pull_left_first"] == xp["send_right_begin"], (xr, xp)
            right_dst = int(counts[r + 1].sum())
        e.push_state(xp["send_left_end"] if r > 0 else 0, left_dst, xp["send_right_begin"] if r < w - 1 else n_next, right_dst)
        self._planned = (nxt, xp)

    def _halo(self, what):
        if self.world == 1:
            return
        if self.fused:      # the pass kernel already stored the values into the neighbours' ghost slots
            self.e.halo_sync()
            return
        sl, rl, sr, rr = self.e.halo(what)
        self._xchg(sl, rl, sr, rr)

    def finish(self):
        """Synchronise and raise if any kernel of the steps so far reported a violated assumption. (During
        a run the flags are looked at once per step, at the layout synchronisation, i.e. one step late for
        the ghost check — without an extra synchronisation.)"""
        self._raise_flags(self.e.flags())

    def _raise_flags(self, f):
        if f & SLAB_FLAG_MIGRATION:
            raise SlabError("rank %d: a particle travelled more than margin=%d planes in one step; results are "
                            "not exact — raise `margin`" % (self.rank, self.margin))
        if f & SLAB_FLAG_TIMEOUT:
            raise SlabError("rank %d: a neighbour's halo completion flag did not arrive (fused halo handshake timed out)" % self.rank)
        if f & SLAB_FLAG_GHOST:
            raise SlabError("rank %d: a particle drifted beyond the %d ghost planes during the Jacobi iterations; "
                            "results are not exact — raise `ghost`" % (self.rank, self.ghost))

    def total_particles(self):
        return int(self._counts().sum())


def plane_of(x, llim_x, h, planes):
    """Cell plane of x-coordinates exactly as the library computes it (reference getGridxyz,
    Simulator.cu:30-35): trunc((x - llim) / h) in fp32, clamped. Works on numpy arrays and torch tensors."""
    if isinstance(x, np.ndarray):
        q = ((x.astype(np.float32) - np.float32(llim_x)) / np.float32(h)).astype(np.float32)
        return np.clip(np.trunc(q).astype(np.int64), 0, planes - 1)
    import torch
    q = (x - torch.tensor(llim_x, dtype=torch.float32, device=x.device)) / torch.tensor(h, dtype=torch.float32, device=x.device)
    return torch.clamp(torch.trunc(q).to(torch.int64), 0, planes - 1)
