"""Stage-by-stage traces of the PBF step from three implementations with one common format:

    trace_oracle(scene)     CPU restatement (oracle/libpbf_oracle.so)            — runs anywhere
    trace_reference(scene)  the reference's own Simulator.cu (oracle/_ref)       — needs a GPU
    trace_product(scene)    libpbf_b200.so through the C-ABI stage entry points  — needs a GPU

A trace is a dict name -> ndarray; per step `s`:
    s{s}.key, s{s}.iid, s{s}.start, s{s}.end          after buildGridHash (sorted order)
    s{s}.npos0                                         advected positions, sorted order
    s{s}.lam{k}, s{s}.pho{k}, s{s}.tpos{k}             after Jacobi iteration k
    s{s}.vel                                           after updateVelocity
    s{s}.nvel, s{s}.npos                               after correctVelocity (step outputs)
All three start from the same host state and follow the caller protocol of
FluidSystem::stepSimulate (reference FluidSystem.cpp:99-120): setLim for a moving wall, step,
swap pos<->npos and vel<->nvel.
"""
import ctypes as C

import numpy as np

import _oracle as O


def make_scene(name):
    """Small named parity scenes. Returns dict(params, ulim, llim, pos, vel, iid, steps, wall)."""
    p = O.default_params()
    wall = None
    if name == "dd32k":  # the reference's own scene (FluidSystem.cpp:55-61)
        pos, vel, iid, ulim, llim = O.scene_double_dam_reference()
        steps = 2
    elif name == "cube2k":  # one suspended block in a small box; full dumps stay small
        ulim, llim = np.float32([1.2, 1.0, 1.5]), np.float32([0, 0, 0])
        pos, vel, iid = O.scene_cube([0.9, 0.8, 1.3], [0.3, 0.2, 0.5], [12, 12, 16])
        steps = 2
    elif name == "floor2k":  # block resting on the floor in a corner: boundary clamps + boundary density on
        ulim, llim = np.float32([1.0, 1.0, 1.0]), np.float32([0, 0, 0])
        pos, vel, iid = O.scene_cube([0.62, 0.62, 0.82], [0.02, 0.02, 0.02], [12, 12, 16])
        p.k_boundaryDensity = 0.5
        steps = 2
    elif name == "wall2k":  # moving wall (FluidSystem.cpp:104-110) squeezing a block, non-cubic box
        ulim, llim = np.float32([0.85, 0.7, 1.1]), np.float32([0, 0, 0])
        pos, vel, iid = O.scene_cube([0.65, 0.62, 0.82], [0.05, 0.02, 0.02], [12, 12, 16])
        wall = dict(a_ulim=np.float32([-0.2, 0, 0]), a_llim=np.float32([0, 0, 0]), w=np.float32(0.7), start=-1)
        steps = 3
    elif name == "ragged":  # n not a multiple of anything, two blocks of different size, niter 3, n_corr 3
        ulim, llim = np.float32([1.3, 0.9, 1.0]), np.float32([-0.2, 0, 0])
        a = O.scene_cube([0.5, 0.5, 0.8], [0.0, 0.1, 0.3], [7, 9, 11], seed=27)
        b = O.scene_cube([1.2, 0.8, 0.6], [0.8, 0.3, 0.1], [5, 10, 9], seed=5, first_iid=len(a[2]))
        pos = np.concatenate([a[0], b[0]]); vel = np.concatenate([a[1], b[1]]); iid = np.concatenate([a[2], b[2]])
        vel[:, 0] = 0.3  # non-zero initial velocity
        p.niter = 3
        p.n_corr = 3.0
        steps = 2
    else:
        raise KeyError(name)
    return dict(name=name, params=p, ulim=ulim, llim=llim, pos=pos, vel=vel, iid=iid, steps=steps, wall=wall)


def lim_for_step(scene, s):
    if scene["wall"] is None:
        return scene["ulim"], scene["llim"]
    w = scene["wall"]
    return O.wall_lim(scene["ulim"], scene["llim"], w["a_ulim"], w["a_llim"], float(w["w"]), s, w["start"])


def max_box(scene):
    u = scene["ulim"].copy()
    if scene["wall"] is not None:
        u = u + np.abs(scene["wall"]["a_ulim"])
    return u, scene["llim"]


def trace_oracle(scene, threads=4):
    p = scene["params"]
    n = len(scene["iid"])
    mu, ml = max_box(scene)
    o = O.Oracle(p, mu, ml, n, threads=threads)
    pos, vel, iid = scene["pos"].copy(), scene["vel"].copy(), scene["iid"].copy()
    npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
    T = {}
    for s in range(scene["steps"]):
        u, l = lim_for_step(scene, s)
        o.set_lim(u, l)
        o.bind(pos, npos, vel, nvel, iid)
        o.advect()
        o.build_grid()
        T["s%d.key" % s] = o.grid_id(); T["s%d.iid" % s] = iid.copy()
        T["s%d.start" % s] = o.grid_start(); T["s%d.end" % s] = o.grid_end()
        T["s%d.npos0" % s] = npos.copy()
        for k in range(p.niter):
            o.correct_density()
            T["s%d.lam%d" % (s, k)] = o.lam(); T["s%d.pho%d" % (s, k)] = o.pho(); T["s%d.tpos%d" % (s, k)] = o.tpos()
        o.update_velocity()
        T["s%d.vel" % s] = vel.copy()
        o.correct_velocity()
        T["s%d.nvel" % s] = nvel.copy(); T["s%d.npos" % s] = npos.copy()
        pos, npos = npos, pos
        vel, nvel = nvel, vel
    o.close()
    return T


def _torch_state(scene):
    import torch
    dev = torch.device("cuda:0")
    pos = torch.from_numpy(scene["pos"].copy()).to(dev)
    vel = torch.from_numpy(scene["vel"].copy()).to(dev)
    iid = torch.from_numpy(scene["iid"].astype(np.int64)).to(dev).to(torch.int32)  # bit pattern of uint32
    npos, nvel = torch.zeros_like(pos), torch.zeros_like(vel)
    return pos, npos, vel, nvel, iid


def _u32(t):
    return t.cpu().numpy().view(np.uint32).copy()


def trace_reference(scene):
    """The reference's own kernels, stage by stage (needs oracle/_ref/libpbf_ref.so and a GPU)."""
    import torch
    import _ref
    p = scene["params"]
    n = len(scene["iid"])
    mu, ml = max_box(scene)
    r = _ref.RefSimulator(p, mu, ml, n)
    pos, npos, vel, nvel, iid = _torch_state(scene)
    T = {}
    for s in range(scene["steps"]):
        u, l = lim_for_step(scene, s)
        r.set_lim(u, l)
        r.bind(pos, npos, vel, nvel, iid, n)
        r.stage(r.ADVECT)
        r.stage(r.GRID)
        T["s%d.key" % s] = r.grid_id(); T["s%d.iid" % s] = _u32(iid)
        T["s%d.start" % s] = r.grid_start(); T["s%d.end" % s] = r.grid_end()
        T["s%d.npos0" % s] = npos.cpu().numpy().copy()
        for k in range(p.niter):
            r.stage(r.DENSITY)
            T["s%d.lam%d" % (s, k)] = r.lam(); T["s%d.pho%d" % (s, k)] = r.pho(); T["s%d.tpos%d" % (s, k)] = r.tpos()
        r.stage(r.VELOCITY_UPDATE)
        T["s%d.vel" % s] = vel.cpu().numpy().copy()
        r.stage(r.VELOCITY_CORRECT)
        torch.cuda.synchronize()
        T["s%d.nvel" % s] = nvel.cpu().numpy().copy(); T["s%d.npos" % s] = npos.cpu().numpy().copy()
        pos, npos = npos, pos
        vel, nvel = nvel, vel
    r.close()
    return T


def trace_product(scene, pbf, exact_pow=True, use_step=False):
    """libpbf_b200.so through the C-ABI. use_step=True runs the fused pbf_step instead of the stage
    entry points (then only the step outputs and the grid are recorded)."""
    import torch
    p = scene["params"]
    gp = pbf.GUIParams()
    C.memmove(C.byref(gp), C.byref(p), C.sizeof(gp))
    n = len(scene["iid"])
    mu, ml = max_box(scene)
    sim = pbf.Simulator(gp, mu, ml, n)
    sim.set_exact_pow(exact_pow)
    pos, npos, vel, nvel, iid = _torch_state(scene)
    T = {}
    for s in range(scene["steps"]):
        u, l = lim_for_step(scene, s)
        sim.setLim(u, l)
        if use_step:
            sim.step(pos, npos, vel, nvel, iid, n)
            torch.cuda.synchronize()
            T["s%d.key" % s] = sim.read(pbf.READ_KEY); T["s%d.iid" % s] = _u32(iid)
            T["s%d.start" % s] = sim.read(pbf.READ_CELL_START); T["s%d.end" % s] = sim.read(pbf.READ_CELL_END)
            T["s%d.pho%d" % (s, p.niter - 1)] = sim.read(pbf.READ_RHO)
            T["s%d.vel" % s] = vel.cpu().numpy().copy()
            T["s%d.pos" % s] = pos.cpu().numpy().copy()
        else:
            sim.begin(pos, npos, vel, nvel, iid, n)
            sim.advect()
            sim.buildGridHash()
            T["s%d.key" % s] = sim.read(pbf.READ_KEY); T["s%d.iid" % s] = sim.read(pbf.READ_IID)
            T["s%d.src" % s] = sim.read(pbf.READ_SRC_INDEX)
            T["s%d.start" % s] = sim.read(pbf.READ_CELL_START); T["s%d.end" % s] = sim.read(pbf.READ_CELL_END)
            T["s%d.npos0" % s] = sim.read(pbf.READ_NPOS)
            T["s%d.ncount" % s] = sim.read(pbf.READ_NEIGHBOR_COUNT)
            for k in range(p.niter):
                sim.correctDensity()
                T["s%d.lam%d" % (s, k)] = sim.read(pbf.READ_LAMBDA); T["s%d.pho%d" % (s, k)] = sim.read(pbf.READ_RHO)
                T["s%d.tpos%d" % (s, k)] = sim.read(pbf.READ_NPOS)
            sim.updateVelocity()
            T["s%d.vel" % s] = sim.read(pbf.READ_VEL)
            sim.correctVelocity()
            sim.end()
            torch.cuda.synchronize()
            T["s%d.pos" % s] = pos.cpu().numpy().copy()
        T["s%d.nvel" % s] = nvel.cpu().numpy().copy(); T["s%d.npos" % s] = npos.cpu().numpy().copy()
        T["s%d.iid_out" % s] = _u32(iid)
        pos, npos = npos, pos
        vel, nvel = nvel, vel
    sim.close()
    return T


def compare(A, B, keys=None):
    """Per-field comparison of two traces: returns {field: dict(exact, max_abs, max_rel_norm)}."""
    out = {}
    for k in sorted(A.keys() if keys is None else keys):
        if k not in B:
            continue
        a, b = A[k], B[k]
        if a.shape != b.shape:
            out[k] = dict(exact=False, max_abs=float("inf"), rel=float("inf"), shape=(a.shape, b.shape))
            continue
        if a.dtype.kind in "iu":
            out[k] = dict(exact=bool(np.array_equal(a, b)), max_abs=float(np.abs(a.astype(np.int64) - b.astype(np.int64)).max() if a.size else 0), rel=0.0)
        else:
            d = np.abs(a.astype(np.float64) - b.astype(np.float64))
            scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
            out[k] = dict(exact=bool(np.array_equal(a, b)), max_abs=float(d.max() if d.size else 0), rel=float((d.max() if d.size else 0) / scale))
    return out


def format_report(title, cmp):
    lines = [title]
    for k, v in cmp.items():
        lines.append("  %-12s exact=%-5s max_abs=%.3e rel_norm=%.3e" % (k, v["exact"], v["max_abs"], v["rel"]))
    return "\n".join(lines)
