"""Multi-rank host logic on CPU: the slab protocol of pbf-cuda_b200/slab.py at world_size 2 and 3 over
gloo, with the oracle as the per-rank engine (tests/_slab_cpu.py), must reproduce the single-domain
oracle step BIT FOR BIT for every particle — positions, velocities, the per-rank cell order — through
migration across the slab boundary and a re-plan of the boundaries. Plus the planner's own invariants."""
import importlib
import os
import socket

import numpy as np
import pytest

import _oracle as O

slab = importlib.import_module("pbf-cuda_b200.slab")


# ---- planner -----------------------------------------------------------------------------------

def test_plan_quantiles_and_min_width():
    t = np.zeros(64, np.int64)
    t[4:36] = 100                      # all the mass in planes 4..35 (a dam in the corner)
    b = slab.plan_boundaries(t, 4, 6)
    assert b[0] == 0 and b[-1] == 64 and len(b) == 5
    assert all(b[r + 1] - b[r] >= 6 for r in range(4))
    per = [int(t[b[r]:b[r + 1]].sum()) for r in range(4)]
    assert max(per) - min(per) <= 100   # balanced to one plane
    with pytest.raises(slab.SlabError):
        slab.plan_boundaries(t, 16, 6)


def test_plan_respects_reach_of_old_boundaries():
    t = np.zeros(64, np.int64)
    t[40:60] = 10
    old = [0, 16, 32, 48, 64]
    b = slab.plan_boundaries(t, 4, 8, old=old, reach=6)
    for r in range(1, 4):
        assert old[r - 1] + 6 <= b[r] <= old[r + 1] - 6
    assert all(b[r + 1] - b[r] >= 8 for r in range(4))


def test_exchange_plan_is_symmetric():
    rng = np.random.default_rng(3)
    planes, world, reach = 48, 3, 4
    old = [0, 16, 32, 48]
    new = [0, 14, 33, 48]
    counts = np.zeros((world, planes), np.int64)
    for r in range(world):
        counts[r, old[r]:old[r + 1]] = rng.integers(0, 50, old[r + 1] - old[r])
    xp = [slab.exchange_plan(counts, old, new, r, reach) for r in range(world)]
    for r in range(world - 1):
        n_r = int(counts[r].sum())
        assert n_r - xp[r]["send_right_begin"] == xp[r + 1]["m_left"]
        assert xp[r + 1]["send_left_end"] == xp[r]["m_right"]
    assert xp[0]["send_left_end"] == 0 and xp[0]["m_left"] == 0
    assert xp[-1]["m_right"] == 0


def test_plane_of_matches_oracle_keys():
    pos, vel, iid, ulim, llim = __import__("_slab_cpu").small_dam()
    p = O.default_params()
    o = O.Oracle(p, ulim, llim, len(iid))
    npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
    o.bind(pos.copy(), npos, vel.copy(), nvel, iid.copy())
    o.advect(); o.build_grid()
    d = o.grid_dim()
    assert np.array_equal(o.grid_id() // (d[1] * d[2]), slab.plane_of(npos[:, 0], llim[0], p.h, d[0]))
    o.close()


# ---- the protocol, multi-process over gloo ---------------------------------------------------------

def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, steps, replan_every, ghost, margin, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import _slab_cpu as S
    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, vel, iid, ulim, llim = S.small_dam()
        p = O.default_params()
        eng = S.OracleEngine(p, ulim, llim, len(iid))
        comm = slab.TorchComm(dist)
        sim = slab.SlabSimulator(eng, comm, p.niter, eng.planes, ghost=ghost, margin=margin, replan_every=replan_every)
        gpos, gvel, giid, gplane = S.sorted_global_state(pos, vel, iid, llim, p.h, eng.dims)
        sim.plan_initial(np.bincount(gplane, minlength=eng.planes))
        if replan_every:   # start from a deliberately lopsided cut so that the re-plan has work to do
            sim.bounds = [0] + [8 * r for r in range(1, world)] + [eng.planes]
        x0, x1 = sim.my_planes()
        mine = (gplane >= x0) & (gplane < x1)
        sim.load_owned(gpos[mine], gvel[mine], giid[mine])
        bounds_seen = [list(sim.bounds)]
        for _ in range(steps):
            sim.step()
            bounds_seen.append(list(sim.bounds))
        sim.finish()
        spos, svel, siid = eng.state()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), pos=spos, vel=svel, iid=siid,
                 bounds=np.asarray(bounds_seen), total=sim.total_particles(), messages=sim.messages)
    finally:
        dist.destroy_process_group()


def _single_domain(steps):
    import _slab_cpu as S
    pos, vel, iid, ulim, llim = S.small_dam()
    p = O.default_params()
    o = O.Oracle(p, ulim, llim, len(iid), threads=4)
    d = [int(np.ceil(np.float32(ulim[a] - llim[a]) / np.float32(p.h))) for a in range(3)]
    pos, vel, iid, _ = S.sorted_global_state(pos, vel, iid, llim, p.h, d)
    npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
    for _ in range(steps):
        o.step(pos, npos, vel, nvel, iid)
        pos, npos, vel, nvel = npos, pos, nvel, vel
    o.close()
    return pos, vel, iid


@pytest.mark.parametrize("world,replan_every", [(2, 0), (2, 3), (3, 2)])
def test_slab_protocol_bit_exact_over_gloo(tmp_path, world, replan_every):
    import torch.multiprocessing as mp
    steps = 8
    port = _free_port()
    mp.spawn(_worker, args=(world, port, steps, replan_every, 2, 2, str(tmp_path)), nprocs=world, join=True)
    ref_pos, ref_vel, ref_iid = _single_domain(steps)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    # the ranks' results, concatenated in rank order, ARE the single-domain arrays: same particles in the
    # same (cell-sorted, stable) order with the same bits
    pos = np.concatenate([q["pos"] for q in parts])
    vel = np.concatenate([q["vel"] for q in parts])
    iid = np.concatenate([q["iid"] for q in parts])
    assert int(parts[0]["total"]) == len(ref_iid) == len(iid)
    assert np.array_equal(iid, ref_iid)
    assert np.array_equal(pos, ref_pos)
    assert np.array_equal(vel, ref_vel)
    assert all(int(q["messages"]) > 0 for q in parts)
    if replan_every:
        b = parts[0]["bounds"]
        assert any(not np.array_equal(b[0], row) for row in b), "the re-plan never moved a boundary"
