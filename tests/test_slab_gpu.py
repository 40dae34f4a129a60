"""GPU tests of the multi-GPU path on ONE device: `world` ranks run as threads of this process, each with its
own pbf_sim handle (pbf-cuda_b200/slab.py GpuEngine) and an in-process transport (ThreadComm). The protocol,
the kernels and the C-ABI calls are exactly those of the NCCL run (bench.py --gpus N); only the byte mover
differs. Bar: the ranks' results, concatenated in rank order, equal the single-GPU pbf_step's arrays BIT FOR
BIT — same particles, same cell-sorted order, same positions and velocities — after several steps with
migration across the boundaries and a re-plan."""
import importlib
import threading

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; the product has no CPU path")
    return torch


def _scene(name):
    if name == "small":
        pos, vel, iid = O.scene_block((0.35, 0.05, 0.05), (40, 8, 14), 0.05, 27, 0)
        return pos, vel, iid, np.asarray((3.2, 0.6, 1.2), np.float32), np.zeros(3, np.float32)
    if name == "dam260k":
        pos, vel, iid = O.scene_block((0.2, 0.2, 0.2), (128, 32, 64), 0.05, 27, 0)
        return pos, vel, iid, np.asarray((12.8, 2.0, 4.8), np.float32), np.zeros(3, np.float32)
    raise KeyError(name)


def _sorted_state(pbf, slab, pos, vel, iid, ulim, llim, h):
    dims = [int(np.ceil(np.float32(ulim[a] - llim[a]) / np.float32(h))) for a in range(3)]
    c = [slab.plane_of(pos[:, a], llim[a], h, dims[a]) for a in range(3)]
    key = (c[0] * dims[1] + c[1]) * dims[2] + c[2]
    order = np.argsort(key, kind="stable")
    return pos[order].copy(), vel[order].copy(), iid[order].copy(), c[0][order], dims


def _single_gpu(pbf, torch, pos, vel, iid, ulim, llim, steps):
    dev = torch.device("cuda:0")
    n = len(iid)
    d = [torch.from_numpy(pos).to(dev), torch.zeros((n, 3), device=dev), torch.from_numpy(vel).to(dev), torch.zeros((n, 3), device=dev)]
    d_iid = torch.from_numpy(iid.view(np.int32)).to(dev)
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    for _ in range(steps):
        sim.step(d[0], d[1], d[2], d[3], d_iid, n)
        d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
    torch.cuda.synchronize()
    out = d[0].cpu().numpy(), d[2].cpu().numpy(), d_iid.cpu().numpy().view(np.uint32)
    sim.close()
    return out


def _run_ranks(pbf, slab, torch, scene, world, steps, ghost, margin, replan_every, skew=False, vel_override=None, fused=False):
    pos, vel, iid, ulim, llim = scene
    if vel_override is not None:
        vel = vel_override
    p = pbf.default_params()
    gpos, gvel, giid, gplane, dims = _sorted_state(pbf, slab, pos, vel, iid, ulim, llim, p.h)
    hub = slab.ThreadComm.Hub(world)
    results, errors = [None] * world, [None] * world
    dev = torch.device("cuda:0")

    def worker(rank):
        try:
            # every emulated rank on its own stream, like every real rank on its own device: the fused halo's
            # flag wait must not sit in front of the neighbour's signal in one shared stream
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                eng = slab.GpuEngine(pbf, p, ulim, llim, len(giid), device_index=0, stream=stream.cuda_stream)
                sim = slab.SlabSimulator(eng, slab.ThreadComm(hub, rank), p.niter, dims[0], ghost=ghost, margin=margin,
                                         replan_every=replan_every, fused_halo=fused)
                sim.plan_initial(np.bincount(gplane, minlength=dims[0]))
                if skew:
                    sim.bounds = [0] + [sim.min_width * r for r in range(1, world)] + [dims[0]]
                x0, x1 = sim.my_planes()
                mine = (gplane >= x0) & (gplane < x1)
                sim.load_owned(torch.from_numpy(gpos[mine]).to(dev), torch.from_numpy(gvel[mine]).to(dev),
                               torch.from_numpy(giid[mine].view(np.int32)).to(dev))
                bounds = [list(sim.bounds)]
                for _ in range(steps):
                    sim.step()
                    bounds.append(list(sim.bounds))
                sim.finish()
                stream.synchronize()
                sp, sv, si = eng.state()
                results[rank] = dict(pos=sp.cpu().numpy(), vel=sv.cpu().numpy(), iid=si.cpu().numpy().view(np.uint32),
                                     bounds=bounds, messages=sim.messages, launches=eng.sim.launch_count(),
                                     pushed=sim.pushed_steps)
                hub.barrier.wait(timeout=120)   # nobody frees arrays a neighbour may still be pushing into
                eng.close()
        except BaseException as ex:   # noqa: BLE001 - reported by the main thread
            errors[rank] = ex
            hub.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    return results, errors, (gpos, gvel, giid)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name,world,replan_every,skew", [
    ("small", 2, 0, False), ("small", 3, 2, True), ("dam260k", 2, 0, False), ("dam260k", 4, 3, True)])
def test_slab_ranks_equal_single_gpu_bit_for_bit(pbf, torch, name, world, replan_every, skew, fused):
    # fused: the ghost refreshes are peer-memory stores from inside the pass kernels + a flag handshake
    # (here between handles of one process on one device; between processes the same pointers come from CUDA IPC)
    slab = importlib.import_module("pbf-cuda_b200.slab")
    scene = _scene(name)
    steps = 6
    margin = 2 if name == "small" else 4
    results, errors, (gpos, gvel, giid) = _run_ranks(pbf, slab, torch, scene, world, steps, 2, margin, replan_every, skew, fused=fused)
    for ex in errors:
        if ex is not None:
            raise ex
    ref_pos, ref_vel, ref_iid = _single_gpu(pbf, torch, gpos, gvel, giid, scene[3], scene[4], steps)
    pos = np.concatenate([r["pos"] for r in results])
    vel = np.concatenate([r["vel"] for r in results])
    iid = np.concatenate([r["iid"] for r in results])
    assert len(iid) == len(ref_iid)
    assert np.array_equal(iid, ref_iid)
    assert np.array_equal(pos, ref_pos)
    assert np.array_equal(vel, ref_vel)
    assert all(r["launches"] > 0 for r in results)
    # fused: from the second step on the raw state of the boundary planes arrives by the neighbours' own stores
    # (pbf_slab_push_state), through re-plans too; otherwise by the transport
    assert all(r["pushed"] == (steps - 1 if fused else 0) for r in results)
    if skew:
        b = results[0]["bounds"]
        assert any(row != b[0] for row in b), "the re-plan never moved a boundary"


def test_slab_world1_equals_plain_step(pbf, torch):
    """The slab entry points on a single rank (no neighbours) are the plain step."""
    slab = importlib.import_module("pbf-cuda_b200.slab")
    scene = _scene("small")
    p = pbf.default_params()
    gpos, gvel, giid, gplane, dims = _sorted_state(pbf, slab, *scene, p.h)
    dev = torch.device("cuda:0")
    eng = slab.GpuEngine(pbf, p, scene[3], scene[4], len(giid))
    sim = slab.SlabSimulator(eng, slab.SingleComm(), p.niter, dims[0])
    sim.plan_initial(np.bincount(gplane, minlength=dims[0]))
    sim.load_owned(torch.from_numpy(gpos).to(dev), torch.from_numpy(gvel).to(dev), torch.from_numpy(giid.view(np.int32)).to(dev))
    for _ in range(3):
        sim.step()
    sim.finish()
    sp, sv, si = eng.state()
    ref = _single_gpu(pbf, torch, gpos, gvel, giid, scene[3], scene[4], 3)
    assert np.array_equal(sp.cpu().numpy(), ref[0]) and np.array_equal(sv.cpu().numpy(), ref[1])
    assert np.array_equal(si.cpu().numpy().view(np.uint32), ref[2])
    eng.close()


def test_slab_flags_a_particle_that_outruns_the_margin(pbf, torch):
    """A particle moving more planes per step than `margin` covers is not silently lost: the step raises."""
    slab = importlib.import_module("pbf-cuda_b200.slab")
    scene = _scene("small")
    pos, vel = scene[0], scene[1].copy()
    ghost, margin = 2, 1
    plane = slab.plane_of(pos[:, 0], 0.0, 0.1, 32)
    b = slab.plan_boundaries(np.bincount(plane, minlength=32), 2, 2 * (ghost + margin))[1]
    # particles of rank 0 that are NOT in the range sent right (planes < b - reach) ...
    fast = np.nonzero((plane >= b - ghost - margin - 3) & (plane < b - ghost - margin - 1))[0][:16]
    assert len(fast) == 16
    vel[fast, 0] = 60.0   # ... jump 0.5 units = 5 planes in one step, into the planes rank 1 stores
    results, errors, _ = _run_ranks(pbf, slab, torch, scene, 2, 2, ghost, margin, 0, vel_override=vel)
    assert any(isinstance(ex, slab.SlabError) and "margin" in str(ex) for ex in errors), errors


def test_scene_block_slice_matches_full_block(pbf, torch):
    dev = torch.device("cuda:0")
    origin, n3 = (0.2, 0.2, 0.2), (24, 6, 10)
    fp, fv, fi = pbf.scene_block_host(origin, n3)
    sp, sv, si = pbf.scene_block_slice_host(origin, n3, 5, 17)
    per = n3[1] * n3[2]
    assert np.array_equal(sp, fp[5 * per:17 * per]) and np.array_equal(si, fi[5 * per:17 * per])
    d_pos = torch.zeros((12 * per, 3), device=dev); d_vel = torch.ones_like(d_pos)
    d_iid = torch.zeros(12 * per, dtype=torch.int32, device=dev)
    assert pbf.scene_block_slice_device(origin, n3, 5, 17, d_pos, d_vel, d_iid) == 12 * per
    torch.cuda.synchronize()
    assert np.array_equal(d_pos.cpu().numpy(), sp) and np.array_equal(d_iid.cpu().numpy().view(np.uint32), si)
    assert float(d_vel.abs().max()) == 0.0


def test_slab_over_nccl_equals_single_gpu(pbf, torch):
    """The same comparison through the real transport (one process per GPU, NCCL send/recv); needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the one-GPU box covers the protocol with thread-emulated ranks)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = min(4, torch.cuda.device_count())
    for mode in ("nccl", "fused"):   # NCCL send/recv halos, then fused peer-memory halos over CUDA IPC
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                            "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "slab_nccl_check.py"),
                            "8", "3", mode], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "BIT-EXACT" in r.stdout, mode + r.stdout[-2000:] + r.stderr[-2000:]


def test_cpp_slab_harness_matches_single_gpu(pbf, torch, tmp_path):
    """host/SlabSimulator.h (C++, one thread per rank, fused transport between handles of one process): 1, 2 and
    3 ranks write the same bytes, and they are the bytes of pbf_step on the cell-sorted scene."""
    import os
    import subprocess
    slab = importlib.import_module("pbf-cuda_b200.slab")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "pbf-cuda_b200", "pbf_slab_headless")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "pbf-cuda_b200"), "harness"], stdout=subprocess.DEVNULL)
    steps, dumps = 6, {}
    for ranks in (1, 2, 3):
        out = str(tmp_path / ("r%d.bin" % ranks))
        r = subprocess.run([exe, "--scene", "small", "--ranks", str(ranks), "--steps", str(steps), "--warmup", "0", "--ghost", "2",
                            "--margin", "2", "--replan", "2", "--dump", out], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        dumps[ranks] = open(out, "rb").read()
    assert dumps[1] == dumps[2] == dumps[3]
    n = int(np.frombuffer(dumps[1][:4], np.int32)[0])
    body = np.frombuffer(dumps[1][4:], np.float32)
    pos, vel = body[:3 * n].reshape(n, 3), body[3 * n:6 * n].reshape(n, 3)
    iid = np.frombuffer(dumps[1][4 + 24 * n:], np.uint32)
    scene = _scene("small")
    gpos, gvel, giid, _, _ = _sorted_state(pbf, slab, *scene, pbf.default_params().h)
    ref = _single_gpu(pbf, torch, gpos, gvel, giid, scene[3], scene[4], steps)
    assert n == len(giid)
    assert np.array_equal(iid, ref[2]) and np.array_equal(pos, ref[0]) and np.array_equal(vel, ref[1])
