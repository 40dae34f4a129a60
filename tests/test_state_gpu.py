"""Checkpoint / resume on the device (pbf_checkpoint_save / _load, pbf_headless --save / --resume / --stats):
a resumed run must reproduce the uninterrupted run BIT FOR BIT — the state file carries the particle order
(the tie-break of the next stable sort), the parameters, the box and the frame counter of the wall schedule."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _run(pbf, torch, sim, d, d_iid, n, ulim, llim, first, steps, moving):
    for s in range(first, first + steps):
        if moving:
            sim.setLim(*pbf.wall_lim(ulim, llim, (2, 0, 0), (0, 0, 0), 0.05, s))
        sim.step(d[0], d[1], d[2], d[3], d_iid, n)
        d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
    torch.cuda.synchronize()


@pytest.mark.parametrize("moving", [0, 1])
def test_resume_is_bit_identical(pbf, torch, tmp_path, moving):
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n = len(iid)
    params = pbf.default_params()
    params.niter = 3   # not the default, so that the file's parameters matter

    def fresh():
        sim = pbf.Simulator(params, (4.0, 2.0, 4.0), llim, 40000)
        sim.setLim(ulim, llim)
        return sim

    def upload():
        d = [torch.from_numpy(a).cuda() for a in (pos, np.zeros_like(pos), vel, np.zeros_like(vel))]
        return d, torch.from_numpy(iid.astype(np.int64)).cuda().to(torch.int32)

    # uninterrupted: 12 steps
    sim = fresh()
    d, d_iid = upload()
    _run(pbf, torch, sim, d, d_iid, n, ulim, llim, 0, 12, moving)
    want = (d[0].cpu().numpy(), d[2].cpu().numpy(), d_iid.cpu().numpy())
    sim.close()

    # 5 steps, checkpoint, throw everything away, resume in a handle created with DEFAULT parameters
    sim = fresh()
    d, d_iid = upload()
    _run(pbf, torch, sim, d, d_iid, n, ulim, llim, 0, 5, moving)
    ck = str(tmp_path / "ck.pbfstate")
    sim.checkpoint_save(ck, d[0], d[2], d_iid, n, frame=5)
    sim.close()
    info = pbf.state_info(ck)
    assert (info.n, info.frame, info.params.niter) == (n, 5, 3)

    sim2 = pbf.Simulator(pbf.default_params(), (4.0, 2.0, 4.0), llim, 40000)
    e = [torch.zeros((40000, 3), dtype=torch.float32, device="cuda") for _ in range(4)]
    e_iid = torch.zeros(40000, dtype=torch.int32, device="cuda")
    n2, frame = sim2.checkpoint_load(ck, e[0], e[2], e_iid, 40000)
    assert (n2, frame) == (n, 5) and sim2.saveParams().niter == 3
    _run(pbf, torch, sim2, e, e_iid, n, ulim, llim, frame, 7, moving)
    got = (e[0][:n].cpu().numpy(), e[2][:n].cpu().numpy(), e_iid[:n].cpu().numpy())
    for a, b in zip(want, got):
        assert a.tobytes() == b.tobytes()

    # a file that does not fit the handle leaves the handle as it was
    big = str(tmp_path / "big.pbfstate")
    pbf.state_write(big, pos, vel, iid, params, (400.0, 200.0, 4.0), llim)
    before = (sim2.saveParams().niter, sim2.getLim())
    with pytest.raises(pbf.PbfError) as err:
        sim2.checkpoint_load(big, e[0], e[2], e_iid, 40000)
    assert err.value.code == pbf.ERR_CAPACITY
    after = (sim2.saveParams().niter, sim2.getLim())
    assert before[0] == after[0] and np.array_equal(np.asarray(before[1]), np.asarray(after[1]))
    with pytest.raises(pbf.PbfError) as err:
        sim2.checkpoint_load(ck, e[0], e[2], e_iid, 100)
    assert err.value.code == pbf.ERR_CAPACITY
    sim2.close()


def _load_dump(path):
    raw = np.fromfile(path, np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    return raw[4:4 + 12 * n].tobytes(), raw[4 + 12 * n:4 + 24 * n].tobytes(), raw[4 + 24 * n:].tobytes()


@pytest.mark.parametrize("moving", [0, 1])
def test_headless_harness_save_resume_and_stats(pbf, torch, tmp_path, moving):
    exe = os.path.join(ROOT, "pbf-cuda_b200", "pbf_headless")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "harness"], stdout=subprocess.DEVNULL)
    full, part, ck, stats = (str(tmp_path / x) for x in ("full.bin", "part.bin", "ck.pbfstate", "stats.jsonl"))
    a = subprocess.run([exe, "20", str(moving), full], capture_output=True, text=True, check=True)
    subprocess.run([exe, "8", str(moving), "--save", ck, "--save-every", "3", "--stats", stats, "--stats-every", "4"],
                   capture_output=True, text=True, check=True)
    assert pbf.state_info(ck).frame == 8
    b = subprocess.run([exe, "12", str(moving), part, "--resume", ck, "--stats", stats, "--stats-every", "4"],
                       capture_output=True, text=True, check=True)
    assert _load_dump(full) == _load_dump(part)
    ja, jb = json.loads(a.stdout), json.loads(b.stdout)
    assert ja["frame"] == jb["frame"] == 20 and ja["kinetic_energy"] == jb["kinetic_energy"]
    lines = [json.loads(l) for l in open(stats)]
    assert [l["step"] for l in lines] == [4, 8, 12, 16, 20]
    assert all(l["ms_per_step"] > 0 and l["kinetic_energy"] > 0 and np.isfinite(l["density_err_max"]) for l in lines)
    assert lines[-1]["kinetic_energy"] == ja["kinetic_energy"]
    r = subprocess.run([exe, "1", "0", "--resume", str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open" in r.stderr


def test_emitter_source_grows_the_scene_and_matches_the_harness(pbf, torch, tmp_path):
    """a ParticleSource whose update() changes the count every other step. The C++
    harness (EmitterSource + stepSource(), host/ParticleSource.h) and the Python mirror must agree bit for
    bit; the count follows the schedule; iid stays a permutation; everything stays inside the box."""
    exe = os.path.join(ROOT, "pbf-cuda_b200", "pbf_headless")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "harness"], stdout=subprocess.DEVNULL)
    steps, total = 41, 4000
    dump = str(tmp_path / "jet.bin")
    r = subprocess.run([exe, str(steps), "0", dump, "--emitter", str(total)], capture_output=True, text=True, check=True)
    info = json.loads(r.stdout)
    layer = 16 * 16
    want_n = min((steps - 1) // 2 + 1, total // layer) * layer   # layers at calls 0, 2, 4, ... while they fit
    assert info["particles"] == want_n == 15 * layer

    cap = 130000
    ulim, llim = (2.0, 2.0, 4.0), (-2.0, -2.0, 0.0)
    sim = pbf.Simulator(pbf.default_params(), (4.0, 2.0, 4.0), llim, cap)
    sim.setLim(ulim, llim)
    d = [torch.zeros((cap, 3), dtype=torch.float32, device="cuda") for _ in range(4)]
    d_iid = torch.zeros(cap, dtype=torch.int32, device="cuda")
    src = pbf.EmitterSource((-1.9, -0.4, 2.0), 16, 16, 0.05, (3.0, 0.0, 0.0), 2, total)
    n = src.initialize(d[0], d[2], d_iid, cap)
    counts = []
    for s in range(steps):
        if s > 0:
            n = src.update(d[0], d[2], d_iid, cap)
        counts.append(n)
        sim.step(d[0], d[1], d[2], d[3], d_iid, n)
        d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
    torch.cuda.synchronize()
    assert counts[0] == layer and counts[1] == layer and counts[2] == 2 * layer and counts[-1] == want_n
    g_pos, g_vel, g_iid = d[0][:n].cpu().numpy(), d[2][:n].cpu().numpy(), d_iid[:n].cpu().numpy().view(np.uint32)
    c_pos, c_vel, c_iid = _load_dump(dump)
    assert g_pos.tobytes() == c_pos and g_vel.tobytes() == c_vel and g_iid.tobytes() == c_iid
    assert np.array_equal(np.sort(g_iid), np.arange(n, dtype=np.uint32))
    assert (g_pos >= np.asarray(llim, np.float32) + 1e-3 - 1e-6).all() and (g_pos <= np.asarray(ulim, np.float32) - 1e-3 + 1e-6).all()
    assert np.isfinite(g_vel).all() and g_pos[:, 0].max() > -1.0   # the jet travelled
    sim.close()


def test_device_digest_equals_host_digest_and_ignores_the_order(pbf, torch):
    """pbf_state_digest_device (the kernel bench.py's `parity` key uses on every rank) == pbf_state_digest_host on the
    downloaded arrays; a step's cell-sorted output and the same particles shuffled have the same digest."""
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n = len(iid)
    dev = torch.device("cuda:0")
    d = [torch.from_numpy(pos).to(dev), torch.zeros((n, 3), device=dev), torch.from_numpy(vel).to(dev), torch.zeros((n, 3), device=dev)]
    d_iid = torch.from_numpy(iid.astype(np.int64)).to(dev).to(torch.int32)
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    for _ in range(3):
        sim.step(d[0], d[1], d[2], d[3], d_iid, n)
        d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
    torch.cuda.synchronize()
    on_device = pbf.state_digest(d[0], d[2], d_iid, n)
    on_host = pbf.state_digest(d[0].cpu().numpy(), d[2].cpu().numpy(), d_iid.cpu().numpy().view(np.uint32))
    assert on_device == on_host and on_device != (0, 0)
    perm = torch.randperm(n, device=dev)
    assert pbf.state_digest(d[0][perm].contiguous(), d[2][perm].contiguous(), d_iid[perm].contiguous(), n) == on_device
    assert pbf.state_digest(d[0], d[2], d_iid, 0) == (0, 0)
    # digests of disjoint parts combine
    parts = [pbf.state_digest(d[0][a:b], d[2][a:b], d_iid[a:b], b - a) for a, b in ((0, 12345), (12345, n))]
    assert pbf.combine_digests(parts) == on_device
    sim.close()


def test_kernel_family_is_a_handle_option_not_an_environment_lookup(pbf, torch, monkeypatch):
    """PBF_OPT_TEAM (include/pbf.h): the family of sweep kernels is a property of the handle — set at create from
    PBF_TEAM, changed with pbf_set_option — and both families give the same bits; changing the environment after
    create changes nothing."""
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n = len(iid)
    dev = torch.device("cuda:0")

    def run(sim):
        d = [torch.from_numpy(pos).to(dev), torch.zeros((n, 3), device=dev), torch.from_numpy(vel).to(dev), torch.zeros((n, 3), device=dev)]
        d_iid = torch.from_numpy(iid.astype(np.int64)).to(dev).to(torch.int32)
        for _ in range(2):
            sim.step(d[0], d[1], d[2], d[3], d_iid, n)
            d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        torch.cuda.synchronize()
        return pbf.state_digest(d[0], d[2], d_iid, n)

    monkeypatch.delenv("PBF_TEAM", raising=False)
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    assert sim.get_option(pbf.OPT_TEAM) == -1
    auto = run(sim)
    monkeypatch.setenv("PBF_TEAM", "0")            # after create: ignored
    assert sim.get_option(pbf.OPT_TEAM) == -1
    sim.set_option(pbf.OPT_TEAM, 0)
    thread = run(sim)
    sim.set_option(pbf.OPT_TEAM, 1)
    team = run(sim)
    assert auto == thread == team
    with pytest.raises(pbf.PbfError):
        sim.set_option(pbf.OPT_TEAM, 7)
    with pytest.raises(pbf.PbfError):
        sim.set_option(99, 0)
    sim.close()
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)   # PBF_TEAM=0 in the environment now: the default of a NEW handle
    assert sim.get_option(pbf.OPT_TEAM) == 0
    sim.close()


def test_rebinned_sweeps_give_the_same_bits(pbf, torch):
    """PBF_OPT_REBIN: which thread of a block computes which particle (consecutive slots, or the block's particles
    re-dealt in the order of their current home cell) must not change a bit — on a scene large enough for the
    thread-per-particle kernels, with a ragged last block, through the neighbour list and without it."""
    dev = torch.device("cuda:0")
    n3 = (48, 40, 37)                                   # 71 040 particles: 555 blocks of 128
    n = n3[0] * n3[1] * n3[2]
    ulim, llim = (4.0, 3.0, 4.0), (0.0, 0.0, 0.0)

    def run(rebin, steps=6, staged=0):
        pos = torch.empty((n, 3), dtype=torch.float32, device=dev)
        vel = torch.empty_like(pos)
        iid = torch.empty(n, dtype=torch.int32, device=dev)
        pbf.scene_block_device((0.2, 0.2, 0.2), n3, pos, vel, iid)
        d = [pos, torch.zeros_like(pos), vel, torch.zeros_like(vel)]
        sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
        sim.set_option(pbf.OPT_TEAM, 0)
        sim.set_option(pbf.OPT_REBIN, rebin)
        sim.set_option(pbf.OPT_STAGED, staged)
        assert sim.get_option(pbf.OPT_REBIN) == rebin and sim.get_option(pbf.OPT_STAGED) == staged
        for _ in range(steps):
            sim.step(d[0], d[1], d[2], d[3], iid, n)
            d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        torch.cuda.synchronize()
        out = pbf.state_digest(d[0], d[2], iid, n)
        sim.close()
        return out

    plain = run(0)
    assert run(1) == plain
    # PBF_OPT_STAGED: the first iteration's candidates staged in shared memory by TMA bulk copies — the same bits
    assert run(0, staged=1) == plain and run(1, staged=1) == plain
    os.environ["PBF_NO_PAIR_REUSE"] = "1"               # (read at create): the full-gather delta-p kernels
    try:
        assert run(1) == plain and run(0) == plain and run(0, staged=1) == plain
    finally:
        del os.environ["PBF_NO_PAIR_REUSE"]


def test_paired_sweeps_give_the_same_bits(pbf, torch):
    """PBF_OPT_PAIRED: two consecutive slots per thread, one walk over the union of their candidate runs (solver.cu
    gather2). Each particle is still accumulated by one thread over exactly its candidates in ascending slot order,
    so no bit may change — odd particle count (a thread with one particle, a ragged last block whose list columns are
    not 0..m-1), enough steps for home cells to drift apart, through the neighbour list and without it."""
    dev = torch.device("cuda:0")
    n3 = (47, 41, 37)                                   # 71 299 particles: odd, 557 blocks of 128, the last one ragged
    n = n3[0] * n3[1] * n3[2]
    ulim, llim = (4.0, 3.0, 4.0), (0.0, 0.0, 0.0)

    def run(paired, steps=10):
        pos = torch.empty((n, 3), dtype=torch.float32, device=dev)
        vel = torch.empty_like(pos)
        iid = torch.empty(n, dtype=torch.int32, device=dev)
        pbf.scene_block_device((0.2, 0.2, 0.2), n3, pos, vel, iid)
        d = [pos, torch.zeros_like(pos), vel, torch.zeros_like(vel)]
        sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
        sim.set_option(pbf.OPT_TEAM, 0)
        sim.set_option(pbf.OPT_PAIRED, paired)
        assert sim.get_option(pbf.OPT_PAIRED) == paired
        for _ in range(steps):
            sim.step(d[0], d[1], d[2], d[3], iid, n)
            d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        torch.cuda.synchronize()
        out = pbf.state_digest(d[0], d[2], iid, n)
        sim.close()
        return out

    plain = run(0)
    assert run(1) == plain
    os.environ["PBF_NO_PAIR_REUSE"] = "1"               # (read at create): the full-gather delta-p kernels
    try:
        assert run(1) == plain
    finally:
        del os.environ["PBF_NO_PAIR_REUSE"]


def test_morton_keys_first_step_exact_then_within_tolerance(pbf, torch):
    """PBF_OPT_MORTON (the A/B of DESIGN.md 3.1): bit-interleaved cell keys, 27 one-cell runs per particle. From a given
    state the FIRST step must give every particle the reference's bits (same within-cell order — the input order —
    and the same dx, dy, dz visiting order; only the array order differs: compared through the order-independent
    digest). From the second step on the within-cell tie order is a different one, so trajectories agree within the
    north-star tolerance (1e-5 relative on positions per step; here 5 steps), not bit for bit."""
    dev = torch.device("cuda:0")
    n3 = (47, 41, 37)
    n = n3[0] * n3[1] * n3[2]
    ulim, llim = (4.0, 3.0, 4.0), (0.0, 0.0, 0.0)

    def run(morton, steps):
        pos = torch.empty((n, 3), dtype=torch.float32, device=dev)
        vel = torch.empty_like(pos)
        iid = torch.empty(n, dtype=torch.int32, device=dev)
        pbf.scene_block_device((0.2, 0.2, 0.2), n3, pos, vel, iid)
        d = [pos, torch.zeros_like(pos), vel, torch.zeros_like(vel)]
        sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
        sim.set_option(pbf.OPT_MORTON, morton)
        assert sim.get_option(pbf.OPT_MORTON) == morton
        for _ in range(steps):
            sim.step(d[0], d[1], d[2], d[3], iid, n)
            d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        torch.cuda.synchronize()
        dig = pbf.state_digest(d[0], d[2], iid, n)
        order = torch.argsort(iid.to(torch.int64))
        out = (dig, d[0][order].cpu().numpy(), d[2][order].cpu().numpy(), iid.cpu().numpy().copy())
        sim.close()
        return out

    lin1, mor1 = run(0, 1), run(1, 1)
    assert mor1[0] == lin1[0], "the first Morton step must reproduce the reference's bits per particle"
    assert not np.array_equal(lin1[3], mor1[3]), "Morton order should differ from the x-major order"
    lin5, mor5 = run(0, 5), run(1, 5)
    scale = np.abs(lin5[1]).max()
    assert np.abs(mor5[1] - lin5[1]).max() <= 5e-5 * scale


def test_cooperative_solver_kernel_gives_the_same_bits(pbf, torch):
    """PBF_OPT_COOP: every solver pass of a small-scene pbf_step in one persistent cooperative kernel (grid-wide barriers
    instead of kernel boundaries; the team kernels' own device code) — the same bits as the separate launches, launched
    directly and replayed from a CUDA graph, with an odd iteration count too, and one launch instead of 2 niter + 1."""
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n = len(iid)
    dev = torch.device("cuda:0")

    def run(coop, graph, niter):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            d = [torch.from_numpy(pos).to(dev), torch.zeros((n, 3), device=dev), torch.from_numpy(vel).to(dev), torch.zeros((n, 3), device=dev)]
            d_iid = torch.from_numpy(iid.astype(np.int64)).to(dev).to(torch.int32)
            p = pbf.default_params()
            p.niter = niter
            sim = pbf.Simulator(p, ulim, llim, n)
            sim.set_option(pbf.OPT_COOP, coop)
            sim.set_option(pbf.OPT_GRAPH, graph)
            assert sim.get_option(pbf.OPT_COOP) == coop
            l0 = sim.launch_count()
            for _ in range(8):
                sim.step(d[0], d[1], d[2], d[3], d_iid, n, st.cuda_stream)
                d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
            st.synchronize()
            out = (pbf.state_digest(d[0], d[2], d_iid, n), sim.launch_count() - l0)
            sim.close()
        return out

    for niter in (4, 3):
        plain = run(0, 0, niter)
        for graph in (0, 1):
            got = run(1, graph, niter)
            assert got[0] == plain[0]
        assert run(1, 0, niter)[1] == plain[1] - 8 * (2 * niter + 1 - 1), "one launch instead of 2 niter + 1 per step"


@pytest.mark.parametrize("moving", [0, 1])
def test_graph_and_pdl_steps_give_the_same_bits(pbf, torch, moving):
    """PBF_OPT_GRAPH / PBF_OPT_PDL (include/pbf.h): a step replayed from a CUDA graph, with or without programmatic
    dependent launches, is the step launched kernel by kernel — over the caller's ping-pong (two graph keys), across
    a parameter change, with a box that changes every step (every step a new key: the library gives up capturing)
    and with read-backs in between (the handle's view of the step must be what the stage functions leave)."""
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n = len(iid)
    dev = torch.device("cuda:0")

    def run(graph, pdl, side_stream=True):
        st = torch.cuda.Stream() if side_stream else torch.cuda.current_stream()
        with torch.cuda.stream(st):
            d = [torch.from_numpy(pos).to(dev), torch.zeros((n, 3), device=dev), torch.from_numpy(vel).to(dev), torch.zeros((n, 3), device=dev)]
            d_iid = torch.from_numpy(iid.astype(np.int64)).to(dev).to(torch.int32)
            sim = pbf.Simulator(pbf.default_params(), (4.0, 2.0, 4.0), llim, n)
            sim.setLim(ulim, llim)
            sim.set_option(pbf.OPT_GRAPH, graph)
            sim.set_option(pbf.OPT_PDL, pdl)
            out = []
            for k in range(14):
                if moving:
                    sim.setLim(*pbf.wall_lim(ulim, llim, (2, 0, 0), (0, 0, 0), 0.05, k))
                if k == 9:
                    p = pbf.default_params()
                    p.niter = 3          # an odd iteration count flips the neighbour-list parity from step to step
                    sim.loadParams(p)
                sim.step(d[0], d[1], d[2], d[3], d_iid, n, st.cuda_stream)
                d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
                if k in (4, 11):
                    out.append(sim.read(pbf.READ_KEY).tobytes())
                    out.append(sim.read(pbf.READ_NPOS).tobytes())
                    out.append(sim.read(pbf.READ_RHO).tobytes())
            st.synchronize()
            out.append(pbf.state_digest(d[0], d[2], d_iid, n, stream=st.cuda_stream))
            launches = sim.launch_count()
            sim.close()
        return out, launches

    plain, l_plain = run(0, 0)
    for graph, pdl in ((0, 1), (1, 0), (1, 1), (-1, 1)):
        got, l = run(graph, pdl)
        assert got == plain, (graph, pdl)
        assert l == l_plain, (graph, pdl, l, l_plain)
    # the legacy default stream cannot be captured: the step is launched directly, same bits
    got, _ = run(1, 1, side_stream=False)
    assert got == plain


def test_checkpoint_load_validates_parameters_and_box_together(pbf, torch, tmp_path):
    """A file whose (h, box) PAIR fits the handle loads, even though (file h, handle box) alone would not: handle created
    at h = 0.1 on a 4x4x4 box (64 000 cells, capacity 256 000), file at h = 0.05 on a 2x2x2 box (64 000 cells) — the
    intermediate state 'h = 0.05 on the 4x4x4 box' has 512 000 cells. And a file that does not fit leaves the handle
    exactly as it was."""
    dev = torch.device("cuda:0")
    n = 64
    rng = np.random.RandomState(1)
    pos = (rng.rand(n, 3).astype(np.float32) * 1.5 + 0.2).astype(np.float32)
    vel = np.zeros((n, 3), np.float32)
    iid = np.arange(n, dtype=np.uint32)
    p_file = pbf.default_params()
    p_file.h = 0.05
    path = str(tmp_path / "pair.pbf")
    pbf.state_write(path, pos, vel, iid, p_file, (2.0, 2.0, 2.0), (0.0, 0.0, 0.0), frame=7)
    sim = pbf.Simulator(pbf.default_params(), (4.0, 4.0, 4.0), (0.0, 0.0, 0.0), 1024)
    d_pos, d_vel = torch.zeros((1024, 3), device=dev), torch.zeros((1024, 3), device=dev)
    d_iid = torch.zeros(1024, dtype=torch.int32, device=dev)
    got_n, frame = sim.checkpoint_load(path, d_pos, d_vel, d_iid, 1024)
    assert (got_n, frame) == (n, 7)
    assert sim.grid_dim() == (40, 40, 40) and abs(sim.saveParams().h - 0.05) < 1e-9
    # a file that cannot fit (h = 0.02 on the 4x4x4 box: 8e6 cells): rejected, handle untouched
    p_big = pbf.default_params()
    p_big.h = 0.02
    big = str(tmp_path / "big.pbf")
    pbf.state_write(big, pos, vel, iid, p_big, (4.0, 4.0, 4.0), (0.0, 0.0, 0.0))
    with pytest.raises(pbf.PbfError) as ei:
        sim.checkpoint_load(big, d_pos, d_vel, d_iid, 1024)
    assert ei.value.code == pbf.ERR_CAPACITY
    assert sim.grid_dim() == (40, 40, 40) and abs(sim.saveParams().h - 0.05) < 1e-9
    u, l = sim.getLim()
    assert np.allclose(u, (2, 2, 2)) and np.allclose(l, (0, 0, 0))
    # stats on more particles than the last step held is an error, not a read past the owned range
    d = [d_pos[:n].contiguous(), torch.zeros((n, 3), device=dev), d_vel[:n].contiguous(), torch.zeros((n, 3), device=dev)]
    sim.step(d[0], d[1], d[2], d[3], d_iid[:n].contiguous(), n)
    torch.cuda.synchronize()
    with pytest.raises(pbf.PbfError):
        sim.stats(d_pos, d_vel, n + 1)
    sim.close()
