"""GPU parity tests: libpbf_b200.so through the C-ABI against (1) the committed golden vectors of
the reference's own CUDA build — BIT-EXACT, (2) the reference's library itself when
oracle/_ref/libpbf_ref.so travelled to the box — BIT-EXACT, (3) the CPU oracle on the same seeded
inputs — exact for keys / order / cell table / neighbour counts, float tolerance stated per field
(device powf vs libm powf is the only arithmetic difference), plus size-independent properties at
the benchmark's full size.

Tolerances (BASELINE.json north_star): keys, sorted order, neighbour counts bit-exact;
lambda / delta-p / positions within 1e-5 relative (norm-wise) of the reference's CUDA path — the
default build is in fact bit-identical to it."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import _oracle as O
import _ref
import _trace as T

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = ["cube2k", "floor2k", "wall2k", "ragged", "dd32k"]
SUB = {"dd32k": 31}


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; the product has no CPU path")
    return torch


@pytest.mark.parametrize("name", SCENES)
def test_bit_exact_vs_reference_golden(pbf, torch, name):
    """Every field of every stage of every step equals what the reference's Simulator.cu produced."""
    scene = T.make_scene(name)
    tr = T.trace_product(scene, pbf)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    checked = 0
    for k in g.files:
        if k.endswith(".sha256"):
            assert hashlib.sha256(np.ascontiguousarray(tr[k[:-7]]).tobytes()).digest() == g[k].tobytes(), k
        elif k.endswith(".sum"):
            v = tr[k[:-4]].astype(np.float64)
            assert np.array_equal([v.sum(), np.abs(v).sum()], g[k]), k
        elif k.endswith(".sub"):
            assert np.array_equal(tr[k[:-4]][::SUB[name]], g[k]), k
        else:
            assert np.array_equal(tr[k], g[k]), k
        checked += 1
    assert checked >= 30


@pytest.mark.parametrize("team", ["0", "1"])
@pytest.mark.parametrize("name", SCENES)
def test_both_kernel_families_are_bit_exact(pbf, torch, monkeypatch, name, team):
    """The sweeps exist twice: one thread per particle (solver.cu, what large scenes run) and four lanes per
    particle (solver_team.cu, what scenes below ~75 K particles run). The golden scenes are small, so left alone
    they would only ever exercise the second family: PBF_TEAM forces each in turn through the golden comparison."""
    monkeypatch.setenv("PBF_TEAM", team)
    test_bit_exact_vs_reference_golden(pbf, torch, name)


@pytest.mark.parametrize("name", SCENES)
def test_bit_exact_vs_reference_library(pbf, torch, name):
    """Same comparison against the reference's library run live on this GPU (all fields, no subsampling)."""
    if not _ref.available():
        pytest.skip("oracle/_ref/libpbf_ref.so not present on this box")
    scene = T.make_scene(name)
    a = T.trace_product(scene, pbf)
    b = T.trace_reference(scene)
    for k in b:
        assert np.array_equal(a[k], b[k]), k
    for s in range(scene["steps"]):   # the caller-visible scratch too: pos = sorted input positions
        assert np.array_equal(a["s%d.iid_out" % s], b["s%d.iid" % s])


@pytest.mark.parametrize("name", SCENES)
def test_step_equals_stage_sequence(pbf, torch, name):
    """pbf_step and the stage entry points run the same kernels: identical outputs."""
    scene = T.make_scene(name)
    a = T.trace_product(scene, pbf, use_step=True)
    b = T.trace_product(scene, pbf, use_step=False)
    for k in a:
        key = k if k in b else None
        if k.endswith(".iid"):
            assert np.array_equal(a[k], b[k.replace(".iid", ".iid_out")]), k
        elif key:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name", SCENES)
def test_vs_cpu_oracle(pbf, torch, name):
    scene = T.make_scene(name)
    a = T.trace_product(scene, pbf)
    b = T.trace_oracle(scene, threads=8)
    loose = name == "dd32k"   # far-from-equilibrium start: ulp differences grow ~4x per iteration
    for s in range(scene["steps"]):
        for f in ("key", "iid", "start", "end"):
            assert np.array_equal(a["s%d.%s" % (s, f)], b["s%d.%s" % (s, f)]), (s, f)
    # step 0: everything before the first powf is bit-exact
    for f in ("s0.npos0", "s0.lam0", "s0.pho0"):
        assert np.array_equal(a[f], b[f]), f
    for k, v in b.items():
        if v.dtype.kind != "f" or not k.startswith("s0."):
            continue
        scale = max(np.abs(v).max(), 1e-30)
        tol = 1e-4 if loose else 1e-5
        if k.split(".")[1] in ("vel", "nvel"):
            tol *= 10   # velocity = position difference / dt (dt = 0.0083)
        assert np.abs(a[k] - v).max() <= tol * scale, (k, np.abs(a[k] - v).max() / scale)


@pytest.mark.parametrize("name", ["cube2k", "ragged", "dd32k"])
def test_neighbor_counts_exact(pbf, torch, name):
    scene = T.make_scene(name)
    a = T.trace_product(scene, pbf)
    p = scene["params"]
    o = O.Oracle(p, scene["ulim"], scene["llim"], len(scene["iid"]), threads=8)
    pos, vel, iid = scene["pos"].copy(), scene["vel"].copy(), scene["iid"].copy()
    npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
    o.bind(pos, npos, vel, nvel, iid)
    o.advect(); o.build_grid()
    assert np.array_equal(a["s0.ncount"], o.neighbor_count())


def test_fast_pow_option_within_tolerance(pbf, torch):
    """exact_pow = 0 ((w*w)^2 instead of powf) stays within the north_star tolerance after one step."""
    scene = T.make_scene("cube2k")
    a = T.trace_product(scene, pbf, exact_pow=False)
    b = T.trace_product(scene, pbf, exact_pow=True)
    for k in ("s0.tpos3", "s0.npos"):
        assert np.abs(a[k] - b[k]).max() <= 1e-5 * np.abs(b[k]).max(), k
    assert np.abs(a["s0.lam3"] - b["s0.lam3"]).max() <= 1e-5 * np.abs(b["s0.lam3"]).max()
    assert not np.array_equal(a["s0.npos"], b["s0.npos"])   # it really is a different evaluation


def test_step_host_equals_device_step(pbf, torch):
    scene = T.make_scene("ragged")
    p = scene["params"]
    gp = pbf.GUIParams(); C.memmove(C.byref(gp), C.byref(p), C.sizeof(gp))
    n = len(scene["iid"])
    sim = pbf.Simulator(gp, scene["ulim"], scene["llim"], n)
    pos, vel, iid = scene["pos"].copy(), scene["vel"].copy(), scene["iid"].copy()
    npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
    sim.step_host(pos, npos, vel, nvel, iid)
    assert np.array_equal(pos, scene["pos"]) and np.array_equal(vel, scene["vel"])   # inputs untouched
    tr = T.trace_product(dict(scene, steps=1), pbf, use_step=True)
    assert np.array_equal(npos, tr["s0.npos"]) and np.array_equal(nvel, tr["s0.nvel"])
    assert np.array_equal(iid, tr["s0.iid"])
    sim.close()


def test_empty_and_tiny_inputs(pbf, torch):
    sim = pbf.Simulator(pbf.default_params(), (1, 1, 1), (0, 0, 0), 64)
    z = torch.zeros((0, 3), device="cuda"); zi = torch.zeros(0, dtype=torch.int32, device="cuda")
    sim.step(z, z.clone(), z.clone(), z.clone(), zi, 0)          # n = 0: no launches, no error
    pos = torch.tensor([[0.5, 0.5, 0.5]], device="cuda"); vel = torch.zeros_like(pos)
    npos, nvel = torch.zeros_like(pos), torch.zeros_like(pos)
    iid = torch.tensor([7], dtype=torch.int32, device="cuda")
    sim.step(pos, npos, vel, nvel, iid, 1)
    torch.cuda.synchronize()
    assert int(iid[0]) == 7 and np.isclose(sim.read(pbf.READ_RHO)[0], 1566.6819, rtol=1e-6)
    q = npos.cpu().numpy()[0]
    assert np.allclose(q[:2], [0.5, 0.5]) and q[2] < 0.5
    with pytest.raises(pbf.PbfError) as e:                        # more particles than the handle holds
        sim.step(pos, npos, vel, nvel, iid, 65)
    assert e.value.code == pbf.ERR_CAPACITY
    with pytest.raises(pbf.PbfError) as e:                        # box larger than the cell table
        sim.setLim((50, 50, 50), (0, 0, 0))
    assert e.value.code == pbf.ERR_CAPACITY
    with pytest.raises(pbf.PbfError) as e:                        # stage out of order
        sim.correctDensity()
    assert e.value.code == pbf.ERR_STATE
    sim.close()


def test_all_particles_in_one_cell(pbf, torch):
    """Collision edge case: 300 particles inside a single cell (list flushes, long runs) == oracle."""
    rng = np.random.RandomState(3)
    n = 300
    pos = (np.float32([0.55, 0.55, 0.55]) + rng.rand(n, 3).astype(np.float32) * np.float32(0.04)).astype(np.float32)
    scene = dict(name="onecell", params=O.default_params(), ulim=np.float32([1, 1, 1]), llim=np.float32([0, 0, 0]),
                 pos=pos, vel=np.zeros_like(pos), iid=np.arange(n, dtype=np.uint32), steps=1, wall=None)
    a = T.trace_product(scene, pbf)
    b = T.trace_oracle(scene, threads=4)
    assert np.array_equal(a["s0.key"], b["s0.key"]) and np.array_equal(a["s0.iid"], b["s0.iid"])
    assert np.array_equal(a["s0.ncount"], np.full(n, n, np.uint32))
    assert np.array_equal(a["s0.pho0"], b["s0.pho0"]) and np.array_equal(a["s0.lam0"], b["s0.lam0"])
    assert np.abs(a["s0.npos"] - b["s0.npos"]).max() <= 1e-5


@pytest.mark.parametrize("team", ["0", "1"])
@pytest.mark.parametrize("n,odd", [(1200, 0), (1999, 1)])
def test_dense_cluster_flushes_and_overflows(pbf, torch, monkeypatch, n, odd, team):
    """Collision edge case for the hit-word list and the neighbour-list hand-over: n particles crammed into a
    2x2x2 block of cells. Every particle sees 4 non-empty runs of ~n/4 slots (tens of hit words: the 15-word
    list drains several times per particle), hundreds of neighbours (the 96-entry pair list overflows, the
    delta-p pass takes its full-gather kernel), run starts at every alignment (n odd: the aligned 4-slot
    groups read before the start and past the end of runs and of the arrays). Two steps, every stage,
    bit for bit against the reference's own library on this GPU (the oracle where that is not present) — through
    the thread-per-particle kernels (team = 0: the 15-word list drains) and the four-lane kernels (team = 1: the
    128-slot neighbour list drains)."""
    monkeypatch.setenv("PBF_TEAM", team)
    rng = np.random.RandomState(11 + odd)
    pos = (np.float32([0.5, 0.5, 0.5]) + rng.rand(n, 3).astype(np.float32) * np.float32(0.2)).astype(np.float32)
    pos[:7] = pos[7:14]                                    # coincident particles: r2 == 0 pairs that are not self
    vel = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(0.2)
    scene = dict(name="cluster", params=O.default_params(), ulim=np.float32([1.2, 1.2, 1.2]), llim=np.float32([0, 0, 0]),
                 pos=pos, vel=vel, iid=rng.permutation(n).astype(np.uint32), steps=2, wall=None)
    a = T.trace_product(scene, pbf)
    assert a["s0.ncount"].max() > 96 and a["s0.ncount"].min() > 15
    if _ref.available():
        b = T.trace_reference(scene)
        for k in b:
            assert np.ascontiguousarray(a[k]).tobytes() == np.ascontiguousarray(b[k]).tobytes(), k
    else:
        b = T.trace_oracle(scene, threads=4)
        assert np.array_equal(a["s0.key"], b["s0.key"]) and np.array_equal(a["s0.iid"], b["s0.iid"])
        assert np.array_equal(a["s0.pho0"], b["s0.pho0"]) and np.array_equal(a["s0.lam0"], b["s0.lam0"])
    # the neighbour-list path and the full-gather path of the delta-p pass agree (PBF_NO_PAIR_REUSE is read at create)
    os.environ["PBF_NO_PAIR_REUSE"] = "1"
    try:
        c = T.trace_product(scene, pbf)
    finally:
        del os.environ["PBF_NO_PAIR_REUSE"]
    for k in a:
        assert np.ascontiguousarray(a[k]).tobytes() == np.ascontiguousarray(c[k]).tobytes(), k


# ---- full-size properties (BASELINE config 2: 1 048 576 particles) ---------------------------------

@pytest.fixture(scope="module")
def big(pbf, torch):
    sc = pbf.SCENES["dam_1m"]
    origin, n3 = sc["blocks"][0]
    n = int(np.prod(n3))
    dev = torch.device("cuda:0")
    pos = torch.empty((n, 3), device=dev); vel = torch.empty_like(pos)
    iid = torch.empty(n, dtype=torch.int32, device=dev)
    pbf.scene_block_device(origin, n3, pos, vel, iid)
    sim = pbf.Simulator(pbf.default_params(), sc["ulim"], sc["llim"], n)
    return dict(sc=sc, n=n, pos=pos, vel=vel, iid=iid, sim=sim, origin=origin, n3=n3)


def test_device_scene_matches_host_scene(pbf, torch, big):
    h_pos, h_vel, h_iid = pbf.scene_block_host(big["origin"], big["n3"])
    assert np.array_equal(big["pos"].cpu().numpy(), h_pos)
    assert np.array_equal(big["iid"].cpu().numpy().view(np.uint32), h_iid)
    assert not big["vel"].any()


def test_full_size_sort_and_grid_properties(pbf, torch, big):
    n, sim = big["n"], big["sim"]
    pos, vel, iid = big["pos"].clone(), big["vel"].clone(), big["iid"].clone()
    npos, nvel = torch.zeros_like(pos), torch.zeros_like(vel)
    pos_in = pos.cpu().numpy()
    for step in range(3):
        sim.step(pos, npos, vel, nvel, iid, n)
        torch.cuda.synchronize()
        key = sim.read(pbf.READ_KEY).astype(np.int64)
        src = sim.read(pbf.READ_SRC_INDEX)
        assert (np.diff(key) >= 0).all(), "keys not sorted"
        assert np.array_equal(np.sort(src), np.arange(n, dtype=np.uint32)), "sort lost or duplicated a particle"
        same = np.diff(key) == 0
        assert (np.diff(src.astype(np.int64))[same] > 0).all(), "sort is not stable within a cell"
        start, end = sim.read(pbf.READ_CELL_START).astype(np.int64), sim.read(pbf.READ_CELL_END).astype(np.int64)
        assert (end - start).sum() == n
        occ = np.nonzero(end > start)[0]
        assert np.array_equal(occ, np.unique(key))
        assert np.array_equal(key[start[occ]], occ) and np.array_equal(key[end[occ] - 1], occ)
        ids = iid.cpu().numpy().view(np.uint32)
        assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32)), "iid is not a permutation"
        q = npos.cpu().numpy()
        lo, hi = np.float32(big["sc"]["llim"]) + np.float32(1e-3), np.float32(big["sc"]["ulim"]) - np.float32(1e-3)
        assert (q >= lo - 1e-6).all() and (q <= hi + 1e-6).all() and np.isfinite(q).all()
        if step == 0:
            # pos now holds the step-input positions in sorted order
            assert np.array_equal(pos.cpu().numpy(), pos_in[src])
            assert np.array_equal(ids, src)   # iid == source index on the first step of this scene
        pos, npos = npos, pos
        vel, nvel = nvel, vel
    st = sim.stats(pos, vel, n)
    assert 0 <= st["density_err_mean"] < 0.2 and st["max_speed"] < 12 and np.isfinite(st["kinetic_energy"])


def test_full_size_matches_oracle_sample_and_is_deterministic(pbf, torch, big):
    """One full-size step: keys and order exact vs the CPU oracle, positions within tolerance,
    and a second run from the same state is bit-identical (no atomics in any ordered path)."""
    n, sim = big["n"], big["sim"]
    outs = []
    for rep in range(2):
        pos, vel, iid = big["pos"].clone(), big["vel"].clone(), big["iid"].clone()
        npos, nvel = torch.zeros_like(pos), torch.zeros_like(vel)
        sim.step(pos, npos, vel, nvel, iid, n)
        torch.cuda.synchronize()
        outs.append((npos.cpu().numpy(), nvel.cpu().numpy(), iid.cpu().numpy().view(np.uint32), sim.read(pbf.READ_KEY)))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    h_pos, h_vel, h_iid = pbf.scene_block_host(big["origin"], big["n3"])
    o = O.Oracle(O.default_params(), big["sc"]["ulim"], big["sc"]["llim"], n, threads=os.cpu_count() or 1)
    o_npos, o_nvel = np.zeros_like(h_pos), np.zeros_like(h_vel)
    o.step(h_pos, o_npos, h_vel, o_nvel, h_iid)
    assert np.array_equal(outs[0][3], o.grid_id())
    assert np.array_equal(outs[0][2], h_iid)
    assert np.abs(outs[0][0] - o_npos).max() <= 1e-5 * np.abs(o_npos).max()
    assert np.abs(outs[0][1] - o_nvel).max() <= 1e-4 * max(np.abs(o_nvel).max(), 1e-30)
    o.close()


def test_full_size_trajectory_is_bit_identical_to_the_reference_library(pbf, torch, big):
    """BASELINE config 2 itself: 1 048 576 particles over the benchmark's whole window, 110 steps into the collapse
    (the compressed regime in which home cells drift and neighbour lists are long) — positions, velocities and
    order after every tenth step equal the reference's own Simulator.cu run live on this GPU, bit for bit. The stable sort makes the orders equal, so
    nothing has to be matched up by iid."""
    if not _ref.available():
        pytest.skip("oracle/_ref/libpbf_ref.so not present on this box")
    n, sc = big["n"], big["sc"]
    sim = big["sim"]
    sim.setLim(sc["ulim"], sc["llim"])
    ref = _ref.RefSimulator(O.default_params(), sc["ulim"], sc["llim"], n)
    ref.set_lim(sc["ulim"], sc["llim"])
    a = [big["pos"].clone(), torch.zeros_like(big["pos"]), big["vel"].clone(), torch.zeros_like(big["vel"])]
    b = [t.clone() for t in a]
    a_iid, b_iid = big["iid"].clone(), big["iid"].clone()
    for step in range(1, 111):
        sim.step(a[0], a[1], a[2], a[3], a_iid, n)
        ref.step(b[0], b[1], b[2], b[3], b_iid, n)
        a[0], a[1], a[2], a[3] = a[1], a[0], a[3], a[2]
        b[0], b[1], b[2], b[3] = b[1], b[0], b[3], b[2]
        if step % 10 == 0:
            torch.cuda.synchronize()
            assert torch.equal(a_iid, b_iid), step
            assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32)), step   # bit patterns, NaN-safe
            assert torch.equal(a[2].view(torch.int32), b[2].view(torch.int32)), step
    ref.close()


def test_sweeping_tank_4m_is_bit_identical_to_the_reference_library(pbf, torch):
    """BASELINE config 3: 4 194 304 particles, the wall moving every step (grid dimensions change with it), XSPH on —
    20 steps, state compared with the reference's library bit for bit after steps 10 and 20."""
    if not _ref.available():
        pytest.skip("oracle/_ref/libpbf_ref.so not present on this box")
    sc = pbf.SCENES["sweep_4m"]
    (origin, n3), = sc["blocks"]
    n = int(np.prod(n3))
    pos = torch.empty((n, 3), device="cuda"); vel = torch.empty_like(pos)
    iid = torch.empty(n, dtype=torch.int32, device="cuda")
    pbf.scene_block_device(origin, n3, pos, vel, iid)
    sim = pbf.Simulator(pbf.default_params(), sc["ulim_max"], sc["llim"], n)
    ref = _ref.RefSimulator(O.default_params(), sc["ulim_max"], sc["llim"], n)
    a = [pos, torch.zeros_like(pos), vel, torch.zeros_like(vel)]
    b = [t.clone() for t in a]
    a_iid, b_iid = iid, iid.clone()
    w = sc["wall"]
    for step in range(1, 21):
        lim = pbf.wall_lim(sc["ulim"], sc["llim"], w["a_ulim"], w["a_llim"], w["w"], step - 1)
        sim.setLim(*lim)
        ref.set_lim(*lim)
        sim.step(a[0], a[1], a[2], a[3], a_iid, n)
        ref.step(b[0], b[1], b[2], b[3], b_iid, n)
        a[0], a[1], a[2], a[3] = a[1], a[0], a[3], a[2]
        b[0], b[1], b[2], b[3] = b[1], b[0], b[3], b[2]
        if step % 10 == 0:
            torch.cuda.synchronize()
            assert torch.equal(a_iid, b_iid), step
            assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32)), step
            assert torch.equal(a[2].view(torch.int32), b[2].view(torch.int32)), step
    ref.close()
    sim.close()


def test_double_dam_16m_is_bit_identical_to_the_reference_library(pbf, torch):
    """BASELINE config 4 on one GPU: 16 777 216 particles in two blocks, 14.2 M cells (24-bit keys: three sort
    passes, 64-bit offsets everywhere) — 4 steps, bit for bit against the reference's library."""
    if not _ref.available():
        pytest.skip("oracle/_ref/libpbf_ref.so not present on this box")
    sc = pbf.SCENES["double_dam_16m"]
    n = sum(int(np.prod(b[1])) for b in sc["blocks"])
    pos = torch.empty((n, 3), device="cuda"); vel = torch.empty_like(pos)
    iid = torch.empty(n, dtype=torch.int32, device="cuda")
    off = 0
    for origin, n3 in sc["blocks"]:
        pbf.scene_block_device(origin, n3, pos[off:], vel[off:], iid[off:], first_iid=off)
        off += int(np.prod(n3))
    sim = pbf.Simulator(pbf.default_params(), sc["ulim"], sc["llim"], n)
    ref = _ref.RefSimulator(O.default_params(), sc["ulim"], sc["llim"], n)
    ref.set_lim(sc["ulim"], sc["llim"])
    a = [pos, torch.zeros_like(pos), vel, torch.zeros_like(vel)]
    b = [t.clone() for t in a]
    a_iid, b_iid = iid, iid.clone()
    for step in range(4):
        sim.step(a[0], a[1], a[2], a[3], a_iid, n)
        ref.step(b[0], b[1], b[2], b[3], b_iid, n)
        a[0], a[1], a[2], a[3] = a[1], a[0], a[3], a[2]
        b[0], b[1], b[2], b[3] = b[1], b[0], b[3], b[2]
    torch.cuda.synchronize()
    assert torch.equal(a_iid, b_iid)
    assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32))
    assert torch.equal(a[2].view(torch.int32), b[2].view(torch.int32))
    ref.close()
    sim.close()


def test_trajectory_statistics_vs_oracle(pbf, torch):
    """50 steps of the reference scene: density error and kinetic energy track the CPU oracle.
    Trajectories are chaotic, so this compares statistics, not particles (north_star); the
    tolerance is ~10x the spread the oracle shows against itself under a 1-ulp input perturbation."""
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n, steps = len(iid), 50
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    d = [torch.from_numpy(a).cuda() for a in (pos, np.zeros_like(pos), vel, np.zeros_like(vel))]
    d_iid = torch.from_numpy(iid.astype(np.int64)).cuda().to(torch.int32)
    o = O.Oracle(O.default_params(), ulim, llim, n, threads=os.cpu_count() or 1)
    h = [pos.copy(), np.zeros_like(pos), vel.copy(), np.zeros_like(vel)]
    h_iid = iid.copy()
    for s in range(steps):
        sim.step(d[0], d[1], d[2], d[3], d_iid, n)
        o.step(h[0], h[1], h[2], h[3], h_iid)
        d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        h[0], h[1], h[2], h[3] = h[1], h[0], h[3], h[2]
        if s in (0, 9, 24, 49):
            gs = sim.stats(d[0], d[2], n)
            hs = O.stats(o.pho(), h[0], h[2], 8000.0)
            assert abs(gs["density_err_mean"] - hs["density_err_mean"]) <= 0.02 * hs["density_err_mean"] + 1e-4, (s, gs, hs)
            assert abs(gs["kinetic_energy"] - hs["kinetic_energy"]) <= 0.02 * hs["kinetic_energy"], (s, gs, hs)
            assert abs(gs["mean_z"] - hs["mean_z"]) <= 1e-3
    # the device statistics reduce exactly what the oracle's reduce on the same arrays
    gs = sim.stats(d[0], d[2], n)
    hs2 = O.stats(sim.read(pbf.READ_RHO), d[0].cpu().numpy(), d[2].cpu().numpy(), 8000.0)
    for k in gs:
        assert np.isclose(gs[k], hs2[k], rtol=1e-12, atol=1e-12), k
    sim.close(); o.close()


def test_timers_and_launch_count(pbf, torch, big):
    n, sim = big["n"], big["sim"]
    pos, vel, iid = big["pos"].clone(), big["vel"].clone(), big["iid"].clone()
    npos, nvel = torch.zeros_like(pos), torch.zeros_like(vel)
    sim.enable_stage_timing(True)
    before = sim.launch_count()
    sim.step(pos, npos, vel, nvel, iid, n)
    ms, kms = sim.stage_ms(), sim.kernel_ms()
    sim.enable_stage_timing(False)
    assert set(ms) == set(pbf.STAGE_NAMES) and all(v > 0 for v in ms.values())
    assert all(v > 0 for v in kms.values()) and kms["lambda"] < ms["DENSITY"]
    assert 10 <= sim.launch_count() - before <= 40


@pytest.mark.parametrize("moving", [0, 1])
def test_cpp_headless_harness_matches_python_path(pbf, torch, tmp_path, moving):
    """pbf_headless (the reference's FluidSystem loop in C++ over the shim Simulator.h / ParticleSource.h)
    and the Python binding drive the same C-ABI: 25 steps of the 32K scene, bit-identical state."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "pbf-cuda_b200", "pbf_headless")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "harness"], stdout=subprocess.DEVNULL)
    steps = 25
    dump = str(tmp_path / "state.bin")
    r = subprocess.run([exe, str(steps), str(moving), dump], capture_output=True, text=True, check=True)
    info = json.loads(r.stdout)
    raw = np.fromfile(dump, np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    assert n == 32000 == info["particles"]
    c_pos = raw[4:4 + 12 * n].view(np.float32).reshape(n, 3)
    c_vel = raw[4 + 12 * n:4 + 24 * n].view(np.float32).reshape(n, 3)
    c_iid = raw[4 + 24 * n:].view(np.uint32)

    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    sim = pbf.Simulator(pbf.default_params(), (4.0, 2.0, 4.0), llim, 130000)
    sim.setLim(ulim, llim)
    d = [torch.from_numpy(a).cuda() for a in (pos, np.zeros_like(pos), vel, np.zeros_like(vel))]
    d_iid = torch.from_numpy(iid.astype(np.int64)).cuda().to(torch.int32)
    for s in range(steps):
        if moving:
            sim.setLim(*pbf.wall_lim(ulim, llim, (2, 0, 0), (0, 0, 0), 0.05, s))
        sim.step(d[0], d[1], d[2], d[3], d_iid, n)
        d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
    torch.cuda.synchronize()
    assert np.array_equal(c_iid, d_iid.cpu().numpy().view(np.uint32))
    assert np.array_equal(c_pos, d[0].cpu().numpy()) and np.array_equal(c_vel, d[2].cpu().numpy())
    st = sim.stats(d[0], d[2], n)
    assert np.isclose(info["kinetic_energy"], st["kinetic_energy"], rtol=1e-7)
    sim.close()


def test_1000_steps_of_the_reference_scene(pbf, torch):
    """BASELINE config 1 / north_star: 1000 steps of the 32 000-particle double dam break. Trajectories are
    chaotic, so the stated criterion is statistical (density error, kinetic energy within tolerance of the
    reference's own CUDA run); because every step is bit-identical the whole trajectory is: final positions,
    velocities and order EQUAL the reference library's, and the statistics therefore agree to the last bit.
    Without the reference library on the box, the statistics are checked for sanity only."""
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n, steps = len(iid), 1000
    dev = torch.device("cuda:0")

    def run(make_step):
        d = [torch.from_numpy(pos).to(dev), torch.zeros((n, 3), device=dev), torch.from_numpy(vel).to(dev), torch.zeros((n, 3), device=dev)]
        d_iid = torch.from_numpy(iid.view(np.int32).copy()).to(dev)
        step = make_step(d_iid)
        for _ in range(steps):
            step(d[0], d[1], d[2], d[3])
            d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        torch.cuda.synchronize()
        return d[0], d[2], d_iid

    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    p_pos, p_vel, p_iid = run(lambda d_iid: (lambda a, b, c, d: sim.step(a, b, c, d, d_iid, n)))
    st = sim.stats(p_pos, p_vel, n)
    q = p_pos.cpu().numpy()
    assert np.isfinite(q).all() and (q >= llim + 1e-3 - 1e-6).all() and (q <= ulim - 1e-3 + 1e-6).all()
    assert np.array_equal(np.sort(p_iid.cpu().numpy().view(np.uint32)), np.arange(n, dtype=np.uint32))
    assert 0.0 < st["density_err_mean"] < 0.5 and st["mean_z"] < 1.0 and st["max_speed"] < 30.0   # settled in the tank
    if _ref.available():
        ref = _ref.RefSimulator(O.default_params(), ulim, llim, n)
        r_pos, r_vel, r_iid = run(lambda d_iid: (lambda a, b, c, d: ref.step(a, b, c, d, d_iid, n)))
        assert torch.equal(p_iid, r_iid) and torch.equal(p_pos, r_pos) and torch.equal(p_vel, r_vel)
        rho = sim.read(pbf.READ_RHO)
        o = O.stats(rho, p_pos.cpu().numpy(), p_vel.cpu().numpy(), pbf.default_params().pho0)
        for k in ("density_err_mean", "kinetic_energy", "mean_z"):
            assert abs(o[k] - st[k]) <= 1e-9 * max(1.0, abs(o[k])), k
    sim.close()


def test_const_division_sequence_is_verified_and_used(pbf, torch):
    """a / pho0 runs as a reciprocal sequence only inside an interval the library verified exhaustively (all
    2^32 dividends) against div.rn on this device; with the default pho0 that interval must cover everything
    but the denormal fringe — and the golden-vector tests above are what prove the bits did not change."""
    sim = pbf.Simulator(pbf.default_params(), (1, 1, 1), (0, 0, 0), 64)
    lo, hi = sim.const_div_interval()
    assert lo <= 1e-30 and hi >= 1e30, (lo, hi)
    p = pbf.default_params()
    p.pho0 = 3.0                       # another divisor: verified again, whatever the outcome it must be consistent
    sim.loadParams(p)
    lo3, hi3 = sim.const_div_interval()
    assert (lo3 > hi3) or (lo3 <= 1e-30 and hi3 >= 1e30)
    sim.close()


def test_fast_spiky_scale_is_verified_for_every_r2(pbf, torch, monkeypatch):
    """The branch-free spiky scale of the lambda pass is used only if it matched the sqrt.rn / div.rn sequence
    for EVERY float r2 in [0, h^2] on this device (about 1e9 values per h). For the default h and two others it
    must verify without a single mismatch; the golden-vector tests above prove the step's bits did not change,
    and a step with PBF_NO_FAST_SPIKY=1 must give the same bits as a step with it."""
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n = len(iid)

    def one_step(sim):
        d = [torch.from_numpy(a).cuda() for a in (pos, np.zeros_like(pos), vel, np.zeros_like(vel))]
        d_iid = torch.from_numpy(iid.astype(np.int64)).cuda().to(torch.int32)
        for _ in range(3):
            sim.step(d[0], d[1], d[2], d[3], d_iid, n)
            d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        torch.cuda.synchronize()
        return d[0].cpu().numpy().tobytes(), d[2].cpu().numpy().tobytes(), sim.read(pbf.READ_RHO).tobytes()

    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    assert sim.fast_spiky() == (1, 0)
    fast = one_step(sim)
    for h in (0.125, 0.07):
        p = pbf.default_params()
        p.h = h
        sim.loadParams(p)
        assert sim.fast_spiky() == (1, 0), (h, sim.fast_spiky())
    sim.close()
    monkeypatch.setenv("PBF_NO_FAST_SPIKY", "1")
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    assert sim.fast_spiky() == (0, 0)
    exact = one_step(sim)
    sim.close()
    assert fast == exact


def test_trimmed_pow_is_verified_for_every_w(pbf, torch, monkeypatch):
    """The special-case-free powf(w, 4) of the delta-p pass is used only if it matched the library's
    powf(w, 4.0f) for EVERY float w in [0, W(0)] on this device (about 1.15e9 values at the default h). It must
    verify without a single mismatch for the default h and two others, and steps with PBF_NO_TRIM_POW=1 (the
    library call) must give the same bits as steps with it."""
    pos, vel, iid, ulim, llim = pbf.scene_double_dam_reference()
    n = len(iid)

    def steps(sim):
        d = [torch.from_numpy(a).cuda() for a in (pos, np.zeros_like(pos), vel, np.zeros_like(vel))]
        d_iid = torch.from_numpy(iid.astype(np.int64)).cuda().to(torch.int32)
        for _ in range(12):
            sim.step(d[0], d[1], d[2], d[3], d_iid, n)
            d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        torch.cuda.synchronize()
        return d[0].cpu().numpy().tobytes(), d[2].cpu().numpy().tobytes(), sim.read(pbf.READ_RHO).tobytes()

    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    assert sim.trim_pow() == (1, 0)
    trimmed = steps(sim)
    for h in (0.125, 0.07):
        p = pbf.default_params()
        p.h = h
        sim.loadParams(p)
        assert sim.trim_pow() == (1, 0), (h, sim.trim_pow())
    sim.close()
    monkeypatch.setenv("PBF_NO_TRIM_POW", "1")
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    assert sim.trim_pow() == (0, 0)
    library = steps(sim)
    sim.close()
    assert trimmed == library
    # the same with the thread-per-particle kernels (the 32 K scene takes the four-lane kernels by default)
    monkeypatch.setenv("PBF_TEAM", "0")
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    assert steps(sim) == trimmed
    sim.close()
    monkeypatch.delenv("PBF_NO_TRIM_POW")
    sim = pbf.Simulator(pbf.default_params(), ulim, llim, n)
    assert sim.trim_pow() == (1, 0)
    assert steps(sim) == trimmed
    sim.close()
