"""CPU tests of the oracle (oracle/pbf_oracle.c): known answers, the reference's golden vectors,
all-pairs cross-check, invariants. The oracle is what the GPU parity tests trust, so it is pinned
here against outputs of the reference's OWN Simulator.cu (tests/golden/*.npz, produced on a B200 by
tests/golden/make_golden.py from oracle/_ref/libpbf_ref.so)."""
import hashlib
import os

import numpy as np
import pytest

import _oracle as O
import _trace as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- known answers ------------------------------------------------

def test_reference_scene_known_answers():
    pos, vel, iid, ulim, llim = O.scene_double_dam_reference()
    assert len(iid) == 32000 and np.array_equal(iid, np.arange(32000, dtype=np.uint32))
    assert np.allclose(pos[0], [-0.82531714, 0.8325885, 1.8921139], rtol=0, atol=1e-7)
    assert np.allclose(pos[-1], [0.93220377, -0.94834656, 3.8793173], rtol=0, atol=1e-7)
    assert np.allclose(pos.min(0), [-1.8574363, -1.9520124, 1.8251725], rtol=0, atol=1e-7)
    assert np.allclose(pos.max(0), [1.9520179, 1.8574691, 3.9573646], rtol=0, atol=1e-7)
    assert not vel.any()
    assert tuple(ulim) == (2.0, 2.0, 4.0) and tuple(llim) == (-2.0, -2.0, 0.0)


def test_initial_grid_statistics():
    pos, vel, iid, ulim, llim = O.scene_double_dam_reference()
    p = O.default_params()
    p.g = 0.0  # keep the initial positions: advect is then the identity (vel = 0)
    o = O.Oracle(p, ulim, llim, 32000, threads=4)
    npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
    o.bind(pos, npos, vel, nvel, iid)
    o.advect()
    o.build_grid()
    assert o.grid_dim() == (40, 40, 40)
    occ = o.grid_end() - o.grid_start()
    assert (occ > 0).sum() == 5405 and occ.max() == 18
    nc, cc = o.neighbor_count(), o.candidate_count()
    # (SURVEY.md App. C quotes 29.3 / 52 and ~172 / 242 from a 2000-particle sample; these are the
    # full-population values)
    assert nc.max() == 52 and int(nc.sum()) == 938066
    assert cc.max() == 243 and int(cc.sum()) == 5456636


def test_kernel_constants():
    p = O.default_params()
    assert p.niter == 4 and p.n_corr == 4.0 and p.pho0 == 8000.0
    assert np.float32(p.delta_q) == np.float32(0.3 * np.float64(np.float32(0.1)))
    o = O.Oracle(p, [2, 2, 4], [-2, -2, 0], 16)
    L = O.lib()
    assert np.isclose(L.orc_poly6_coef(o.h), 1.5666815e9, rtol=1e-7)
    assert np.isclose(L.orc_spiky_coef(o.h), -1.4323942e7, rtol=1e-7)
    assert np.isclose(L.orc_poly6(o.h, 0.0), 1566.6819, rtol=1e-6)
    assert L.orc_poly6(o.h, np.float32(0.1) * np.float32(0.1)) == 0.0
    assert np.isclose(L.orc_poly6(o.h, p.delta_q * p.delta_q), 1180.6058, rtol=1e-6)
    pos = np.float32([[0, 0, 1], [0.05, 0, 1]]); vel = np.zeros_like(pos); iid = np.arange(2, dtype=np.uint32)
    npos, nvel = np.zeros_like(pos), np.zeros_like(pos)
    o.bind(pos, npos, vel, nvel, iid)
    o.advect(); o.build_grid(); o.correct_density()
    assert np.isclose(o.coef_corr(), -5.14731e-16, rtol=1e-6)


def test_kernel_eps_predicate_equivalence():
    """(double)rlen < 1e-4 (reference, KERNAL_EPS is a double literal) == rlen <= float(1e-4):
    the product's kernels use the float form (pbf_math.cuh)."""
    f = np.float32(1e-4)
    around = np.array([np.nextafter(f, np.float32(0)), f, np.nextafter(f, np.float32(1))], np.float32)
    assert list(around.astype(np.float64) < 1e-4) == list(around <= f) == [True, True, False]


# ---- golden vectors of the reference's own CUDA build ------------------------------------------

def _check_against_golden(name, trace, tol_pos, tol_rel):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    checked = 0
    for k in g.files:
        if k.endswith(".sha256"):
            v = trace[k[:-7]]
            assert hashlib.sha256(np.ascontiguousarray(v).tobytes()).digest() == g[k].tobytes(), k
        elif k.endswith(".sum"):
            v = trace[k[:-4]].astype(np.float64)
            assert np.allclose([v.sum(), np.abs(v).sum()], g[k], rtol=1e-5, atol=1e-3), k
        else:
            sub = k.endswith(".sub")
            field = k[:-4] if sub else k
            v = trace[field][::T_SUB[name]] if sub else trace[field]
            ref = g[k]
            assert v.shape == ref.shape, k
            if ref.dtype.kind in "iu":
                assert np.array_equal(v, ref), k
            else:
                kind = field.split(".")[1]
                if kind.startswith(("tpos", "npos")):
                    assert np.abs(v - ref).max() <= tol_pos, (k, np.abs(v - ref).max())
                else:
                    scale = max(np.abs(ref).max(), 1e-30)
                    assert np.abs(v - ref).max() <= tol_rel * scale, (k, np.abs(v - ref).max() / scale)
        checked += 1
    assert checked > 10


T_SUB = {"dd32k": 31}


@pytest.mark.parametrize("name", ["cube2k", "floor2k", "wall2k", "ragged"])
def test_oracle_matches_reference_golden_small(name):
    """Everything upstream of powf is bit-exact (keys, order, cell table, first-iteration lambda
    and rho); device powf vs libm powf differ by a few ulp in s_corr, hence the tolerances."""
    scene = T.make_scene(name)
    tr = T.trace_oracle(scene, threads=4)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    for f in ("s0.key", "s0.iid", "s0.start", "s0.end", "s0.npos0", "s0.lam0", "s0.pho0"):
        assert np.array_equal(tr[f], g[f]), f
    _check_against_golden(name, tr, tol_pos=2e-6, tol_rel=5e-5)


def test_oracle_matches_reference_golden_32k():
    scene = T.make_scene("dd32k")
    tr = T.trace_oracle(scene, threads=8)
    g = np.load(os.path.join(GOLDEN, "dd32k.npz"))
    assert np.array_equal(tr["s0.key"], g["s0.key"]) and np.array_equal(tr["s1.key"], g["s1.key"])
    # this scene starts far from equilibrium (max speed 10 after one step): ulp-level differences
    # grow ~4x per Jacobi iteration, so the second step is compared at a looser tolerance
    _check_against_golden("dd32k", tr, tol_pos=2e-4, tol_rel=2e-3)
    for f in ("s0.lam0.sub", "s0.pho0.sub", "s0.npos0.sub"):
        assert np.array_equal(tr[f[:-4]][::31], g[f]), f
    assert np.abs(tr["s0.npos"][::31] - g["s0.npos.sub"]).max() <= 2e-6


# ---- the reference's dead DEBUG_NO_HASH_GRID idea: grid search == all pairs ---------------------

@pytest.mark.parametrize("name", ["cube2k", "ragged"])
def test_grid_search_equals_all_pairs(name):
    scene = T.make_scene(name)
    p = scene["params"]
    o = O.Oracle(p, scene["ulim"], scene["llim"], len(scene["iid"]), threads=4)
    pos, vel, iid = scene["pos"].copy(), scene["vel"].copy(), scene["iid"].copy()
    npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
    o.bind(pos, npos, vel, nvel, iid)
    o.advect(); o.build_grid()
    cnt_grid = o.neighbor_count()
    lam_ap, pho_ap, cnt_ap = o.lambda_allpairs()
    assert np.array_equal(cnt_grid, cnt_ap)          # same neighbour sets, exactly
    o.correct_density()                               # (lambda/pho computed before positions move)
    # same terms, different summation order
    assert np.allclose(o.pho(), pho_ap, rtol=2e-6)
    assert np.abs(o.lam() - lam_ap).max() <= 2e-5 * np.abs(lam_ap).max()


# ---- invariants ----------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["floor2k", "wall2k"])
def test_conservation_and_box(name):
    scene = T.make_scene(name)
    tr = T.trace_oracle(scene, threads=4)
    n = len(scene["iid"])
    for s in range(scene["steps"]):
        u, l = T.lim_for_step(scene, s)
        assert np.array_equal(np.sort(tr["s%d.iid" % s]), np.arange(n, dtype=np.uint32))
        npos = tr["s%d.npos" % s]
        assert (npos >= l + np.float32(1e-3) - 1e-6).all() and (npos <= u - np.float32(1e-3) + 1e-6).all()
        key = tr["s%d.key" % s]
        assert (np.diff(key.astype(np.int64)) >= 0).all()
        start, end = tr["s%d.start" % s], tr["s%d.end" % s]
        assert (end - start).sum() == n
        occ = np.nonzero(end > start)[0]
        assert np.array_equal(np.unique(key), occ)


def test_wall_schedule():
    """FluidSystem.cpp:104-110 with the reference's numbers: ulim.x = 2 + 2*sin(0.05*(frame-start))."""
    u, l = O.wall_lim([2, 2, 4], [-2, -2, 0], [2, 0, 0], [0, 0, 0], 0.05, 31, 0)
    t = np.float32(np.float32(0.05) * np.float32(31))
    assert u[0] == np.float32(2) + np.float32(2) * np.float32(np.sin(np.float64(t)))
    assert tuple(u[1:]) == (2.0, 4.0) and tuple(l) == (-2.0, -2.0, 0.0)


def test_threads_do_not_change_results():
    scene = T.make_scene("cube2k")
    a = T.trace_oracle(scene, threads=1)
    b = T.trace_oracle(scene, threads=4)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_empty_and_single_particle():
    p = O.default_params()
    o = O.Oracle(p, [1, 1, 1], [0, 0, 0], 4)
    z3, zu = np.zeros((0, 3), np.float32), np.zeros(0, np.uint32)
    o.step(z3.copy(), z3.copy(), z3.copy(), z3.copy(), zu.copy())      # n = 0 is a no-op
    pos = np.float32([[0.5, 0.5, 0.5]]); vel = np.zeros_like(pos); iid = np.uint32([7])
    npos, nvel = np.zeros_like(pos), np.zeros_like(pos)
    o.step(pos, npos, vel, nvel, iid)
    # a lone particle: rho = W(0), free fall for one step, then lambda pulls nothing (no neighbours)
    assert iid[0] == 7 and np.isclose(o.pho()[0], 1566.6819, rtol=1e-6)
    assert np.allclose(npos[0, :2], [0.5, 0.5]) and npos[0, 2] < 0.5
