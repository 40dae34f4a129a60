"""CPU tests of the C-ABI library: it loads, exports every symbol include/pbf.h declares, its
host-side scene generators agree bit for bit with the oracle's independent restatement, and the
compute entry points fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(pbf):
    hdr = open(os.path.join(ROOT, "include", "pbf.h")).read()
    declared = re.findall(r"PBF_API\s+[\w\s\*]+?\b(pbf_\w+)\s*\(", hdr)
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(pbf.lib(), name), "libpbf_b200.so does not export %s" % name
    assert set(declared) == set(pbf.EXPORTS)


def test_library_does_not_link_the_oracle(pbf):
    import subprocess
    out = subprocess.run(["ldd", pbf.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "pbf_ref" not in out
    syms = subprocess.run(["nm", "-D", pbf.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in syms


def test_default_params_match_reference_defaults(pbf):
    p, q = pbf.default_params(), O.default_params()
    for f, _ in p._fields_:
        assert getattr(p, f) == getattr(q, f), f
    assert (p.niter, p.pho0, p.lambda_eps, p.n_corr, p.c_XSPH) == (4, 8000.0, 1000.0, 4.0, 0.5)


def test_scene_generators_agree_with_oracle(pbf):
    a = pbf.scene_double_dam_reference()
    b = O.scene_double_dam_reference()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    src = pbf.DoubleDamSource([-1.8, 1.8, 3.8], [-0.8, 0.8, 1.8], [20, 20, 40], [0.8, -0.8, 3.8], [1.8, -1.8, 1.8], [20, 20, 40])
    pos, vel, iid = src.initialize()
    # same blocks as FluidSystem.cpp:55-61 up to how the corner coordinates were rounded there
    assert src.update() == 32000 and len(iid) == 32000 and np.abs(pos - a[0]).max() < 1e-5
    cube = pbf.FixedCubeSource([0.9, 0.8, 1.3], [0.3, 0.2, 0.5], [12, 12, 16])
    c = cube.initialize()
    d = O.scene_cube([0.9, 0.8, 1.3], [0.3, 0.2, 0.5], [12, 12, 16])
    for x, y in zip(c, d):
        assert np.array_equal(x, y)
    assert np.array_equal(cube.reset()[0], c[0])


def test_block_scene_host_matches_oracle(pbf):
    a = pbf.scene_block_host([0.2, 0.2, 0.2], [16, 8, 12], 0.05, seed=27, first_iid=100)
    b = O.scene_block([0.2, 0.2, 0.2], [16, 8, 12], 0.05, seed=27, first_iid=100)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    pos = a[0]
    assert pos.min() >= 0.2 and (pos.max(0) <= np.float32([1.0, 0.6, 0.8]) + 1e-6).all()
    # jitter is U[0,1)*0.2*spacing on top of the lattice
    lattice = (np.stack(np.meshgrid(np.arange(16), np.arange(8), np.arange(12), indexing="ij"), -1).reshape(-1, 3) + 0.5) * 0.05 + 0.2
    j = pos - lattice.astype(np.float32)
    assert j.min() >= -1e-6 and j.max() < 0.2 * 0.05 + 1e-6 and 0.3 < j.mean() / (0.2 * 0.05) < 0.7


def test_scene_capacity_is_checked(pbf):
    pos = np.zeros((10, 3), np.float32); iid = np.zeros(10, np.uint32)
    cnt = C.c_int64()
    u = (C.c_float * 3)(1, 1, 1); l = (C.c_float * 3)(0, 0, 0); ns = (C.c_int32 * 3)(4, 4, 4)
    rng = C.c_uint32(27)
    rc = pbf.lib().pbf_scene_cube(u, l, ns, C.byref(rng), 0, pos.ctypes.data, pos.ctypes.data, iid.ctypes.data, 10, C.byref(cnt))
    assert rc == pbf.ERR_CAPACITY


def test_wall_lim_matches_oracle(pbf):
    for frame in (0, 1, 17, 126):
        a = pbf.wall_lim([19.2, 6.8, 9.6], [0, 0, 0], [4.8, 0, 0], [0, 0, 0], 0.05, frame)
        b = O.wall_lim([19.2, 6.8, 9.6], [0, 0, 0], [4.8, 0, 0], [0, 0, 0], 0.05, frame)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_no_cpu_fallback(pbf):
    """Without a usable sm_100 device every compute entry point must fail with an error code."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pbf.PbfError) as e:
        pbf.Simulator(pbf.default_params(), (1, 1, 1), (0, 0, 0), 1000)
    assert e.value.code == pbf.ERR_CUDA
    assert pbf.lib().pbf_step(None, None, None, None, None, None, 0, None) == pbf.ERR_INVALID
    assert b"null" in pbf.lib().pbf_last_error()


def test_cpp_shim_builds_and_fails_loudly_without_gpu(pbf):
    """The C++ mirror of the reference's Simulator / ParticleSource (pbf-cuda_b200/host) compiles with
    plain g++ against include/pbf.h, and keeps the reference's print-and-exit error convention."""
    import subprocess
    import torch
    pkg = os.path.join(ROOT, "pbf-cuda_b200")
    subprocess.check_call(["make", "-C", pkg, "harness"], stdout=subprocess.DEVNULL)
    exe = os.path.join(pkg, "pbf_headless")
    assert os.path.exists(exe)
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "1"], capture_output=True, text=True)
        assert r.returncode == 1 and "PBF error at" in r.stderr and "pbf_create" in r.stderr


def test_scene_block_slice_is_a_slice_of_the_block(pbf):
    """pbf_scene_block_slice_host (what lets every rank generate only its part of a scene): lattice layers
    [a, b) carry exactly the bits of the full block's particles, and match the oracle's generator."""
    import _oracle as O
    origin, n3 = (0.2, 0.2, 0.2), (24, 6, 10)
    fp, fv, fi = pbf.scene_block_host(origin, n3)
    op, ov, oi = O.scene_block(origin, n3, 0.05, 27, 0)
    assert np.array_equal(fp, op) and np.array_equal(fi, oi)
    per = n3[1] * n3[2]
    for a, b in ((0, 24), (5, 17), (23, 24), (7, 7)):
        sp, sv, si = pbf.scene_block_slice_host(origin, n3, a, b)
        assert np.array_equal(sp, fp[a * per:b * per]) and np.array_equal(si, fi[a * per:b * per]) and not sv.any()
    with pytest.raises(pbf.PbfError):
        pbf.scene_block_slice_host(origin, n3, 3, 25)
