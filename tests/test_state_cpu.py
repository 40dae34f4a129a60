"""State files (include/pbf.h pbf_state_*): host-side I/O of the checkpoint format — no device involved.
SURVEY.md 8(f) rank 1; the reference has no dump / resume, so the contract is the header's own."""
import os
import struct

import numpy as np
import pytest


def _state(n, seed=3):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3), dtype=np.float32) * 4 - 2
    vel = rng.standard_normal((n, 3)).astype(np.float32)
    iid = rng.permutation(n).astype(np.uint32)
    return pos, vel, iid


@pytest.mark.parametrize("n", [0, 1, 7, 32000])
def test_round_trip_is_bit_exact(pbf, tmp_path, n):
    pos, vel, iid = _state(n)
    p = pbf.default_params()
    p.niter, p.n_corr, p.k_boundaryDensity = 3, 3.0, 0.25
    f = str(tmp_path / "a.pbfstate")
    pbf.state_write(f, pos, vel, iid, p, (4.0, 2.0, 4.0), (-2.0, -2.0, 0.0), frame=123, exact_pow=0)
    assert os.path.getsize(f) == 128 + 28 * n and not os.path.exists(f + ".tmp")
    info, a, b, c = pbf.state_read(f)
    assert (info.n, info.frame, info.exact_pow) == (n, 123, 0)
    assert (info.params.niter, info.params.n_corr, info.params.k_boundaryDensity) == (3, 3.0, 0.25)
    assert list(info.ulim) == [4.0, 2.0, 4.0] and list(info.llim) == [-2.0, -2.0, 0.0]
    assert a.tobytes() == pos.tobytes() and b.tobytes() == vel.tobytes() and c.tobytes() == iid.tobytes()
    assert pbf.state_info(f).checksum == info.checksum


def test_header_layout_is_the_documented_one(pbf, tmp_path):
    pos, vel, iid = _state(5)
    f = str(tmp_path / "h.pbfstate")
    pbf.state_write(f, pos, vel, iid, pbf.default_params(), (2, 2, 4), (-2, -2, 0), frame=9)
    raw = open(f, "rb").read()
    assert raw[:8] == b"PBFSTAT1"
    version, header_bytes, n, frame = struct.unpack_from("<IIqq", raw, 8)
    assert (version, header_bytes, n, frame) == (1, 128, 5, 9)
    niter, pho0, g, h, dt = struct.unpack_from("<iffff", raw, 32)
    assert (niter, pho0, h) == (4, 8000.0, np.float32(0.1)) and np.float32(dt) == np.float32(0.0083)
    assert np.frombuffer(raw, np.float32, 15, 128).tobytes() == pos.tobytes()
    # the checksum covers the payload: same payload, different header -> same checksum
    f2 = str(tmp_path / "h2.pbfstate")
    pbf.state_write(f2, pos, vel, iid, pbf.default_params(), (2, 2, 4), (-2, -2, 0), frame=10)
    assert pbf.state_info(f).checksum == pbf.state_info(f2).checksum
    iid2 = iid.copy()
    iid2[0] ^= 1
    pbf.state_write(f2, pos, vel, iid2, pbf.default_params(), (2, 2, 4), (-2, -2, 0), frame=10)
    assert pbf.state_info(f).checksum != pbf.state_info(f2).checksum


def test_damaged_files_are_refused(pbf, tmp_path):
    pos, vel, iid = _state(100)
    f = str(tmp_path / "d.pbfstate")
    pbf.state_write(f, pos, vel, iid, pbf.default_params(), (2, 2, 4), (-2, -2, 0))
    good = open(f, "rb").read()

    def expect(data, what, code=pbf.ERR_INVALID):
        open(f, "wb").write(data)
        with pytest.raises(pbf.PbfError) as e:
            pbf.state_read(f)
        assert e.value.code == code and what in str(e.value)

    flipped = bytearray(good)
    flipped[128 + 777] ^= 0x10
    expect(bytes(flipped), "checksum mismatch")
    expect(good[:-1], "truncated")
    expect(good + b"\0", "longer than its header")
    expect(good[:64], "shorter than a state header")
    expect(b"NOTASTAT" + good[8:], "bad magic")
    expect(good[:8] + struct.pack("<I", 2) + good[12:], "unsupported state version")
    with pytest.raises(pbf.PbfError):
        pbf.state_info(str(tmp_path / "missing.pbfstate"))
    # capacity of the caller's buffers is checked before anything is written into them
    open(f, "wb").write(good)
    import ctypes as C
    info = pbf.StateInfo()
    small = np.zeros((10, 3), np.float32)
    rc = pbf.lib().pbf_state_read(os.fsencode(f), C.byref(info), small.ctypes.data, small.ctypes.data, small.ctypes.data, 10)
    assert rc == pbf.ERR_CAPACITY and not small.any()


def test_emitter_source_schedule_on_host_buffers(pbf):
    """EmitterSource (the Python mirror of host/ParticleSource.h): one ny x nz layer every `period` calls, appended
    behind the current particles, until `total` or the buffers are full; iid continues the running count."""
    cap = 1000
    pos, vel, iid = np.zeros((cap, 3), np.float32), np.zeros((cap, 3), np.float32), np.zeros(cap, np.uint32)
    src = pbf.EmitterSource((-1.9, -0.4, 2.0), 4, 5, 0.05, (3.0, 0.0, 0.0), 3, 70)
    counts = [src.initialize(pos, vel, iid, cap)] + [src.update(pos, vel, iid, cap) for _ in range(11)]
    assert counts == [20, 20, 20, 40, 40, 40, 60, 60, 60, 60, 60, 60]          # 70 < 80: the fourth layer never fits
    assert np.array_equal(iid[:60], np.arange(60, dtype=np.uint32)) and not iid[60:].any()
    assert np.all(pos[:60, 0] == np.float32(-1.9)) and np.all(vel[:60] == np.float32([3.0, 0.0, 0.0]))
    layer = pos[:20].reshape(4, 5, 3)
    assert np.array_equal(layer[:, 0, 1], np.float32(-0.4) + np.float32(0.05) * np.arange(4, dtype=np.float32))
    assert np.array_equal(layer[0, :, 2], np.float32(2.0) + np.float32(0.05) * np.arange(5, dtype=np.float32))
    assert np.array_equal(pos[20:40], pos[:20]) and not pos[60:].any()
    # the buffers, not `total`, can be the limit
    small = pbf.EmitterSource((0, 0, 0), 4, 5, 0.05, (1, 0, 0), 1, 10 ** 6)
    assert [small.initialize(pos, vel, iid, 45)] + [small.update(pos, vel, iid, 45) for _ in range(3)] == [20, 40, 40, 40]
    assert src.reset(pos, vel, iid, cap) == 20


def test_state_digest_is_order_independent_and_matches_the_numpy_restatement(pbf):
    """pbf_state_digest_host (include/pbf.h): the same particles in any order give the same 128 bits, disjoint parts
    combine by + and ^, one flipped bit changes it — and bench.py's numpy restatement (what the reference arm
    uses, which must not load the product's library) computes the same value."""
    import bench
    rng = np.random.RandomState(5)
    n = 4097
    pos = rng.randn(n, 3).astype(np.float32)
    vel = rng.randn(n, 3).astype(np.float32)
    iid = rng.permutation(n).astype(np.uint32)
    d = pbf.state_digest(pos, vel, iid)
    assert d == bench.digest_numpy(pos, vel, iid)
    p = rng.permutation(n)
    assert pbf.state_digest(pos[p], vel[p], iid[p]) == d
    parts = [pbf.state_digest(pos[a:b], vel[a:b], iid[a:b]) for a, b in ((0, 1000), (1000, 1001), (1001, n))]
    assert pbf.combine_digests(parts) == d
    q = pos.copy()
    q.view(np.uint32)[17, 2] ^= 1
    assert pbf.state_digest(q, vel, iid) != d
    w = iid.copy()
    w[[3, 4]] = w[[4, 3]]                      # two particles swap their ids: a different state
    assert pbf.state_digest(pos, vel, w) != d
    assert pbf.state_digest(pos[:0], vel[:0], iid[:0]) == (0, 0) == bench.digest_numpy(pos[:0], vel[:0], iid[:0])


def test_scene_table_loads_without_the_library():
    """bench.py's reference arm reads scenes.py by path, so that the arm that times the reference never maps
    libpbf_b200.so; the table it sees is the package's."""
    import importlib
    import bench
    S = bench.load_scenes()
    pkg = importlib.import_module("pbf-cuda_b200")
    assert S.SCENES == pkg.SCENES
    assert S.scene_particles(S.SCENES["double_dam_16m"]) == 16777216 and S.scene_particles(S.SCENES["double_dam_32k"]) == 32000
    assert S.scene_dims(S.SCENES["dam_64m"]) == [768, 260, 96]
    u, l = S.wall_lim((19.2, 6.8, 9.6), (0, 0, 0), (4.8, 0, 0), (0, 0, 0), 0.05, 31)
    u2, l2 = pkg.wall_lim((19.2, 6.8, 9.6), (0, 0, 0), (4.8, 0, 0), (0, 0, 0), 0.05, 31)
    assert np.array_equal(u, u2) and np.array_equal(l, l2)
