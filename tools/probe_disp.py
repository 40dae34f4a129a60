"""Scratch probe: how far do particles move from their post-advect position during the Jacobi iterations?
Decides whether a skin-radius neighbour list built in iteration 0 can be replayed by the later passes."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pbf = importlib.import_module("pbf-cuda_b200")
sys.argv = sys.argv + [""] * 3
scene = sys.argv[1] or "dam_1m"
probe_steps = [int(v) for v in (sys.argv[2] or "5,30,60,100").split(",")]
sc = pbf.SCENES[scene]
dev = torch.device("cuda:0")
if "blocks" in sc:
    n = sum(int(np.prod(b[1])) for b in sc["blocks"])
    pos = torch.empty((n, 3), device=dev); vel = torch.empty_like(pos); iid = torch.empty(n, dtype=torch.int32, device=dev)
    off = 0
    for origin, n3 in sc["blocks"]:
        off += pbf.scene_block_device(origin, n3, pos[off:], vel[off:], iid[off:], first_iid=off)
else:
    p, v, i, _, _ = pbf.scene_double_dam_reference(); n = len(i)
    pos = torch.from_numpy(p).to(dev); vel = torch.from_numpy(v).to(dev); iid = torch.from_numpy(i.astype(np.int64)).to(dev).to(torch.int32)
npos = torch.zeros_like(pos); nvel = torch.zeros_like(vel)
params = pbf.default_params()
sim = pbf.Simulator(params, sc["ulim"], sc["llim"], n)
dims = sim.grid_dim()
h = 0.1
b = [pos, npos, vel, nvel]
for step in range(max(probe_steps) + 1):
    if step in probe_steps:
        sim.begin(b[0], b[1], b[2], b[3], iid, n); sim.advect(); sim.buildGridHash()
        x0 = sim.read(pbf.READ_NPOS)
        key = sim.read(pbf.READ_KEY).astype(np.int64)
        cx, cy, cz = key // (dims[1] * dims[2]), (key // dims[2]) % dims[1], key % dims[2]
        out = []
        for it in range(params.niter):
            sim.correctDensity()
            x = sim.read(pbf.READ_NPOS)
            d = np.sqrt(((x - x0) ** 2).sum(1)) / h
            row = "it%d: mean %.4f p99 %.4f max %.3f |" % (it, d.mean(), np.quantile(d, 0.99), d.max())
            for thr in (0.025, 0.05, 0.1):
                bad = d > thr
                dirty = np.zeros(dims, bool)
                bx, by, bz = cx[bad], cy[bad], cz[bad]
                for ax in (-1, 0, 1):
                    for ay in (-1, 0, 1):
                        for az in (-1, 0, 1):
                            dirty[np.clip(bx + ax, 0, dims[0] - 1), np.clip(by + ay, 0, dims[1] - 1), np.clip(bz + az, 0, dims[2] - 1)] = True
                # fraction of PARTICLES whose home cell is dirty (they fall back to the full gather)
                row += " thr %.3f: moved %.4f fallback %.4f |" % (thr, bad.mean(), dirty[cx, cy, cz].mean())
            # home-cell changes
            c2 = np.clip(np.trunc((x[:, 0] - sc["llim"][0]) / np.float32(h)), 0, dims[0] - 1)
            out.append(row + " home-x changed %.4f" % (c2 != cx).mean())
        sim.updateVelocity(); sim.correctVelocity(); sim.end()
        nc = sim.read(pbf.READ_NEIGHBOR_COUNT)
        print("step %d  neighbours mean %.1f max %d" % (step, nc.mean(), nc.max())); [print("   ", r) for r in out]
        sys.stdout.flush()
    else:
        sim.step(b[0], b[1], b[2], b[3], iid, n)
    b[0], b[1], b[2], b[3] = b[1], b[0], b[3], b[2]
