"""Run under torchrun with N >= 2 ranks on N GPUs: the NCCL slab run of a 262 144-particle dam break must
equal the single-GPU pbf_step bit for bit (rank 0 runs the single-GPU reference run and compares).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/slab_nccl_check.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    pbf = importlib.import_module("pbf-cuda_b200")
    slab = importlib.import_module("pbf-cuda_b200.slab")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    replan = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    fused = (sys.argv[3] == "fused") if len(sys.argv) > 3 else False
    origin, n3 = (0.2, 0.2, 0.2), (128, 32, 64)
    ulim, llim = np.asarray((12.8, 2.0, 4.8), np.float32), np.zeros(3, np.float32)
    pos, vel, iid = pbf.scene_block_host(origin, n3)
    p = pbf.default_params()
    dims = [int(np.ceil(np.float32(ulim[a] - llim[a]) / np.float32(p.h))) for a in range(3)]
    c = [slab.plane_of(pos[:, a], llim[a], p.h, dims[a]) for a in range(3)]
    order = np.argsort((c[0] * dims[1] + c[1]) * dims[2] + c[2], kind="stable")
    gpos, gvel, giid, gplane = pos[order].copy(), vel[order].copy(), iid[order].copy(), c[0][order]

    eng = slab.GpuEngine(pbf, p, ulim, llim, len(giid), device_index=local)
    sim = slab.SlabSimulator(eng, slab.TorchComm(dist, device=dev), p.niter, dims[0], ghost=2, margin=4, replan_every=replan,
                             fused_halo=fused)
    sim.plan_initial(np.bincount(gplane, minlength=dims[0]))
    if replan:
        sim.bounds = [0] + [sim.min_width * r for r in range(1, world)] + [dims[0]]   # lopsided on purpose
    first_bounds = list(sim.bounds)
    x0, x1 = sim.my_planes()
    mine = (gplane >= x0) & (gplane < x1)
    sim.load_owned(torch.from_numpy(gpos[mine]).to(dev), torch.from_numpy(gvel[mine]).to(dev),
                   torch.from_numpy(giid[mine].view(np.int32)).to(dev))
    for _ in range(steps):
        sim.step()
    sim.finish()
    sp, sv, si = eng.state()
    n_all = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(n_all, torch.tensor([eng.n_own], dtype=torch.int64, device=dev))
    n_all = [int(x) for x in n_all]
    if rank == 0:
        parts = [(sp.clone(), sv.clone(), si.clone())]
        for r in range(1, world):
            a = torch.empty((n_all[r], 3), dtype=torch.float32, device=dev)
            b = torch.empty_like(a)
            c_ = torch.empty(n_all[r], dtype=torch.int32, device=dev)
            dist.recv(a, r); dist.recv(b, r); dist.recv(c_, r)
            parts.append((a, b, c_))
        got = [torch.cat([q[i] for q in parts]).cpu().numpy() for i in range(3)]
        n = len(giid)
        d = [torch.from_numpy(gpos).to(dev), torch.zeros((n, 3), device=dev), torch.from_numpy(gvel).to(dev), torch.zeros((n, 3), device=dev)]
        d_iid = torch.from_numpy(giid.view(np.int32)).to(dev)
        one = pbf.Simulator(p, ulim, llim, n, device=local)
        for _ in range(steps):
            one.step(d[0], d[1], d[2], d[3], d_iid, n)
            d[0], d[1], d[2], d[3] = d[1], d[0], d[3], d[2]
        torch.cuda.synchronize()
        ok = (np.array_equal(got[2], d_iid.cpu().numpy()) and np.array_equal(got[0], d[0].cpu().numpy())
              and np.array_equal(got[1], d[2].cpu().numpy()))
        print("slab_nccl_check halo=%s world=%d steps=%d particles=%d per-rank=%s bounds %s -> %s messages/step=%.1f : %s"
              % ("fused-p2p" if fused else "nccl", world, steps, n, n_all, first_bounds, list(sim.bounds), sim.messages / steps, "BIT-EXACT" if ok else "MISMATCH"), flush=True)
        rc = 0 if ok else 1
    else:
        dist.send(sp.contiguous(), 0); dist.send(sv.contiguous(), 0); dist.send(si.contiguous(), 0)
        rc = 0
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
