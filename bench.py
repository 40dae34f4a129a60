#!/usr/bin/env python
"""bench.py — particle-steps/s of the PBF step (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl product|reference] [--scene NAME]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pbf_step (advect, grid, 4 Jacobi iterations of lambda / delta-p, velocity update,
XSPH) over the whole scene. Default workload at N=1: BASELINE config 2, the 1 048 576-particle single
dam break (pbf-cuda_b200 SCENES["dam_1m"]), synthetic, reference default parameters.

Timed region (`value`): K steps, state resident in HBM, each step bracketed by CUDA events on the
launching stream; L2 is flushed (a 256 MB memset, not timed) between steps. `e2e`: the same metric
through pbf_step_host with pinned HOST buffers (upload pos/vel/iid, step, download npos/nvel/iid
inside the timed region). `roofline`: the dominant kernel (the lambda or delta-p pass), timed live
with CUDA events inside the library, against the measured HBM peak of MEASURED_PEAKS.json.
`cpu_baseline`: the scalar oracle (oracle/pbf_oracle.c, OpenMP over particles) on the host cores on a
bounded sample of the same workload. `--impl reference` times the reference's own Simulator.cu
(oracle/_ref/libpbf_ref.so, built headless for sm_100) on the same workload; the reference has no
CPU implementation of this path, its path IS CUDA — if that library is absent the CPU oracle port is
timed instead.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

ALG_BYTES_STEP = 488          # 296 + 48*K at K = 4, per particle-step
ALG_BYTES = {"lambda": 20, "delta_p": 28}   # per particle per pass: R 12, W 4+4 / R 12+4, W 12
FALLBACK_HBM_GBS = 6650.0     # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def scene_state(pbf, torch, name, dev, sc=None):
    sc = sc or pbf.SCENES[name]
    if "blocks" in sc:
        n = sum(int(np.prod(b[1])) for b in sc["blocks"])
        pos = torch.empty((n, 3), dtype=torch.float32, device=dev)
        vel = torch.empty_like(pos)
        iid = torch.empty(n, dtype=torch.int32, device=dev)
        off = 0
        for origin, n3 in sc["blocks"]:
            m = int(np.prod(n3))
            pbf.scene_block_device(origin, n3, pos[off:], vel[off:], iid[off:], first_iid=off)
            off += m
    else:
        p, v, i, _, _ = pbf.scene_double_dam_reference()
        n = len(i)
        pos, vel = torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)
        iid = torch.from_numpy(i.astype(np.int64)).to(dev).to(torch.int32)
    torch.cuda.synchronize()
    return sc, n, pos, vel, iid


def workload_name(name, sc, n):
    d = [int(np.ceil(np.float32(np.float32(u) - np.float32(l)) / np.float32(0.1))) for u, l in zip(sc["ulim"], sc["llim"])]
    return "%s: %d particles, niter 4, box %dx%dx%d cells, reference default parameters" % (name, n, d[0], d[1], d[2])


def lim_at(pbf, sc, frame):
    if "wall" not in sc:
        return None
    w = sc["wall"]
    return pbf.wall_lim(sc["ulim"], sc["llim"], w["a_ulim"], w["a_llim"], w["w"], frame)


def run_product(args, rank, world, dist):
    import torch
    pbf = importlib.import_module("pbf-cuda_b200")   # raises if libpbf_b200.so is missing: no fallback
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sc, n, pos, vel, iid = scene_state(pbf, torch, args.scene, dev)
    npos, nvel = torch.zeros_like(pos), torch.zeros_like(vel)
    params = pbf.default_params()
    sim = pbf.Simulator(params, sc.get("ulim_max", sc["ulim"]), sc["llim"], n, device=local)
    sim.setLim(sc["ulim"], sc["llim"])
    stream = torch.cuda.current_stream().cuda_stream
    bufs = [pos, npos, vel, nvel]
    frame = [0]

    def one_step():
        lim = lim_at(pbf, sc, frame[0])
        if lim is not None:
            sim.setLim(*lim)
        sim.step(bufs[0], bufs[1], bufs[2], bufs[3], iid, n, stream)
        bufs[0], bufs[1] = bufs[1], bufs[0]
        bufs[2], bufs[3] = bufs[3], bufs[2]
        frame[0] += 1

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = sim.launch_count()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        one_step()
        ev[k][1].record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches = sim.launch_count() - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0])
    value = world * n * args.steps / (total_ms * 1e-3)

    # back-to-back (no flush) for information, and the per-kernel device times for the roofline
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1) / args.steps
    sim.enable_stage_timing(True)
    kacc, sacc, reps = {}, {}, 5
    for _ in range(reps):
        flush.zero_()
        one_step()
        for k, v in sim.kernel_ms().items():
            kacc[k] = kacc.get(k, 0.0) + v / reps
        for k, v in sim.stage_ms().items():
            sacc[k] = sacc.get(k, 0.0) + v / reps
    sim.enable_stage_timing(False)
    stats = sim.stats(bufs[0], bufs[2], n)

    if rank != 0:
        return None
    peak, peak_src = hbm_peak()
    dom = "lambda" if kacc["lambda"] >= kacc["delta_p"] else "delta_p"
    achieved = ALG_BYTES[dom] * n / (kacc[dom] * 1e-3) / 1e9
    step_gbs = ALG_BYTES_STEP * (value / world) / 1e9
    roofline = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": TRAFFIC.get(dom),
                "peak_source": peak_src, "kernel_ms": round(kacc[dom], 4),
                "algorithmic_bytes_per_particle": ALG_BYTES[dom],
                "whole_step": {"algorithmic_bytes_per_particle_step": ALG_BYTES_STEP, "achieved": round(step_gbs, 2),
                               "frac": round(step_gbs / peak, 5)},
                "note": "the lambda / XSPH sweeps are bound by the L1 wavefront rate and instruction issue, not by HBM "
                        "(SURVEY.md App. D, DESIGN.md 5): ncu at step 100 shows 67% issue-slot utilisation, 74-89% of the L1 "
                        "data-pipe wavefront rate and 12% DRAM throughput; the delta-p pass replays the lambda pass's neighbour "
                        "list and evaluates the exact powf: 76% issue, 27% DRAM (see delta_p below)",
                "delta_p": {"kernel_ms": round(kacc["delta_p"], 4), "traffic": TRAFFIC["delta_p"],
                            "dram_GBps_from_traffic": round(TRAFFIC["delta_p"] / (kacc["delta_p"] * 1e-3) / 1e9, 1)}}

    # ---- end to end through the host-buffer entry point --------------------------------------------
    # the SAME window of the SAME scene as `value`: a fresh initial state, `warmup` untimed steps, `steps`
    # timed steps, every one of them uploading its inputs from pinned host memory and downloading its result
    sc2, n2, pos0, vel0, iid0 = scene_state(pbf, torch, args.scene, dev)
    h = [torch.empty((n, 3), dtype=torch.float32).pin_memory() for _ in range(4)]
    h_iid = torch.empty(n, dtype=torch.int32).pin_memory()
    h[0].copy_(pos0); h[2].copy_(vel0); h_iid.copy_(iid0)
    del pos0, vel0, iid0
    hn = [t.numpy() for t in h]
    hi = h_iid.numpy().view(np.uint32)
    sim.setLim(sc["ulim"], sc["llim"])
    e2e_frame = [0]

    def host_step():
        lim = lim_at(pbf, sc, e2e_frame[0])
        if lim is not None:
            sim.setLim(*lim)
        sim.step_host(hn[0], hn[1], hn[2], hn[3], hi)
        hn[0], hn[1], hn[2], hn[3] = hn[1], hn[0], hn[3], hn[2]
        e2e_frame[0] += 1

    for _ in range(args.warmup):
        host_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": round(n * args.steps / e2e_s, 1), "unit": "particle-steps/s", "steps": args.steps,
           "h2d_bytes_per_step": 28 * n, "d2h_bytes_per_step": 28 * n,
           "api": "pbf_step_host (pinned host buffers; upload pos/vel/iid, step, download npos/nvel/iid), "
                  "same scene and step window as `value`, wall clock around the synchronous calls"}

    out = {"metric": "particle-steps/s", "value": round(value, 1), "unit": "particle-steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 5),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(args.scene, sc, n), "particles_per_gpu": n,
                      "parallelism": "single GPU" if world == 1 else "replicas x%d (one independent scene per GPU)" % world,
                      "l2": "flushed between timed steps (256 MB memset outside the event pairs)",
                      "ms_per_step_back_to_back": round(b2b_ms, 5), "exact_pow": True,
                      "ms_per_step_series": [round(a.elapsed_time(b), 2) for a, b in ev],
                      "stage_ms": {k: round(v, 4) for k, v in sacc.items()},
                      "kernel_ms": {k: round(v, 4) for k, v in kacc.items()},
                      "stats_after_run": {k: round(v, 6) for k, v in stats.items()}},
           "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, n, sc)
    return out


# ---- N > 1: one scene across the ranks (x-slab decomposition, pbf-cuda_b200/slab.py) ----------------

def slab_scene(pbf, name, world, scaling):
    """Block list and box of the multi-GPU workload. weak: the scene's single block and its box are
    repeated `world` times along x (per-GPU work fixed: SURVEY.md 8d config 5 'weak'); strong: the named
    scene as it is, cut into `world` slabs."""
    sc = dict(pbf.SCENES[name])
    if "blocks" not in sc:
        raise SystemExit("scene %s has no block description; multi-GPU runs need a block scene" % name)
    if scaling == "weak":
        if len(sc["blocks"]) != 1 or "wall" in sc:
            raise SystemExit("weak scaling is defined for single-block scenes without a moving wall")
        (origin, n3), = sc["blocks"]
        sc["blocks"] = [(origin, (n3[0] * world, n3[1], n3[2]))]
        sc["ulim"] = (sc["ulim"][0] * world, sc["ulim"][1], sc["ulim"][2])
    return sc


def slab_generate(pbf, slab, torch, dev, sc, planes, layer_ranges, keep=None, chunk_layers=64):
    """Generates lattice layers [a, b) of every block (layer_ranges[k] = (a, b)) on the device, in global
    input order. keep = (x0, x1): only particles whose cell plane is in [x0, x1) are returned; keep = None:
    only the per-plane histogram is returned."""
    h, llx = 0.1, float(sc["llim"][0])
    hist = torch.zeros(planes, dtype=torch.int64, device=dev)
    parts, first_iid = [], 0
    for (origin, n3), (a, b) in zip(sc["blocks"], layer_ranges):
        per = int(n3[1]) * int(n3[2])
        for ia in range(a, b, chunk_layers):
            ib = min(b, ia + chunk_layers)
            m = (ib - ia) * per
            pos = torch.empty((m, 3), dtype=torch.float32, device=dev)
            vel = torch.empty_like(pos)
            iid = torch.empty(m, dtype=torch.int32, device=dev)
            pbf.scene_block_slice_device(origin, n3, ia, ib, pos, vel, iid, first_iid=first_iid)
            pl = slab.plane_of(pos[:, 0], llx, h, planes)
            if keep is None:
                hist += torch.bincount(pl, minlength=planes)
            else:
                msk = (pl >= keep[0]) & (pl < keep[1])
                parts.append((pos[msk], vel[msk], iid[msk]))
        first_iid += int(n3[0]) * per
    if keep is None:
        return hist
    return tuple(torch.cat([q[i] for q in parts]) for i in range(3))


def run_product_slab(args, rank, world, dist):
    import torch
    pbf = importlib.import_module("pbf-cuda_b200")   # raises if libpbf_b200.so is missing: no fallback
    slab = importlib.import_module("pbf-cuda_b200.slab")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sc = slab_scene(pbf, args.scene, world, args.scaling)
    params = pbf.default_params()
    h = np.float32(0.1)
    dims = [int(np.ceil(np.float32(np.float32(u) - np.float32(l)) / h)) for u, l in zip(sc["ulim"], sc["llim"])]
    planes = dims[0]
    n_total = sum(int(np.prod(b[1])) for b in sc["blocks"])
    ghost, margin = args.ghost, args.margin
    comm = slab.TorchComm(dist, device=dev)

    # 1. per-plane histogram of the whole scene (every rank generates 1/world of the layers), the plan
    shares = [(int(n3[0]) * rank // world, int(n3[0]) * (rank + 1) // world) for _, n3 in sc["blocks"]]
    hist = slab_generate(pbf, slab, torch, dev, sc, planes, shares)
    dist.all_reduce(hist)
    hist = hist.cpu().numpy()
    bounds = slab.plan_boundaries(hist, world, 2 * (ghost + margin))
    x0, x1 = bounds[rank], bounds[rank + 1]
    # 2. capacity: the rank's share with head-room for imbalance + the planes it receives and mirrors
    per_rank = max(int(hist[bounds[r]:bounds[r + 1]].sum()) for r in range(world))
    capacity = int(1.3 * per_rank) + 2 * (2 * ghost + margin) * int(hist.max()) + 4096
    eng = slab.GpuEngine(pbf, params, sc["ulim"], sc["llim"], capacity, device_index=local,
                         stream=torch.cuda.current_stream().cuda_stream)
    sim = slab.SlabSimulator(eng, comm, params.niter, planes, ghost=ghost, margin=margin, replan_every=args.replan_every,
                             fused_halo=args.halo == "fused")
    sim.bounds = bounds
    # 3. the rank's own particles: the lattice layers that can reach its planes, filtered exactly
    delta = 0.05
    ranges = []
    for origin, n3 in sc["blocks"]:
        lo = int(np.floor((x0 * 0.1 + sc["llim"][0] - origin[0]) / delta - 0.7)) - 1
        hi = int(np.ceil((x1 * 0.1 + sc["llim"][0] - origin[0]) / delta - 0.5)) + 1
        ranges.append((min(max(lo, 0), int(n3[0])), min(max(hi, 0), int(n3[0]))))
    pos, vel, iid = slab_generate(pbf, slab, torch, dev, sc, planes, ranges, keep=(x0, x1))
    sim.load_owned(pos, vel, iid)
    del pos, vel, iid
    if sim.total_particles() != n_total:
        raise SystemExit("slab initialisation lost particles: %d of %d" % (sim.total_particles(), n_total))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    for _ in range(args.warmup):
        sim.step()
    sim.finish()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0, msgs0, bytes0 = eng.sim.launch_count(), sim.messages, sim.bytes_sent
    dist.barrier()
    torch.cuda.synchronize()
    trace = os.environ.get("PBF_BENCH_TRACE") == "1"
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        if trace:
            print("TRACE rank %d step %d begins %.4f" % (rank, k, time.time()), flush=True)
        sim.step()
        ev[k][1].record()
    torch.cuda.synchronize()
    dist.barrier()
    sim.finish()
    launches = eng.sim.launch_count() - launches0
    msgs, sent = sim.messages - msgs0, sim.bytes_sent - bytes0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, float(eng.n_own), -float(eng.n_own)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, n_max, n_min = float(t[0]), int(t[1]), int(-t[2])
    value = n_total * args.steps / (total_ms * 1e-3)

    # per-kernel device times on this rank for the roofline (same timers as the single-GPU run)
    eng.sim.enable_stage_timing(True)
    kacc, reps = {}, 3
    for _ in range(reps):
        flush.zero_()
        sim.step()
        for k, v in eng.sim.kernel_ms().items():
            kacc[k] = kacc.get(k, 0.0) + v / reps
    eng.sim.enable_stage_timing(False)
    n_own = eng.n_own
    # where a step's device time goes on every rank (phase stamps on the stream), and who is slowest
    phases = {}
    for _ in range(reps):
        flush.zero_()
        for k, v in sim.profile_step().items():
            phases[k] = phases.get(k, 0.0) + v / reps
    names = ["raw_exchange", "keys_sort_layout", "count_allgather", "lambda", "delta_p", "update_velocity", "xsph", "halo"]
    pt = torch.tensor([phases.get(k, 0.0) for k in names] + [sum(a.elapsed_time(b) for a, b in ev) / args.steps, float(eng.n_own)],
                      dtype=torch.float64, device=dev)
    allp = [torch.zeros_like(pt) for _ in range(world)]
    dist.all_gather(allp, pt)
    per_rank = [{**{k: round(float(q[i]), 3) for i, k in enumerate(names)}, "ms_per_step": round(float(q[-2]), 3),
                 "particles": int(q[-1])} for q in allp]

    # ---- end to end: every rank's state starts each step in pinned HOST memory and ends there ----------
    e2e_steps = max(3, min(args.steps, 10))
    cap = eng.capacity
    hp = [torch.empty((cap, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    hi = torch.empty(cap, dtype=torch.int32).pin_memory()
    n = eng.n_own
    hp[0][:n].copy_(eng.pos[:n]); hp[1][:n].copy_(eng.vel[:n]); hi[:n].copy_(eng.iid[:n])
    torch.cuda.synchronize()
    dist.barrier()
    up = down = 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.pos[:n].copy_(hp[0][:n], non_blocking=True); eng.vel[:n].copy_(hp[1][:n], non_blocking=True)
        eng.iid[:n].copy_(hi[:n], non_blocking=True)
        up += 28 * n
        sim.step()
        n = eng.n_own
        hp[0][:n].copy_(eng.pos[:n], non_blocking=True); hp[1][:n].copy_(eng.vel[:n], non_blocking=True)
        hi[:n].copy_(eng.iid[:n], non_blocking=True)
        down += 28 * n
        torch.cuda.synchronize()
    dist.barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s, float(up), float(down)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sim.finish()
    stats = eng.sim.stats(eng.pos, eng.vel, eng.n_own)
    torch.cuda.synchronize()
    dist.barrier()     # nobody frees arrays a neighbour may still be pushing ghost values into
    eng.close()
    if rank != 0:
        return None
    peak, peak_src = hbm_peak()
    dom = "lambda" if kacc["lambda"] >= kacc["delta_p"] else "delta_p"
    achieved = ALG_BYTES[dom] * n_own / (kacc[dom] * 1e-3) / 1e9
    step_gbs = ALG_BYTES_STEP * (value / world) / 1e9
    roofline = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": None, "peak_source": peak_src,
                "kernel_ms": round(kacc[dom], 4), "algorithmic_bytes_per_particle": ALG_BYTES[dom],
                "particles_on_this_rank": n_own,
                "whole_step": {"algorithmic_bytes_per_particle_step": ALG_BYTES_STEP, "achieved_per_gpu": round(step_gbs, 2),
                               "frac": round(step_gbs / peak, 5)},
                "note": "rank 0's kernels; the lambda / XSPH sweeps are L1-wavefront / issue bound, not HBM bound (DESIGN.md 5)"}
    e2e = {"value": round(n_total * e2e_steps / float(t[0]), 1), "unit": "particle-steps/s", "steps": e2e_steps,
           "h2d_bytes_per_step": int(float(t[1]) / e2e_steps), "d2h_bytes_per_step": int(float(t[2]) / e2e_steps),
           "api": "every rank uploads its slab's pos/vel/iid from pinned host memory, SlabSimulator.step, downloads "
                  "the result; bytes are the largest rank's"}
    return {"metric": "particle-steps/s", "value": round(value, 1), "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 5),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.scene + (" x%d along x (weak)" % world if args.scaling == "weak" else ""), sc, n_total),
                       "particles_total": n_total, "particles_per_gpu": {"min": n_min, "max": n_max},
                       "parallelism": "x-slab decomposition over %d ranks (one scene): per step one NCCL send/recv of the raw state of "
                                      "%d planes per side, then (2*niter+1) ghost refreshes %s; ghost=%d margin=%d replan_every=%d"
                                      % (world, ghost + margin,
                                         "FUSED into the pass kernels (stores into the neighbour's ghost slots over NVLink peer memory "
                                         "+ a flag handshake, no collective)" if sim.fused else "as NCCL send/recv of float4 ranges" +
                                         (" [%s]" % sim.fused_note if sim.fused_note else ""),
                                         ghost, margin, args.replan_every),
                       "slab_boundaries": [int(b) for b in sim.bounds],
                       "messages_per_step_rank0": round(msgs / args.steps, 1), "bytes_sent_per_step_rank0": int(sent / args.steps),
                       "l2": "flushed between timed steps (256 MB memset outside the event pairs)", "exact_pow": True,
                       "kernel_ms_rank0": {k: round(v, 4) for k, v in kacc.items()},
                       "phase_ms_per_rank": per_rank,
                       "ms_per_step_series_rank0": [round(a.elapsed_time(b), 2) for a, b in ev],
                       "stats_rank0": {k: round(v, 6) for k, v in stats.items()}},
            "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "clocks": clocks}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full`, dam_1m at step ~100 — the state the
# kernel timers above see (profiles/r01h_solver_ncu_summary.txt): the lambda pass writes the 8-byte neighbour
# records (344 MB) the delta-p replay reads back (420 MB with its float4 gathers)
TRAFFIC = {"lambda": 432.8e6, "delta_p": 453.8e6}


def cpu_baseline(args, n, sc, steps=None):
    """The oracle port on the host cores: bounded sample = `steps` whole steps of the same workload."""
    import _oracle as O
    threads = os.cpu_count() or 1
    pos, vel, iid = host_scene(args.scene, sc)
    npos, nvel = np.zeros_like(pos), np.zeros_like(vel)
    o = O.Oracle(O.default_params(), sc["ulim"], sc["llim"], n, threads=threads)
    if steps is None:
        steps = args.cpu_steps or max(1, min(10, int(round(4.0e6 / n)) or 1))
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(pos, npos, vel, nvel, iid)
        pos, npos, vel, nvel = npos, pos, nvel, vel
    dt = time.perf_counter() - t0
    o.close()
    return {"value": round(n * steps / dt, 1), "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": "%d whole step(s) of the same workload from its initial state (%.1f s)" % (steps, dt)}


def host_scene(name, sc):
    import _oracle as O
    if "blocks" in sc:
        parts, off = [], 0
        for origin, n3 in sc["blocks"]:
            parts.append(O.scene_block(origin, n3, 0.05, 27, off))
            off += int(np.prod(n3))
        return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
    pos, vel, iid, _, _ = O.scene_double_dam_reference()
    return pos, vel, iid


def run_reference(args, rank, world):
    """The reference arm: the reference's own Simulator.cu (headless, sm_100) on this scene; rank 0 only."""
    if rank != 0:
        return None
    import _ref
    import _oracle as O
    pbf_scenes = importlib.import_module("pbf-cuda_b200").SCENES   # scene table and generators only
    sc = pbf_scenes[args.scene]
    if not _ref.available():
        n = len(host_scene(args.scene, sc)[2])
        cb = cpu_baseline(args, n, sc, steps=max(1, min(args.steps, 3)))
        return {"impl": "reference", "metric": "particle-steps/s", "value": cb["value"], "unit": "particle-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args.scene, sc, n), "note": "oracle/_ref/libpbf_ref.so absent: CPU oracle port timed"},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    import torch
    pbf = importlib.import_module("pbf-cuda_b200")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    # N > 1: the product arm runs ONE scene over N GPUs (weak: the block repeated N times along x; strong: the
    # named scene); the single-GPU reference gets that same scene, whole, on one GPU
    sc_n = slab_scene(pbf, args.scene, world, args.scaling) if (world > 1 and not args.replicas) else None
    sc, n, pos, vel, iid = scene_state(pbf, torch, args.scene, dev, sc_n)   # same generator, same bits
    npos, nvel = torch.zeros_like(pos), torch.zeros_like(vel)
    ref = _ref.RefSimulator(O.default_params(), sc.get("ulim_max", sc["ulim"]), sc["llim"], n)
    ref.set_lim(sc["ulim"], sc["llim"])
    bufs = [pos, npos, vel, nvel]
    frame = [0]

    def one_step():
        lim = lim_at(pbf, sc, frame[0])
        if lim is not None:
            ref.set_lim(*lim)
        ref.step(bufs[0], bufs[1], bufs[2], bufs[3], iid, n)
        bufs[0], bufs[1] = bufs[1], bufs[0]
        bufs[2], bufs[3] = bufs[3], bufs[2]
        frame[0] += 1

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(0)
    sampler.start()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        one_step()
        ev[k][1].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    value = n * args.steps / (total_ms * 1e-3)
    return {"impl": "reference", "metric": "particle-steps/s", "value": round(value, 1), "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 5),
            "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.scene + (" x%d along x (weak)" % world if (sc_n is not None and args.scaling == "weak") else ""), sc, n),
                       "note": "the reference's own Simulator.cu + Simulator_kernel.cuh compiled unchanged for sm_100 "
                               "(oracle/_ref/libpbf_ref.so), Thrust sort and its cudaDeviceSynchronize fences kept, GL interop "
                               "excluded; runs on one GPU (the reference is single-GPU, rank 0 only)",
                       "l2": "flushed between timed steps"},
            "cpu_baseline": {"value": round(value, 1), "unit": "particle-steps/s", "kind": "reference", "cores": 0,
                             "sample": "all %d steps; the reference has no CPU implementation of this path — it ran its "
                                       "CUDA path on cuda:0" % args.steps},
            "e2e": {"value": round(value, 1), "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "clocks": clocks}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--scene", default="dam_1m")
    ap.add_argument("--cpu-steps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = the scene's block repeated N times along x; strong = the named scene cut in N slabs")
    ap.add_argument("--replicas", action="store_true", help="N>1: N independent copies of the scene instead of one scene in slabs")
    ap.add_argument("--ghost", type=int, default=5,
                    help="ghost planes per side; niter+1 is the guaranteed bound (MAX_DP = one cell per iteration)")
    ap.add_argument("--margin", type=int, default=6, help="planes a particle may travel between two sorts")
    ap.add_argument("--replan-every", type=int, default=20)
    ap.add_argument("--halo", default="fused", choices=["fused", "nccl"],
                    help="N>1 ghost refreshes: fused = peer-memory stores from inside the kernels; nccl = send/recv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    dist = None
    # stdout carries exactly ONE line, the JSON result: anything a library prints on fd 1 meanwhile (NCCL's
    # version banner, for one) goes to stderr
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        out = run_reference(args, rank, world)
    else:
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
            dist.init_process_group("nccl")
        if world > 1 and not args.replicas:
            out = run_product_slab(args, rank, world, dist)
        else:
            out = run_product(args, rank, world, dist)
        if dist:
            dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0 and out is not None:
        os.write(result_fd, (json.dumps(out) + "\n").encode())
    os.close(result_fd)


if __name__ == "__main__":
    main()
