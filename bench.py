#!/usr/bin/env python
"""bench.py — particle-steps/s of the PBF step (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl product|reference] [--scene NAME]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pbf_step (advect, grid, 4 Jacobi iterations of lambda / delta-p, velocity update,
XSPH) over the whole scene. The JSON line's headline (`value`, `ms_per_step`, `e2e`, `roofline`) is
BASELINE config 2 at N=1 — the 1 048 576-particle single dam break (scenes.py SCENES["dam_1m"]),
synthetic, reference default parameters — and that block repeated N times along x at N>1 (weak).

Every OTHER named size of BASELINE.json rides along in the same line:
  N=1   `sizes`: double_dam_32k (config 1, 200 steps), sweep_4m (config 3, moving wall, 20 steps),
        double_dam_16m (config 4 on one GPU, 10 steps) — ms/step, particle-steps/s, whole-step HBM fraction,
        the state digest after the window, and the ratio / digest comparison with the reference arm's run of
        the same window on this box when that ran first (it leaves gpurun_out/bench_reference_last.json).
  N>1   `legs`: double_dam_16m STRONG over the N ranks (config 4), and at N=8 dam_64m strong (config 5; the
        same scene as its weak form, the 8.4 M-particle block per GPU repeated 8 times along x).
  `parity`: after every multi-GPU window the order-independent digest (pbf_state_digest) of all ranks' owned
        particles is compared with the digest of ONE GPU running the same scene for the same number of steps
        (rank 0, after the timed region): "bit_exact" or "MISMATCH".

Timed region (`value`): K steps, state resident in HBM, each step bracketed by CUDA events on the
launching stream; L2 is flushed (a 256 MB memset, not timed) between steps. `e2e`: the same metric
through pbf_step_host with pinned HOST buffers (upload pos/vel/iid, step, download npos/nvel/iid
inside the timed region). `roofline`: the dominant kernel (the lambda or delta-p pass), timed live
with CUDA events inside the library, against the measured HBM peak of MEASURED_PEAKS.json.
`cpu_baseline`: the scalar oracle (oracle/pbf_oracle.c, OpenMP over particles) on the host cores on a
bounded sample of the same workload. `--impl reference` times the reference's own Simulator.cu
(oracle/_ref/libpbf_ref.so, built headless for sm_100) on the same workloads; the reference has no
CPU implementation of this path, its path IS CUDA — if that library is absent the CPU oracle port is
timed instead. The reference arm never maps libpbf_b200.so: scenes come from the oracle's host generator.
"""
import argparse
import importlib
import importlib.util
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
REF_FILE = os.path.join(ROOT, "gpurun_out", "bench_reference_last.json")

# (scene, timed steps, warm-up steps) of the `sizes` entries at N=1 — the same windows in both arms
SIZES = (("double_dam_32k", 200, 5), ("sweep_4m", 20, 5), ("double_dam_16m", 10, 3))
# multi-GPU legs behind the weak headline: (scene, scaling, timed steps, warm-up, smallest world)
LEGS = (("double_dam_16m", "strong", 10, 3, 2), ("dam_64m", "strong", 10, 3, 8))


def load_scenes():
    """scenes.py alone, by path: the reference arm must not import the package (that maps libpbf_b200.so)."""
    spec = importlib.util.spec_from_file_location("_pbf_scenes", os.path.join(ROOT, "pbf-cuda_b200", "scenes.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def digest_numpy(pos, vel, iid):
    """include/pbf.h pbf_state_digest_* restated in numpy (tests/test_state_cpu.py checks it against the library):
    what the reference arm uses, which must not load the product's library."""
    M = np.uint64
    def mix(z):
        z = z + M(0x9E3779B97F4A7C15)
        z = (z ^ (z >> M(30))) * M(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> M(27))) * M(0x94D049BB133111EB)
        return z ^ (z >> M(31))
    p = np.ascontiguousarray(pos, np.float32).view(np.uint32).reshape(-1, 3).astype(np.uint64)
    v = np.ascontiguousarray(vel, np.float32).view(np.uint32).reshape(-1, 3).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = mix(np.ascontiguousarray(iid).view(np.uint32).astype(np.uint64))
        h = mix(h ^ (p[:, 0] | (p[:, 1] << M(32))))
        h = mix(h ^ (p[:, 2] | (v[:, 0] << M(32))))
        h = mix(h ^ (v[:, 1] | (v[:, 2] << M(32))))
        s = int(np.add.reduce(h, dtype=np.uint64)) if len(h) else 0
        x = int(np.bitwise_xor.reduce(mix(h))) if len(h) else 0
    return s, x


def hexd(d):
    return "%016x%016x" % (d[0], d[1])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe). The query loop is
    started when the object is made — before the warm-up steps, nvidia-smi takes ~0.2 s to deliver its first line and
    a 20-step window lasts 40 ms — and `start()` / `stop()` bracket the timed region: the samples that arrived between
    them count (plus the first one after `stop()` if the window was shorter than the 50 ms period: the device is still
    at its load clocks then)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, launch=True):
        self.index, self.proc, self.lines, self.t0 = index, None, [], None
        if launch:
            self._launch()

    def _launch(self):
        if self.proc:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def start(self):
        self._launch()
        self.t0 = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = time.time()
        deadline = t1 + 1.0
        while not any(t >= t0 for t, _ in self.lines) and time.time() < deadline:
            time.sleep(0.01)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [ln for t, ln in self.lines if t0 <= t <= t1]
        if not inside:
            inside = [ln for t, ln in self.lines if t >= t0][:1]
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in inside:
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


def bind_to_gpu_cpus(index):
    """One process per GPU: run this rank on the CPU cores next to its GPU (NVML's affinity mask), BEFORE anything is
    allocated — pinned host buffers are then first touched on the GPU's NUMA node, and the e2e leg's uploads / downloads
    do not cross the socket interconnect. Best effort: a box without NVML or with a restricted cpuset is left alone."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return sorted(os.sched_getaffinity(0))
    except Exception:   # noqa: BLE001
        return None


def workload_name(S, name, sc, n):
    d = S.scene_dims(sc)
    return "%s: %d particles, niter 4, box %dx%dx%d cells, reference default parameters" % (name, n, d[0], d[1], d[2])


def lim_at(S, sc, frame):
    if "wall" not in sc:
        return None
    w = sc["wall"]
    return S.wall_lim(sc["ulim"], sc["llim"], w["a_ulim"], w["a_llim"], w["w"], frame)


def host_scene(sc):
    """The scene's initial state from the ORACLE's host generator (bit-identical to the product's generators,
    tests/test_capi_cpu.py, tests/test_parity_gpu.py) — used by the cpu_baseline leg and the reference arm."""
    import _oracle as O
    if "blocks" in sc:
        parts, off = [], 0
        for origin, n3 in sc["blocks"]:
            parts.append(O.scene_block(origin, n3, 0.05, 27, off))
            off += int(np.prod(n3))
        return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
    pos, vel, iid, _, _ = O.scene_double_dam_reference()
    return pos, vel, iid


def timed_window(torch, step, steps, warmup, flush, before=None):
    """`warmup` untimed steps, then `steps` steps each bracketed by CUDA events on the current stream, L2 flushed
    (outside the event pairs) between them. Returns (total ms, per-step ms)."""
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if before:
        before()
    torch.cuda.synchronize()
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        step()
        ev[k][1].record()
    torch.cuda.synchronize()
    series = [a.elapsed_time(b) for a, b in ev]
    return sum(series), series


class ProductRun:
    """One scene on one GPU through pbf_step on device buffers (the product's C-ABI via the ctypes mirror)."""

    def __init__(self, pbf, torch, local, name, sc=None):
        self.pbf, self.torch, self.name = pbf, torch, name
        self.dev = torch.device("cuda", local)
        sc = self.sc = sc or pbf.SCENES[name]
        if "blocks" in sc:
            n = pbf.scene_particles(sc)
            pos = torch.empty((n, 3), dtype=torch.float32, device=self.dev)
            vel = torch.empty_like(pos)
            iid = torch.empty(n, dtype=torch.int32, device=self.dev)
            off = 0
            for origin, n3 in sc["blocks"]:
                pbf.scene_block_device(origin, n3, pos[off:], vel[off:], iid[off:], first_iid=off)
                off += int(np.prod(n3))
        else:
            p, v, i, _, _ = pbf.scene_double_dam_reference()
            n = len(i)
            pos, vel = torch.from_numpy(p).to(self.dev), torch.from_numpy(v).to(self.dev)
            iid = torch.from_numpy(i.astype(np.int64)).to(self.dev).to(torch.int32)
        torch.cuda.synchronize()
        self.n, self.iid = n, iid
        self.bufs = [pos, torch.zeros_like(pos), vel, torch.zeros_like(vel)]
        self.sim = pbf.Simulator(pbf.default_params(), sc.get("ulim_max", sc["ulim"]), sc["llim"], n, device=local)
        self.sim.setLim(sc["ulim"], sc["llim"])
        self.stream = torch.cuda.current_stream().cuda_stream
        self.frame = 0
        self.local = local

    def step(self):
        lim = lim_at(self.pbf, self.sc, self.frame)
        if lim is not None:
            self.sim.setLim(*lim)
        b = self.bufs
        self.sim.step(b[0], b[1], b[2], b[3], self.iid, self.n, self.stream)
        b[0], b[1] = b[1], b[0]
        b[2], b[3] = b[3], b[2]
        self.frame += 1

    def digest(self):
        return self.pbf.state_digest(self.bufs[0], self.bufs[2], self.iid, self.n, device=self.local, stream=self.stream)

    def close(self):
        self.sim.close()
        self.bufs = self.iid = None


def read_reference_file():
    try:
        with open(REF_FILE) as f:
            return json.load(f)
    except Exception:
        return {}


def size_entry(S, name, sc, n, steps, warmup, total_ms, digest, peak, ref):
    value = n * steps / (total_ms * 1e-3)
    e = {"scene": name, "workload": workload_name(S, name, sc, n), "particles": n, "steps": steps, "warmup": warmup,
         "ms_per_step": round(total_ms / steps, 5), "value": round(value, 1), "unit": "particle-steps/s",
         "whole_step_frac": round(ALG_BYTES_STEP * value / 1e9 / peak, 5), "digest": hexd(digest)}
    r = (ref or {}).get("%s/%d/%d" % (name, steps, warmup))
    if r:
        e["reference_value"] = r["value"]
        e["reference_ratio"] = round(value / r["value"], 2)
        e["reference_source"] = "bench.py --impl reference, same window, this box (gpurun_out/bench_reference_last.json)"
        if r.get("digest"):
            e["parity_vs_reference"] = "bit_exact" if r["digest"] == e["digest"] else "MISMATCH"
    return e


def run_product(args, rank, world, dist):
    import torch
    pbf = importlib.import_module("pbf-cuda_b200")   # raises if libpbf_b200.so is missing: no fallback
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # a stream of its own (torch's default is the legacy stream 0, which cannot be captured into a CUDA graph:
    # pbf_step replays small scenes from a graph, include/pbf.h PBF_OPT_GRAPH); events, flushes and steps all go here
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    run = ProductRun(pbf, torch, local, args.scene)
    sc, n, sim = run.sc, run.n, run.sim
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    sampler = ClockSampler(local, launch=(rank == 0))
    launches0 = [0]

    def before():
        if rank == 0:
            sampler.start()
        launches0[0] = sim.launch_count()
        if dist:
            dist.barrier()

    total_ms, series = timed_window(torch, run.step, args.steps, args.warmup, flush, before)
    if dist:
        dist.barrier()
    launches = sim.launch_count() - launches0[0]
    clocks = sampler.stop() if rank == 0 else None
    digest = run.digest()
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
        run.step()
    e1.record()
    torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1) / args.steps
    sim.enable_stage_timing(True)
    kacc, sacc, reps = {}, {}, 5
    for _ in range(reps):
        flush.zero_()
        run.step()
        for k, v in sim.kernel_ms().items():
            kacc[k] = kacc.get(k, 0.0) + v / reps
        for k, v in sim.stage_ms().items():
            sacc[k] = sacc.get(k, 0.0) + v / reps
    sim.enable_stage_timing(False)
    stats = sim.stats(run.bufs[0], run.bufs[2], n)

    if rank != 0:
        return None
    peak, peak_src = hbm_peak()
    dom = "lambda" if kacc["lambda"] >= kacc["delta_p"] else "delta_p"
    achieved = ALG_BYTES[dom] * n / (kacc[dom] * 1e-3) / 1e9
    step_gbs = ALG_BYTES_STEP * (value / world) / 1e9
    roofline = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": NCU[dom]["traffic"],
                "traffic_source": NCU["source"] + " — a separate ncu run, NOT measured in this run",
                "issue_slot_frac": NCU[dom]["issue"], "l1_frac": NCU[dom]["l1"], "dram_frac": NCU[dom]["dram"],
                "peak_source": peak_src, "kernel_ms": round(kacc[dom], 4),
                "algorithmic_bytes_per_particle": ALG_BYTES[dom],
                "whole_step": {"algorithmic_bytes_per_particle_step": ALG_BYTES_STEP, "achieved": round(step_gbs, 2),
                               "frac": round(step_gbs / peak, 5)},
                "note": "the lambda / XSPH sweeps are bound by the L1 wavefront rate and instruction issue, not by HBM "
                        "(SURVEY.md App. D, DESIGN.md 5): issue_slot_frac / l1_frac / dram_frac are the ncu figures of "
                        "the same kernel in the state the kernel timers see; the delta-p pass replays the lambda pass's "
                        "neighbour list and evaluates the exact powf (see delta_p below)",
                "delta_p": {"kernel_ms": round(kacc["delta_p"], 4), "traffic": NCU["delta_p"]["traffic"],
                            "issue_slot_frac": NCU["delta_p"]["issue"], "dram_frac": NCU["delta_p"]["dram"],
                            "dram_GBps_from_traffic": round(NCU["delta_p"]["traffic"] / (kacc["delta_p"] * 1e-3) / 1e9, 1)}}

    # ---- end to end through the host-buffer entry point --------------------------------------------
    # the SAME window of the SAME scene as `value`: a fresh initial state, `warmup` untimed steps, `steps`
    # timed steps, every one of them uploading its inputs from pinned host memory and downloading its result
    fresh = ProductRun(pbf, torch, local, args.scene)
    h = [torch.empty((n, 3), dtype=torch.float32).pin_memory() for _ in range(4)]
    h_iid = torch.empty(n, dtype=torch.int32).pin_memory()
    h[0].copy_(fresh.bufs[0]); h[2].copy_(fresh.bufs[2]); h_iid.copy_(fresh.iid)
    fresh.close()
    del fresh
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
    e2e_digest = pbf.state_digest(hn[0], hn[2], hi)
    e2e = {"value": round(n * args.steps / e2e_s, 1), "unit": "particle-steps/s", "steps": args.steps,
           "h2d_bytes_per_step": 28 * n, "d2h_bytes_per_step": 28 * n,
           "parity_vs_device_path": "bit_exact" if e2e_digest == digest else "MISMATCH",
           "api": "pbf_step_host (pinned host buffers; upload pos/vel/iid, step, download npos/nvel/iid), "
                  "same scene and step window as `value`, wall clock around the synchronous calls"}
    run.close()
    del run, h, h_iid, hn, hi
    torch.cuda.empty_cache()

    ref = read_reference_file() if world == 1 else {}
    out = {"metric": "particle-steps/s", "value": round(value, 1), "unit": "particle-steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 5),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(pbf, args.scene, sc, n), "particles_per_gpu": n,
                      "parallelism": "single GPU" if world == 1 else "replicas x%d (one independent scene per GPU)" % world,
                      "l2": "flushed between timed steps (256 MB memset outside the event pairs)",
                      "ms_per_step_back_to_back": round(b2b_ms, 5), "exact_pow": True,
                      "neighbour_list": dict(zip(("in_use", "bytes"), sim.pair_list())),
                      "ms_per_step_series": [round(x, 2) for x in series],
                      "stage_ms": {k: round(v, 4) for k, v in sacc.items()},
                      "kernel_ms": {k: round(v, 4) for k, v in kacc.items()},
                      "stats_after_run": {k: round(v, 6) for k, v in stats.items()}},
           "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "clocks": clocks,
           "digest": hexd(digest)}
    r = ref.get("%s/%d/%d" % (args.scene, args.steps, args.warmup))
    if r and r.get("digest"):
        out["parity_vs_reference"] = "bit_exact" if r["digest"] == out["digest"] else "MISMATCH"
    # ---- the other named single-GPU sizes (BASELINE configs 1, 3, 4) ----------------------------------
    if world == 1 and not args.no_sizes:
        sizes = []
        for name, steps, warmup in SIZES:
            if name == args.scene and steps == args.steps:
                continue
            try:
                q = ProductRun(pbf, torch, local, name)
                ms, _ = timed_window(torch, q.step, steps, warmup, flush)
                sizes.append(size_entry(pbf, name, q.sc, q.n, steps, warmup, ms, q.digest(), peak, ref))
                q.close()
                del q
                torch.cuda.empty_cache()
            except Exception as ex:   # noqa: BLE001 - an entry that failed says so, the headline stays
                sizes.append({"scene": name, "error": repr(ex)})
        out["sizes"] = sizes
    # ---- opt-in tolerance mode (north_star: lambda / delta-p / positions within 1e-5 relative; the headline above
    # is the bit-exact default): s_corr's w^4 as (w*w)^2 instead of the exact powf. Same scene, same window; the
    # error is measured the way the tolerance is defined — ONE step from an identical state, against the exact path.
    if world == 1 and not args.no_sizes:
        try:
            q = ProductRun(pbf, torch, local, args.scene)
            q.sim.set_exact_pow(False)
            ms, _ = timed_window(torch, q.step, args.steps, args.warmup, flush)
            state = [t.clone() for t in (q.bufs[0], q.bufs[2], q.iid)]
            q.step()                                     # one step, tolerance mode ...
            fast = (q.bufs[0].clone(), q.bufs[2].clone(), q.iid.clone())
            q.bufs[0].copy_(state[0]); q.bufs[2].copy_(state[1]); q.iid.copy_(state[2])
            q.frame -= 1
            q.sim.set_exact_pow(True)
            q.step()                                     # ... and the same step, bit-exact mode
            torch.cuda.synchronize()
            same_order = bool(torch.equal(fast[2], q.iid))
            dpos = float((fast[0] - q.bufs[0]).abs().max() / q.bufs[0].abs().max()) if same_order else None
            dvel = float((fast[1] - q.bufs[2]).abs().max() / q.bufs[2].abs().max()) if same_order else None
            out["tolerance_mode"] = {"option": "pbf_set_option_exact_pow(sim, 0): (w*w)^2 for n_corr = 4; everything else as in the default",
                                     "value": round(n * args.steps / (ms * 1e-3), 1), "unit": "particle-steps/s",
                                     "ms_per_step": round(ms / args.steps, 5), "steps": args.steps, "warmup": args.warmup,
                                     "one_step_max_rel_err_positions": dpos, "one_step_max_rel_err_velocities": dvel,
                                     "same_sorted_order_as_exact": same_order,
                                     "note": "NOT the headline: results differ from the reference's bits (within the stated 1e-5 on "
                                             "positions after one step; trajectories then diverge chaotically like any two fp32 runs)"}
            q.close()
            del q, state, fast
            torch.cuda.empty_cache()
        except Exception as ex:   # noqa: BLE001
            out["tolerance_mode"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, pbf.SCENES[args.scene], n)
    return out


# ---- N > 1: one scene across the ranks (x-slab decomposition, pbf-cuda_b200/slab.py) ----------------

def slab_scene(S, name, world, scaling):
    """Block list and box of the multi-GPU workload. weak: the scene's single block and its box are
    repeated `world` times along x (per-GPU work fixed: SURVEY.md 8d config 5 'weak'); strong: the named
    scene as it is, cut into `world` slabs."""
    sc = dict(S.SCENES[name])
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


def slab_leg(args, pbf, slab, torch, dist, rank, world, local, name, scaling, steps, warmup, full):
    """One scene over the `world` ranks: `warmup` + `steps` SlabSimulator steps, device-timed (max over ranks),
    then the digest of all owned particles; full = the headline leg (per-kernel / per-phase timers, e2e).
    Afterwards rank 0 runs the same scene for the same number of steps on ONE GPU and compares digests."""
    dev = torch.device("cuda", local)
    sc = slab_scene(pbf, name, world, scaling)
    params = pbf.default_params()
    planes = pbf.scene_dims(sc)[0]
    n_total = pbf.scene_particles(sc)
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
    sampler = ClockSampler(local, launch=(rank == 0 and full))
    mark = {}

    def before():
        sim.finish()
        if rank == 0 and full:
            sampler.start()
        mark["l"], mark["m"], mark["b"] = eng.sim.launch_count(), sim.messages, sim.bytes_sent
        dist.barrier()

    total_ms, series = timed_window(torch, sim.step, steps, warmup, flush, before)
    dist.barrier()
    sim.finish()
    launches = eng.sim.launch_count() - mark["l"]
    msgs, sent = sim.messages - mark["m"], sim.bytes_sent - mark["b"]
    clocks = sampler.stop() if (rank == 0 and full) else None
    own_ms = total_ms
    # the state after exactly warmup + steps steps: every rank's owned particles, digests combined over the ranks
    d = pbf.state_digest(*eng.state(), device=local, stream=torch.cuda.current_stream().cuda_stream)
    dt = torch.tensor([d[0] - (1 << 64) if d[0] >= (1 << 63) else d[0], d[1] - (1 << 64) if d[1] >= (1 << 63) else d[1]],
                      dtype=torch.int64, device=dev)
    alld = [torch.zeros_like(dt) for _ in range(world)]
    dist.all_gather(alld, dt)
    digest = pbf.combine_digests([(int(q[0]) & 0xFFFFFFFFFFFFFFFF, int(q[1]) & 0xFFFFFFFFFFFFFFFF) for q in alld])
    t = torch.tensor([total_ms, float(eng.n_own), -float(eng.n_own)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, n_max, n_min = float(t[0]), int(t[1]), int(-t[2])
    value = n_total * steps / (total_ms * 1e-3)
    peak, peak_src = hbm_peak()
    res = {"scene": name, "scaling": scaling, "workload": workload_name(pbf, name + (" x%d along x (weak)" % world if scaling == "weak" else ""), sc, n_total),
           "particles_total": n_total, "particles_per_gpu": {"min": n_min, "max": n_max}, "steps": steps, "warmup": warmup,
           "ms_per_step": round(total_ms / steps, 5), "value": round(value, 1), "unit": "particle-steps/s",
           "whole_step_frac_per_gpu": round(ALG_BYTES_STEP * value / world / 1e9 / peak, 5),
           "slab_boundaries": [int(b) for b in sim.bounds], "digest": hexd(digest)}
    extra = {}
    if full:
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
        pt = torch.tensor([phases.get(k, 0.0) for k in names] + [own_ms / steps, float(eng.n_own)], dtype=torch.float64, device=dev)
        allp = [torch.zeros_like(pt) for _ in range(world)]
        dist.all_gather(allp, pt)
        per_rank_ph = [{**{k: round(float(q[i]), 3) for i, k in enumerate(names)}, "ms_per_step": round(float(q[-2]), 3),
                        "particles": int(q[-1])} for q in allp]

        # ---- end to end: every rank's state starts each step in pinned HOST memory and ends there, the SAME number
        # of steps as `value`. Uploads go ahead of the step on its stream; the final positions and iid are complete
        # after the velocity update and come down on a copy stream while the XSPH sweep still runs (what
        # pbf_step_host does on one GPU); the velocities follow the sweep.
        cap = eng.capacity
        hp = [torch.empty((cap, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        hi = torch.empty(cap, dtype=torch.int32).pin_memory()
        n = eng.n_own
        hp[0][:n].copy_(eng.pos[:n]); hp[1][:n].copy_(eng.vel[:n]); hi[:n].copy_(eng.iid[:n])
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream()
        state = {"n": n}

        def after_velocity():
            m = int(eng.layout.own_count)
            state["n"] = m
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev)
                hp[0][:m].copy_(eng.npos[:m], non_blocking=True)

        torch.cuda.synchronize()
        dist.barrier()
        up = down = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            n = state["n"]
            eng.pos[:n].copy_(hp[0][:n], non_blocking=True); eng.vel[:n].copy_(hp[1][:n], non_blocking=True)
            eng.iid[:n].copy_(hi[:n], non_blocking=True)
            up += 28 * n
            sim.step(after_velocity=after_velocity)
            n = eng.n_own
            hp[1][:n].copy_(eng.vel[:n], non_blocking=True)
            hi[:n].copy_(eng.iid[:n], non_blocking=True)
            down += 28 * n
            torch.cuda.synchronize()
        dist.barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s, float(up), float(down)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sim.finish()
        stats = eng.sim.stats(eng.pos, eng.vel, eng.n_own)
        dom = "lambda" if kacc["lambda"] >= kacc["delta_p"] else "delta_p"
        achieved = ALG_BYTES[dom] * n_own / (kacc[dom] * 1e-3) / 1e9
        step_gbs = ALG_BYTES_STEP * (value / world) / 1e9
        extra["roofline"] = {"bound": "hbm", "kernel": dom + "_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 5), "traffic": None, "peak_source": peak_src,
                             "kernel_ms": round(kacc[dom], 4), "algorithmic_bytes_per_particle": ALG_BYTES[dom],
                             "particles_on_this_rank": n_own,
                             "whole_step": {"algorithmic_bytes_per_particle_step": ALG_BYTES_STEP, "achieved_per_gpu": round(step_gbs, 2),
                                            "frac": round(step_gbs / peak, 5)},
                             "note": "rank 0's kernels; the lambda / XSPH sweeps are L1-wavefront / issue bound, not HBM bound (DESIGN.md 5)"}
        extra["e2e"] = {"value": round(n_total * steps / float(t[0]), 1), "unit": "particle-steps/s", "steps": steps,
                        "h2d_bytes_per_step": int(float(t[1]) / steps), "d2h_bytes_per_step": int(float(t[2]) / steps),
                        "api": "every rank uploads its slab's pos/vel/iid from pinned host memory, SlabSimulator.step, downloads the "
                               "result (positions on a copy stream under the XSPH sweep, velocities + iid behind it); bytes are the "
                               "largest rank's"}
        extra["config"] = {"messages_per_step_rank0": round(msgs / steps, 1), "bytes_sent_per_step_rank0": int(sent / steps),
                           "kernel_ms_rank0": {k: round(v, 4) for k, v in kacc.items()}, "phase_ms_per_rank": per_rank_ph,
                           "ms_per_step_series_rank0": [round(x, 2) for x in series],
                           "stats_rank0": {k: round(v, 6) for k, v in stats.items()},
                           "parallelism": "x-slab decomposition over %d ranks (one scene): per step one exchange of the raw state of "
                                          "%d planes per side, then (2*niter+1) ghost refreshes %s; ghost=%d margin=%d replan_every=%d"
                                          % (world, ghost + margin,
                                             "FUSED into the pass kernels (stores into the neighbour's ghost slots over NVLink peer memory "
                                             "+ a flag handshake, no collective)" if sim.fused else "as NCCL send/recv of float4 ranges" +
                                             (" [%s]" % sim.fused_note if sim.fused_note else ""),
                                             ghost, margin, args.replan_every)}
        extra["gpu_launches"] = int(launches)
        extra["clocks"] = clocks
    torch.cuda.synchronize()
    dist.barrier()     # nobody frees arrays a neighbour may still be pushing ghost values into
    eng.close()
    del eng, sim, flush
    torch.cuda.empty_cache()
    # ---- parity: ONE GPU, the same scene, the same number of steps --------------------------------------
    if rank == 0 and not args.no_parity:
        try:
            q = ProductRun(pbf, torch, local, name, sc)
            for _ in range(warmup + steps):
                q.step()
            single = q.digest()
            q.close()
            del q
            torch.cuda.empty_cache()
            res["parity"] = "bit_exact" if single == digest else "MISMATCH"
            res["parity_how"] = ("digest of the %d ranks' owned particles after %d steps == digest of pbf_step on one GPU "
                                 "after %d steps (order-independent 128-bit digest of iid, pos, vel bits)" % (world, warmup + steps, warmup + steps))
        except Exception as ex:   # noqa: BLE001
            res["parity"] = "unchecked: %r" % (ex,)
    dist.barrier()
    return res, extra


def run_product_slab(args, rank, world, dist):
    import torch
    pbf = importlib.import_module("pbf-cuda_b200")   # raises if libpbf_b200.so is missing: no fallback
    slab = importlib.import_module("pbf-cuda_b200.slab")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", local)))
    head, extra = slab_leg(args, pbf, slab, torch, dist, rank, world, local, args.scene, args.scaling, args.steps, args.warmup, True)
    legs = []
    if not args.no_legs:
        for name, scaling, steps, warmup, min_world in LEGS:
            if world < min_world or (name == args.scene and scaling == args.scaling):
                continue
            try:
                leg, _ = slab_leg(args, pbf, slab, torch, dist, rank, world, local, name, scaling, min(steps, args.steps), warmup, False)
                if name == "dam_64m":
                    leg["also"] = "identical to BASELINE config 5 WEAK at 8 GPUs: the 8 388 608-particle block per GPU (dam_8m) repeated 8 times along x"
                legs.append(leg)
            except SystemExit as ex:
                legs.append({"scene": name, "scaling": scaling, "error": str(ex)})
    if rank != 0:
        return None
    cfg = {"workload": head["workload"], "particles_total": head["particles_total"], "particles_per_gpu": head["particles_per_gpu"],
           "slab_boundaries": head["slab_boundaries"],
           "l2": "flushed between timed steps (256 MB memset outside the event pairs)", "exact_pow": True}
    cfg.update(extra["config"])
    out = {"metric": "particle-steps/s", "value": head["value"], "unit": "particle-steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": cfg, "gpu_launches": extra["gpu_launches"], "e2e": extra["e2e"], "roofline": extra["roofline"],
           "clocks": extra["clocks"], "digest": head["digest"], "parity": head.get("parity"), "parity_how": head.get("parity_how"),
           "legs": legs}
    return out


# ncu --set full of the dominant kernels, dam_1m at step 100 (profiles/r03_solver_ncu_summary.txt, the build of the last
# session of round 2; means over the four launches of the step): dram__bytes_read.sum + dram__bytes_write.sum per launch,
# issue slots busy, l1tex throughput, DRAM throughput. A separate run under the profiler: labelled as such in the line.
NCU = {"source": "profiles/r03_solver_ncu_summary.txt (ncu --set full, dam_1m, step 100)",
       "lambda": {"traffic": 462.8e6, "issue": 0.70, "l1": 0.77, "dram": 0.12},
       "delta_p": {"traffic": 493.1e6, "issue": 0.63, "l1": 0.50, "dram": 0.32}}


def cpu_baseline(args, sc, n, steps=None):
    """The oracle port on the host cores: bounded sample = `steps` whole steps of the same workload."""
    import _oracle as O
    threads = os.cpu_count() or 1
    pos, vel, iid = host_scene(sc)
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


class ReferenceRun:
    """One scene through the reference's own Simulator.cu (oracle/_ref/libpbf_ref.so) on cuda:0. The initial state
    comes from the oracle's host generator; the product's library is not loaded."""

    def __init__(self, S, torch, name, sc=None):
        import _oracle as O
        import _ref
        self.S, self.torch = S, torch
        sc = self.sc = sc or S.SCENES[name]
        dev = torch.device("cuda", 0)
        p, v, i = host_scene(sc)
        self.n = len(i)
        pos, vel = torch.from_numpy(p).to(dev), torch.from_numpy(v).to(dev)
        self.iid = torch.from_numpy(i.view(np.int32)).to(dev)
        self.bufs = [pos, torch.zeros_like(pos), vel, torch.zeros_like(vel)]
        self.ref = _ref.RefSimulator(O.default_params(), sc.get("ulim_max", sc["ulim"]), sc["llim"], self.n)
        self.ref.set_lim(sc["ulim"], sc["llim"])
        self.frame = 0

    def step(self):
        lim = lim_at(self.S, self.sc, self.frame)
        if lim is not None:
            self.ref.set_lim(*lim)
        b = self.bufs
        self.ref.step(b[0], b[1], b[2], b[3], self.iid, self.n)
        b[0], b[1] = b[1], b[0]
        b[2], b[3] = b[3], b[2]
        self.frame += 1

    def digest(self):
        return digest_numpy(self.bufs[0].cpu().numpy(), self.bufs[2].cpu().numpy(), self.iid.cpu().numpy())

    def close(self):
        self.ref.close()
        self.bufs = self.iid = None


def run_reference(args, rank, world):
    """The reference arm: the reference's own Simulator.cu (headless, sm_100) on the same workloads; rank 0 only."""
    if rank != 0:
        return None
    import _ref
    S = load_scenes()
    sc = S.SCENES[args.scene]
    if not _ref.available():
        n = S.scene_particles(sc)
        cb = cpu_baseline(args, sc, n, steps=max(1, min(args.steps, 3)))
        return {"impl": "reference", "metric": "particle-steps/s", "value": cb["value"], "unit": "particle-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(S, args.scene, sc, n), "note": "oracle/_ref/libpbf_ref.so absent: CPU oracle port timed"},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    # N > 1: the product arm runs ONE scene over N GPUs (weak: the block repeated N times along x; strong: the
    # named scene); the single-GPU reference gets that same scene, whole, on one GPU
    sc_n = slab_scene(S, args.scene, world, args.scaling) if (world > 1 and not args.replicas) else None
    run = ReferenceRun(S, torch, args.scene, sc_n)
    sc, n = run.sc, run.n
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(0)
    total_ms, _ = timed_window(torch, run.step, args.steps, args.warmup, flush, sampler.start)
    clocks = sampler.stop()
    value = n * args.steps / (total_ms * 1e-3)
    record = {"%s/%d/%d" % (args.scene, args.steps, args.warmup): {"value": round(value, 1), "digest": hexd(run.digest())}} if world == 1 else {}
    run.close()
    del run
    torch.cuda.empty_cache()
    out = {"impl": "reference", "metric": "particle-steps/s", "value": round(value, 1), "unit": "particle-steps/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 5),
           "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(S, args.scene + (" x%d along x (weak)" % world if (sc_n is not None and args.scaling == "weak") else ""), sc, n),
                      "note": "the reference's own Simulator.cu + Simulator_kernel.cuh compiled unchanged for sm_100 "
                              "(oracle/_ref/libpbf_ref.so), Thrust sort and its cudaDeviceSynchronize fences kept, GL interop "
                              "excluded; runs on one GPU (the reference is single-GPU, rank 0 only); initial state from the "
                              "oracle's host generator, libpbf_b200.so not loaded",
                      "l2": "flushed between timed steps"},
           "cpu_baseline": {"value": round(value, 1), "unit": "particle-steps/s", "kind": "reference", "cores": 0,
                            "sample": "all %d steps; the reference has no CPU implementation of this path — it ran its "
                                      "CUDA path on cuda:0" % args.steps},
           "e2e": {"value": round(value, 1), "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "clocks": clocks}
    # the other named sizes, the same windows as the product arm (N=1: `sizes`; N>1: the legs' scenes on ONE GPU)
    if not args.no_sizes:
        todo = [(nm, st, wu) for nm, st, wu in SIZES] if world == 1 else \
               [(nm, min(st, 3) if nm == "dam_64m" else st, wu) for nm, _, st, wu, mw in LEGS if world >= mw]
        entries = []
        for name, steps, warmup in todo:
            if world == 1 and name == args.scene and steps == args.steps:
                continue
            try:
                q = ReferenceRun(S, torch, name)
                ms, _ = timed_window(torch, q.step, steps, warmup, flush)
                v = q.n * steps / (ms * 1e-3)
                e = {"scene": name, "workload": workload_name(S, name, q.sc, q.n), "particles": q.n, "steps": steps,
                     "warmup": warmup, "ms_per_step": round(ms / steps, 5), "value": round(v, 1), "unit": "particle-steps/s",
                     "digest": hexd(q.digest())}
                entries.append(e)
                if world == 1:
                    record["%s/%d/%d" % (name, steps, warmup)] = {"value": e["value"], "digest": e["digest"]}
                q.close()
                del q
                torch.cuda.empty_cache()
            except Exception as ex:   # noqa: BLE001
                entries.append({"scene": name, "error": repr(ex)})
        out["sizes" if world == 1 else "legs"] = entries
    if record:
        try:
            os.makedirs(os.path.dirname(REF_FILE), exist_ok=True)
            with open(REF_FILE, "w") as f:
                json.dump(record, f)
        except OSError:
            pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--scene", default="dam_1m")
    ap.add_argument("--cpu-steps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sizes", action="store_true", help="N=1: skip the `sizes` entries (the other named scenes)")
    ap.add_argument("--no-legs", action="store_true", help="N>1: skip the strong-scaling legs behind the headline")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the single-GPU digest comparison")
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
            bind_to_gpu_cpus(int(os.environ.get("LOCAL_RANK", 0)))
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
