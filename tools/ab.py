"""A/B of library variants / options on a named scene: ms per step over a window, per-kernel times at its end, and the
state digest (every variant must print the same one: the variants differ in speed, never in bits).

    python tools/ab.py SCENE WARMUP STEPS  name[:ENV=VAL[,ENV=VAL...]] ...

Each variant runs in its own process (the PBF_* variables are read when the handle is created; PBF_LIB picks a
tuning build made with `make -C pbf-cuda_b200 VARIANT=... EXTRA=...`). L2 is not flushed: back-to-back steps.
"""
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(scene, warm, steps):
    sys.path.insert(0, ROOT)
    import torch
    bench = importlib.import_module("bench")
    pbf = importlib.import_module("pbf-cuda_b200")
    torch.cuda.set_stream(torch.cuda.Stream())   # (a capturable stream: PBF_OPT_GRAPH)
    run = bench.ProductRun(pbf, torch, 0, scene)
    for _ in range(warm):
        run.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    d = run.digest()
    run.sim.enable_stage_timing(True)
    acc = {}
    for _ in range(5):
        run.step()
        for k, v in run.sim.kernel_ms().items():
            acc[k] = acc.get(k, 0.0) + v / 5
    print(json.dumps({"ms_per_step": round(ms, 4), "Mps": round(run.n / ms / 1e3, 1), "digest": bench.hexd(d),
                      "kernel_ms": {k: round(v, 4) for k, v in acc.items()}}))


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    scene, warm, steps = sys.argv[1], sys.argv[2], sys.argv[3]
    for spec in sys.argv[4:]:
        name, _, envs = spec.partition(":")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        if "PBF_LIB" in env and not os.path.isabs(env["PBF_LIB"]):
            env["PBF_LIB"] = os.path.join(ROOT, env["PBF_LIB"])
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", scene, warm, steps], env=env,
                           capture_output=True, text=True)
        out = r.stdout.strip().splitlines()
        print("%-18s %s %s" % (name, scene, out[-1] if out else "FAILED rc=%d %s" % (r.returncode, r.stderr[-400:])), flush=True)
