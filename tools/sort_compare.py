"""The product's onesweep sort next to cub::DeviceRadixSort::SortPairs (comparison only; tools/cub_sort.cu).
Ours: device time of hist_scan + all onesweep passes of one pbf_step (kernel timer slot `sort`; the digit
histograms are produced by advect_key, whose whole time is printed beside it as the upper bound of that share)."""
import importlib, json, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pbf = importlib.import_module("pbf-cuda_b200")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import probe

for scene in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["dam_1m", "double_dam_16m"]):
    sc, n, pos, vel, iid = probe.make_state(scene)
    npos, nvel = torch.zeros_like(pos), torch.zeros_like(vel)
    sim = pbf.Simulator(pbf.default_params(), sc.get("ulim_max", sc["ulim"]), sc["llim"], n)
    sim.setLim(sc["ulim"], sc["llim"])
    sim.enable_stage_timing(True)
    best = {}
    bufs = [pos, npos, vel, nvel]
    for _ in range(8):
        sim.step(bufs[0], bufs[1], bufs[2], bufs[3], iid, n)
        bufs[0], bufs[1], bufs[2], bufs[3] = bufs[1], bufs[0], bufs[3], bufs[2]
        torch.cuda.synchronize()
        for k, v in sim.kernel_ms().items():
            best[k] = min(best.get(k, 1e9), v)
    cells = int(np.prod(sim.grid_dim()))
    sim.close()
    cub = json.loads(subprocess.run([os.path.join(ROOT, "tools", "cub_sort"), str(n), str(cells)], capture_output=True, text=True, check=True).stdout)
    print(json.dumps({"scene": scene, "n": n, "cells": cells, "ours_sort_us": round(best["sort"] * 1e3, 1),
                      "ours_advect_key_us (includes the histograms)": round(best["advect_key"] * 1e3, 1),
                      "cub_best_us": cub["best_us"], "cub_median_us": cub["median_us"], "key_bits": cub["key_bits"]}))
