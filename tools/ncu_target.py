"""The process ncu wraps: N plain steps of a named scene through pbf_step on a stream of its own — no event
timers (the library's kernel timers keep the velocity update a separate launch), no flush, no read-backs.

    ncu --set full --clock-control none --import-source on -k regex:"lambda_kernel|delta_p|xsph" -s <skip> -c <n> \
        -o gpurun_out/<name> python tools/ncu_target.py SCENE STEPS
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

bench = importlib.import_module("bench")
pbf = importlib.import_module("pbf-cuda_b200")
torch.cuda.set_stream(torch.cuda.Stream())
run = bench.ProductRun(pbf, torch, 0, sys.argv[1])
for _ in range(int(sys.argv[2])):
    run.step()
torch.cuda.synchronize()
print("launches per step:", run.sim.launch_count() / int(sys.argv[2]), "digest", bench.hexd(run.digest()))
