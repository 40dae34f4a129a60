"""scenes.py — the named benchmark scenes (BASELINE.md section 3 / SURVEY.md 8d) and the moving-wall schedule.

Pure numpy, no library: bench.py's reference arm loads THIS FILE by path (not the package, whose import maps
libpbf_b200.so) so that the arm that times the reference never has the product's library in its process.
"""
import numpy as np


def wall_lim(ulim0, llim0, a_ulim, a_llim, w, frame, start_frame=0):
    """Moving-wall schedule of FluidSystem::stepSimulate (FluidSystem.cpp:104-110):
    float t = w*(frame-start); float phi = sin(t); lim = lim0 + A*phi (all float)."""
    t = np.float32(np.float32(w) * np.float32(frame - start_frame))
    phi = np.float32(np.sin(np.float64(t)))
    u = np.asarray(ulim0, np.float32) + np.asarray(a_ulim, np.float32) * phi
    l = np.asarray(llim0, np.float32) + np.asarray(a_llim, np.float32) * phi
    return u.astype(np.float32), l.astype(np.float32)


# The named benchmark scenes (BASELINE.md section 3 / SURVEY.md 8d).
SCENES = {
    "double_dam_32k": dict(ulim=(2.0, 2.0, 4.0), llim=(-2.0, -2.0, 0.0), n=32000),
    # intermediate sizes (tuning of the small-scene kernels; same shape as dam_1m, scaled)
    "dam_128k": dict(ulim=(8.0, 2.0, 4.8), llim=(0.0, 0.0, 0.0), blocks=[((0.2, 0.2, 0.2), (64, 32, 64))]),
    "dam_256k": dict(ulim=(8.0, 3.6, 4.8), llim=(0.0, 0.0, 0.0), blocks=[((0.2, 0.2, 0.2), (64, 64, 64))]),
    "dam_1m": dict(ulim=(16.0, 3.6, 9.6), llim=(0.0, 0.0, 0.0), blocks=[((0.2, 0.2, 0.2), (128, 64, 128))]),
    "sweep_4m": dict(ulim=(19.2, 6.8, 9.6), llim=(0.0, 0.0, 0.0), blocks=[((0.2, 0.2, 0.2), (256, 128, 128))],
                     wall=dict(a_ulim=(4.8, 0.0, 0.0), a_llim=(0.0, 0.0, 0.0), w=0.05), ulim_max=(24.0, 6.8, 9.6)),
    "double_dam_16m": dict(ulim=(38.4, 38.4, 9.6), llim=(0.0, 0.0, 0.0),
                           blocks=[((0.2, 25.4, 0.2), (256, 256, 128)), ((25.4, 0.2, 0.2), (256, 256, 128))]),
    # per-GPU block of the 64M weak-scaling run (SURVEY.md 8d config 5 "weak": box x-extent 9.6 per GPU)
    "dam_8m": dict(ulim=(9.6, 26.0, 9.6), llim=(0.0, 0.0, 0.0), blocks=[((0.2, 0.2, 0.2), (128, 512, 128))]),
    "dam_64m": dict(ulim=(76.8, 26.0, 9.6), llim=(0.0, 0.0, 0.0), blocks=[((0.2, 0.2, 0.2), (1024, 512, 128))]),
}


def scene_dims(sc, h=0.1):
    """Grid dimensions ceil((ulim - llim) / h) in fp32, as the reference recomputes them (Simulator.cu:187-188)."""
    return [int(np.ceil(np.float32(np.float32(u) - np.float32(l)) / np.float32(h))) for u, l in zip(sc["ulim"], sc["llim"])]


def scene_particles(sc):
    return sum(int(np.prod(b[1])) for b in sc["blocks"]) if "blocks" in sc else int(sc["n"])
