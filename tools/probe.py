"""Scratch GPU probe: times the product and the reference's own CUDA build on a named scene."""
import importlib, os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pbf = importlib.import_module("pbf-cuda_b200")
import _ref, _oracle

def make_state(scene_name):
    sc = pbf.SCENES[scene_name]
    dev = torch.device("cuda:0")
    if "blocks" in sc:
        n = sum(int(np.prod(b[1])) for b in sc["blocks"])
        pos = torch.empty((n, 3), dtype=torch.float32, device=dev); vel = torch.empty_like(pos)
        iid = torch.empty(n, dtype=torch.int32, device=dev)
        off = 0
        for origin, n3 in sc["blocks"]:
            m = int(np.prod(n3))
            pbf.scene_block_device(origin, n3, pos[off:], vel[off:], iid[off:], first_iid=off)
            off += m
    else:
        p, v, i, _, _ = pbf.scene_double_dam_reference()
        n = len(i)
        pos = torch.from_numpy(p).to(dev); vel = torch.from_numpy(v).to(dev)
        iid = torch.from_numpy(i.astype(np.int64)).to(dev).to(torch.int32)
    torch.cuda.synchronize()
    return sc, n, pos, vel, iid

def run(scene_name, impl, steps, warm):
    sc, n, pos, vel, iid = make_state(scene_name)
    npos = torch.zeros_like(pos); nvel = torch.zeros_like(vel)
    params = pbf.default_params()
    ulim = sc.get("ulim_max", sc["ulim"])
    if impl == "product":
        sim = pbf.Simulator(params, ulim, sc["llim"], n); sim.setLim(sc["ulim"], sc["llim"])
        sim.enable_stage_timing(True)
        stepf = lambda a, b, c, d: sim.step(a, b, c, d, iid, n)
    else:
        op = _oracle.Params(); C.memmove(C.byref(op), C.byref(params), C.sizeof(op))
        sim = _ref.RefSimulator(op, ulim, sc["llim"], n); sim.set_lim(sc["ulim"], sc["llim"])
        stepf = lambda a, b, c, d: sim.step(a, b, c, d, iid, n)
    bufs = [pos, npos, vel, nvel]
    def one():
        stepf(bufs[0], bufs[1], bufs[2], bufs[3])
        bufs[0], bufs[1] = bufs[1], bufs[0]; bufs[2], bufs[3] = bufs[3], bufs[2]
    for _ in range(warm): one()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): one()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    extra = ""
    if impl == "product":
        extra = " stages " + str({k: round(v, 4) for k, v in sim.stage_ms().items()}) + " launches/step %d" % (sim.launch_count() // (steps + warm))
        st = sim.stats(bufs[0], bufs[2], n); extra += " stats " + str({k: round(v, 5) for k, v in st.items()})
    print("%-14s %-9s n=%9d  %.4f ms/step  %.4f G particle-steps/s%s" % (scene_name, impl, n, ms, n / ms / 1e6, extra), flush=True)

if __name__ == "__main__":
    scenes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["double_dam_32k", "dam_1m"]
    impls = sys.argv[2].split(",") if len(sys.argv) > 2 else ["product", "reference"]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    for s in scenes:
        for i in impls:
            try:
                run(s, i, steps, 5)
            except Exception as e:
                print(s, i, "FAILED", repr(e), flush=True)
