"""Debug aid: first field (stage / iteration) at which the product and the reference library differ on a scene."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pbf = importlib.import_module("pbf-cuda_b200")
import _oracle as O, _trace as T

def onecell(n=300, seed=3):
    rng = np.random.RandomState(seed)
    pos = (np.float32([0.55, 0.55, 0.55]) + rng.rand(n, 3).astype(np.float32) * np.float32(0.04)).astype(np.float32)
    return dict(name="onecell", params=O.default_params(), ulim=np.float32([1, 1, 1]), llim=np.float32([0, 0, 0]),
                pos=pos, vel=np.zeros_like(pos), iid=np.arange(n, dtype=np.uint32), steps=1, wall=None)

scene = onecell() if len(sys.argv) < 2 or sys.argv[1] == "onecell" else T.make_scene(sys.argv[1])
a = T.trace_product(scene, pbf)
b = T.trace_reference(scene)
for k in b:
    x, y = np.ascontiguousarray(a[k]), np.ascontiguousarray(b[k])
    same = x.tobytes() == y.tobytes()
    bad = 0 if same else int((x.reshape(len(x), -1).view(np.uint32) != y.reshape(len(y), -1).view(np.uint32)).any(axis=1).sum())
    print("%-12s %s %d" % (k, "ok" if same else "DIFF", bad))
    if not same:
        idx = np.nonzero((x.reshape(len(x), -1).view(np.uint32) != y.reshape(len(y), -1).view(np.uint32)).any(axis=1))[0]
        print("  first bad rows", idx[:10], "ncount there", a["s0.ncount"][idx[:10]] if "s0.ncount" in a else None)
        break
