"""Generate the golden vectors that pin the oracle (and the product) to the reference itself.

The reference ships no tests or fixtures for this path, so the known answers are
produced by running the reference's OWN Simulator.cu — compiled unchanged into
oracle/_ref/libpbf_ref.so (oracle/Makefile target `ref`, built where /root/reference is mounted)
— on a B200, stage by stage, on the small parity scenes of tests/_trace.py:

    gpurun -- python tests/golden/make_golden.py          # writes gpurun_out/golden/*.npz + report
    cp gpurun_out/golden/*.npz tests/golden/               # commit

Small scenes are stored in full; the 32 000-particle reference scene stores every 31st sorted
slot plus whole-array checksums. It also prints how far the CPU oracle and the product are from
the reference on the same inputs (gpurun_out/golden/parity_report.txt).
"""
import hashlib
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path.insert(0, TESTS)
sys.path.insert(0, ROOT)

import _trace as T  # noqa: E402

SCENES = ["cube2k", "floor2k", "wall2k", "ragged", "dd32k"]
SUBSAMPLE = {"dd32k": 31}


def pack(name, trace):
    """Reduce a trace to what is committed."""
    step = SUBSAMPLE.get(name)
    out = {}
    for k, v in trace.items():
        field = k.split(".")[1]
        if step is None:
            out[k] = v
            continue
        per_cell = field in ("start", "end")
        if v.dtype.kind in "iu":
            out[k + ".sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(v).tobytes()).digest(), np.uint8).copy()
            if field == "key":
                out[k] = v  # sorted keys compress to almost nothing
            elif not per_cell:
                out[k + ".sub"] = v[::step].copy()
        else:
            out[k + ".sub"] = v[::step].copy()
            out[k + ".sum"] = np.array([v.astype(np.float64).sum(), np.abs(v.astype(np.float64)).sum()])
    return out


def main():
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    pbf = importlib.import_module("pbf-cuda_b200")
    report = []
    for name in SCENES:
        scene = T.make_scene(name)
        ref = T.trace_reference(scene)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **pack(name, ref))
        orc = T.trace_oracle(scene, threads=8)
        report.append(T.format_report("== %s: oracle vs reference (n=%d)" % (name, len(scene["iid"])), T.compare(orc, ref)))
        for exact in (True, False):
            try:
                prod = T.trace_product(scene, pbf, exact_pow=exact)
                report.append(T.format_report("== %s: product (exact_pow=%s) vs reference" % (name, exact), T.compare(prod, ref)))
            except Exception as e:  # keep the golden run alive if the product is broken
                report.append("== %s: product (exact_pow=%s) FAILED: %r" % (name, exact, e))
        try:
            prod = T.trace_product(scene, pbf, use_step=True)
            report.append(T.format_report("== %s: product pbf_step vs reference" % name, T.compare(prod, ref)))
        except Exception as e:
            report.append("== %s: product pbf_step FAILED: %r" % (name, e))
    txt = "\n".join(report)
    with open(os.path.join(outdir, "parity_report.txt"), "w") as f:
        f.write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
