"""Per-kernel launch count, total time and share from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys

def main(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        if r[ix["Metric Unit"]].startswith("us"):
            v *= 1e3
        elif r[ix["Metric Unit"]].startswith("ms"):
            v *= 1e6
        tot[name] += v
        cnt[name] += 1
    SETUP = ("scene", "stats", "const_div_check", "spiky_check", "pow4_check", "digest")   # one-off kernels outside the step
    own = sum(v for k, v in tot.items() if "pbf::" in k and not any(s in k for s in SETUP))
    print("%-66s %5s %12s %7s %10s" % ("kernel", "n", "total_ns", "share", "avg_ns"))
    for k, v in tot.most_common():
        step = "pbf::" in k and not any(s in k for s in SETUP)
        print("%-66s %5d %12d %7s %10d" % (k[:66], cnt[k], v, "%.4f" % (v / own) if step else "(n/a)", v / cnt[k]))

if __name__ == "__main__":
    main(sys.argv[1])
