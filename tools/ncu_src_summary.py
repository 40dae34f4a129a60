"""Summarise an `ncu --page source --csv` dump: executed instructions and stall samples by opcode."""
import csv, collections, sys
def main(path, nwarps):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows if r and r[0].startswith("0x") and len(r) == len(hdr)]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    byop, samp, tot = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ix["Source"]].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        byop[op] += int(r[ix["Instructions Executed"]])
        samp[op] += int(r[ix["# Samples"]])
        for c in stall_cols:
            tot[c] += int(r[ix[c]])
    total = sum(byop.values())
    print("SASS instructions: %d; executed warp-instr: %d (%.0f per warp)" % (len(data), total, total / nwarps))
    print("executed per warp by opcode:", [(k, round(v / nwarps, 1)) for k, v in byop.most_common(32)])
    print("stall samples by opcode:", samp.most_common(14), "total", sum(samp.values()))
    print("stall reasons:", tot.most_common(12))
if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
