#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top SASS instructions by stall samples and the stall-reason totals.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<k> > src.csv && python tools/ncu_source_top.py src.csv [N]
"""
import csv
import sys


def sections(path):
    cur, hdr = None, None
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            if cur:
                yield cur
            cur = {"name": r[1], "hdr": None, "rows": []}
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    if cur:
        yield cur


def main():
    path = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    for sec in sections(path):
        hdr = sec["hdr"]
        idx = {h: i for i, h in enumerate(hdr)}
        data = sec["rows"]
        num = lambda r, h: int(float(r[idx[h]] or 0))
        tot = sum(num(r, "# Samples") for r in data)
        inst = sum(num(r, "Instructions Executed") for r in data)
        print("== %s: %d SASS instructions, %d warp-instructions executed, %d samples" % (sec["name"][:60], len(data), inst, tot))
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(num(r, h) for r in data) for h in stalls}
        print("   stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
        for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:ntop]:
            s = num(r, "# Samples")
            st = {h[6:]: num(r, h) for h in stalls if num(r, h) > 0.15 * s}
            print("   %6s %5.1f%% exec=%-8d %-72s %s" % (r[idx["Address"]][-5:], 100.0 * s / max(tot, 1), num(r, "Instructions Executed"), r[idx["Source"]][:72], st))
        break


if __name__ == "__main__":
    main()
