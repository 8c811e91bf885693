#!/usr/bin/env python
"""Hot SASS instructions of one kernel from an .ncu-rep source page: per-opcode totals and the top individual
instructions by stall samples.  Usage: python profiles/ncu_source_hot.py <file.ncu-rep> <kernel regex> [topN]"""
import csv, io, subprocess, sys, collections

def main(path, kre, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # first kernel only
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in rows[hi + 1:]:
        if not r or r[0] == "Kernel Name": break
        body.append(r)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = 0; byop = collections.Counter(); execs = collections.Counter(); items = []
    for r in body:
        s = int(r[idx["# Samples"]] or 0); ex = int(r[idx["Instructions Executed"]] or 0)
        op = r[idx["Source"]].split()[0] if r[idx["Source"]].split() else "?"
        if op.startswith("@"): op = r[idx["Source"]].split()[1]
        op = op.split(".")[0]
        tot += s; byop[op] += s; execs[op] += ex
        reasons = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        items.append((s, ex, r[idx["Source"]][:70], reasons))
    print("total samples", tot, "instructions", len(body), "executed", sum(execs.values()))
    print("by opcode (samples%, executed%):")
    te = sum(execs.values()) or 1
    for op, s in byop.most_common(14):
        print("   %-10s %5.1f%%  %5.1f%%" % (op, 100.0 * s / tot, 100.0 * execs[op] / te))
    print("top instructions:")
    for s, ex, src, reasons in sorted(items, reverse=True)[:top]:
        print("   %5.1f%%  %-70s %s" % (100.0 * s / tot, src, ", ".join("%s=%d" % (n, v) for v, n in reasons)))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
