#!/usr/bin/env python
"""Per-kernel share of the timed region from an ncu launch list (gpu__time_duration.sum per launch, cold-cache and
serialised: compare SHARES with bench.py's own CUDA-event phases, not absolutes).
Usage: python profiles/launch_shares.py <launches.csv> <out.txt> <timed steps>"""
import collections
import csv
import io
import sys


def main(path, out, steps):
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(txt)))
    names = [r["Kernel Name"] for r in rows]
    t = [float(r["Metric Value"]) for r in rows]
    # the timed region starts after the zero-inflation cell pass of the last warm-up step
    idx = [i for i, n in enumerate(names) if "k_zi_cells_mma" in n or "k_zi_cells_i8" in n]
    start = idx[-(steps + 1)] + 1
    # and ends before the roofline micro-benchmarks bench.py runs afterwards
    stop = next((i for i in range(start, len(names)) if "k_gather_peak" in names[i] or "k_fill_idx" in names[i]), len(names))
    agg, cnt = collections.Counter(), collections.Counter()
    for n_, v in zip(names[start:stop], t[start:stop]):
        k = n_.split("(")[0][:52]
        agg[k] += v
        cnt[k] += 1
    tot = sum(agg.values())
    lines = ["ncu launch list (gpu__time_duration.sum, --clock-control none) of `python bench.py --steps %d --no-cpu --no-e2e`, C4, 1 GPU" % steps,
             "timed region = the last %d steps + the final ELBO data-term gene pass of pcmf_iterate" % steps,
             "%d launches in the whole process; %d in the timed region" % (len(rows), stop - start), ""]
    for k, v in agg.most_common(16):
        lines.append("%-54s %3d launches %9.3f ms  %5.1f%%" % (k, cnt[k], v / 1e6, 100 * v / tot))
    lines.append("total %.1f ms" % (tot / 1e6))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
