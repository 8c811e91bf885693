#!/usr/bin/env python
"""Writes profiles/r01_ncu_traffic.json from an `ncu --set full` capture of one bench step: per phase of the iteration,
dram__bytes_read.sum + dram__bytes_write.sum of that kernel launch (bench.py reports it as `traffic`).
Usage: python profiles/ncu_traffic.py <file.ncu-rep> <config name, e.g. c4>"""
import csv
import io
import json
import os
import subprocess
import sys


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)


def main(path, config):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out, seen_gene = {}, 0
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * unit_scale(units[idx["dram__bytes_read.sum"]])
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * unit_scale(units[idx["dram__bytes_write.sum"]])
        dur = r[idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]]
        if "k_estep_rows" in name:
            phase = "estep_cells_mstep_u"
        elif "k_gene_pass" in name:
            phase = "estep_genes" if seen_gene == 0 else "sparse_gene_pass"
            seen_gene += 1
        elif "k_zi_cells_mma" in name:
            phase = "zi_pass_cells"
        elif "k_zi_genes_mma" in name or "k_zi_genes_stored" in name:
            phase = "zi_pass_genes"
        else:
            continue
        out[phase] = rd + wr
        print("%-22s %-60s read %.3f GB write %.3f GB  (%s under ncu)" % (phase, name[:60], rd / 1e9, wr / 1e9, dur))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r01_ncu_traffic.json")
    allcfg = {}
    if os.path.exists(dst):
        with open(dst) as f:
            allcfg = json.load(f)
    allcfg.setdefault(config, {}).update(out)
    with open(dst, "w") as f:
        json.dump(allcfg, f, indent=1, sort_keys=True)
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
