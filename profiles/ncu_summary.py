#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per kernel duration, registers, occupancy, DRAM bytes, pipe
utilisation and the top warp stall reasons (pc-sampling).  Usage: python profiles/ncu_summary.py <file.ncu-rep>"""
import csv
import io
import subprocess
import sys


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    basic = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
             "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
             # the FP64 / DMMA dispatch model of DESIGN.md section 4.1 as counters (round 2): a DFMA-class instruction keeps the
             # shared FP64 pipe 2 cycles, a DMMA 16; "shared" = fp64 + tensor
             "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "sm__inst_executed.sum",
             "sm__issue_active.avg.pct_of_peak_sustained_active",
             "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"]
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    for r in rows[2:]:
        print("==== %s" % r[idx["Kernel Name"]][:90])
        for b in basic:
            if b in idx:
                print("   %-72s %14s %s" % (b, r[idx[b]], units[idx[b]]))
        vals = []
        for n in stall:
            try:
                vals.append((float(r[idx[n]].replace(",", "")), n.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
        tot = sum(v for v, _ in vals) or 1.0
        print("   stall samples: " + ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in sorted(vals, reverse=True)[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
