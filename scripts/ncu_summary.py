#!/usr/bin/env python3
"""Compact summary of an .ncu-rep: key raw metrics + the hottest SASS lines by stall samples.
usage: ncu_summary.py report.ncu-rep [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 8
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_lookup_hit.sum", "lts__t_sectors_lookup_miss.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    if h in KEYS:
        print("%-70s %s %s" % (h, v, u))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
agg = {}
for r in data:
    for c in stall_cols:
        agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
print("samples", tot, "sass lines", len(data), "stalls:", ", ".join("%s=%.0f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]))
for idx, r in sorted(enumerate(data), key=lambda x: -int(x[1][iS]))[:top]:
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols), reverse=True)[0]
    print("  #%-5d %-52s samp %5.1f%% inst %-9s thr %-3s %s" % (idx, r[1].strip()[:52], 100.0 * int(r[iS]) / max(tot, 1), r[iI], r[iT], st[1]))
