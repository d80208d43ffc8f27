#!/usr/bin/env python
"""ncu_summary.py -- prints the metrics quoted in profiles/*.md from an .ncu-rep (run where ncu is installed).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else -1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__inst_executed_realtime.avg.pct_of_peak_sustained_elapsed", "inst_executed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "lts__t_bytes.sum", "sm__cycles_active.avg"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-85s %-14s %s" % (w, units[i], [r[i][:60] for r in data]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
sec = secs[kidx]
h = sec["hdr"]
print("\nsource page of:", sec["name"])
st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = collections.Counter()
for r in sec["rows"]:
    for i in st:
        tot[h[i]] += int(r[i] or 0)
s = sum(tot.values())
print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / s) for k, v in tot.most_common(8)))
iS, iI, iSamp, iT = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
ops, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r in sec["rows"]:
    toks = r[iS].split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.split(".")[0]
    ops[op] += int(r[iI]); samp[op] += int(r[iSamp]); thr[op] += int(r[iT])
ti, ts = sum(ops.values()), sum(samp.values())
print("warp instructions:", ti)
for op, c in ops.most_common(22):
    print("  %-10s %11d %5.1f%%  lanes %4.1f  samples %5.1f%%" % (op, c, 100 * c / ti, thr[op] / max(c, 1), 100 * samp[op] / ts))
print("hottest instructions:")
for r in sorted(sec["rows"], key=lambda r: -int(r[iSamp]))[:14]:
    print("  %6s %10s  %-60s %s" % (r[iSamp], r[iI], r[iS].strip()[:60], {h[i][6:]: r[i] for i in st if int(r[i] or 0) > 0.25 * int(r[iSamp])}))
