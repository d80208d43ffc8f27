"""Stall samples and executed instructions per code region of one kernel launch in an .ncu-rep
(`--set full --import-source on`).  Regions are split at BAR.SYNC instructions and at user-given opcode
markers.  usage: python tools/ncu_regions.py source.csv   (csv from `ncu -i rep --page source --csv --print-source sass`)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
# first kernel only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
body = rows[hdr_i + 1:end]
c = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_s = sum(int(r[c["# Samples"]]) for r in body)
tot_i = sum(int(r[c["Instructions Executed"]]) for r in body)
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
region, regions = [], []
for r in body:
    region.append(r)
    if "BAR.SYNC" in r[c["Source"]] or r[c["Source"]].strip().startswith("EXIT"):
        regions.append(region)
        region = []
if region:
    regions.append(region)
base = int(body[0][c["Address"]], 16)
for reg in regions:
    s = sum(int(r[c["# Samples"]]) for r in reg)
    n = sum(int(r[c["Instructions Executed"]]) for r in reg)
    if s < tot_s * 0.005:
        continue
    top = sorted(((sum(int(r[c[h]]) for r in reg), h) for h in stalls), reverse=True)[:5]
    a0, a1 = int(reg[0][c["Address"]], 16) - base, int(reg[-1][c["Address"]], 16) - base
    fp64 = sum(int(r[c["Instructions Executed"]]) for r in reg if r[c["Source"]].strip().split()[0].lstrip("@!P0123456789 ").startswith(("DFMA", "DMUL", "DADD", "DSETP")))
    print("%#06x-%#06x: samples %5.1f%%  instr %5.1f%% (fp64 %4.1f%% of all)  %s" % (a0, a1, 100.0 * s / tot_s, 100.0 * n / tot_i, 100.0 * fp64 / tot_i,
          ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / tot_s) for v, h in top)))
