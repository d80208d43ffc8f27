"""Instructions of the first kernel in a source-page csv (`ncu -i rep --page source --csv --print-source sass [--launch-skip n --launch-count 1]`)
with the most samples of one stall reason.  usage: python tools/ncu_top_stalls.py source.csv stall_long_sb [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
c = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[0].startswith("0x") and not r[0][:1].isdigit():
        break
    body.append(r)
key = sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
base = int(body[0][c["Address"]], 16)
tot = sum(int(r[c["# Samples"]]) for r in body)
for r in sorted(body, key=lambda r: -int(r[c[key]]))[:n]:
    print("%#07x  %s %5.2f%%  all %5.2f%%  %s" % (int(r[c["Address"]], 16) - base, key, 100 * int(r[c[key]]) / tot, 100 * int(r[c["# Samples"]]) / tot, r[c["Source"]][:110]))
