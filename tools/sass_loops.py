"""List the loops (backward branches) of one kernel in a cuobjdump -sass dump with their opcode mix.
usage: cuobjdump -sass lib.so > all.sass; python tools/sass_loops.py all.sass <mangled-name-substring> [min_instr]"""
import collections
import re
import sys

path, name = sys.argv[1], sys.argv[2]
min_instr = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lines = open(path).read().split("\n")
start = next(i for i, l in enumerate(lines) if "Function :" in l and name in l)
end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
ops = []
for l in lines[start:end]:
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops.append((int(m.group(1), 16), m.group(3), l.strip()))
print(lines[start].strip(), len(ops), "instructions")
FP64 = ("DFMA", "DADD", "DMUL", "DSETP")
for addr, op, l in ops:
    if op.startswith("BRA"):
        m = re.search(r"(0x[0-9a-f]+)", l.split("BRA")[1])
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            body = [o for o in ops if tgt <= o[0] <= addr]
            if len(body) < min_instr:
                continue
            cc = collections.Counter(o[1].split(".")[0] for o in body)
            d = sum(v for k, v in cc.items() if k in FP64)
            print("%#x -> %#x: %d instr, fp64 %d, %s" % (tgt, addr, len(body), d, dict(cc.most_common(14))))
