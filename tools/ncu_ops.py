"""Summarise an `ncu --page source --csv` export: executed warp instructions and stall samples per opcode."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data)
tots = sum(int(r[iSm]) for r in data)
print("total warp instr", tot, "static", len(data), "samples", tots)
ops, smp = collections.Counter(), collections.Counter()
for r in data:
    t = r[iS].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    ops[op] += int(r[iE])
    smp[op] += int(r[iSm])
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:12s} {c:10d} {100 * c / tot:5.1f}%  samples {100 * smp[op] / max(tots, 1):5.1f}%")
