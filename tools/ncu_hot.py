"""Top stall-sample SASS lines of an `ncu --page source --csv` export (in program order)."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[iE].isdigit()]
tots = sum(int(r[iSm]) for r in data)
print("samples", tots, "instr", sum(int(r[iE]) for r in data), "static", len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
top = sorted(range(len(data)), key=lambda k: -int(data[k][iSm]))[:n]
for k in sorted(top):
    r = data[k]
    print(f"{k:5d} {int(r[iE]):9d} {100 * int(r[iSm]) / max(tots, 1):5.1f}%  {r[iS].strip()[:100]}")
