"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
iN, iV, iM = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iU = hdr.index("Metric Unit")
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v = v / 1e3 if r[iU] in ("ns", "nsecond") else v
    name = re.sub(r"\(.*", "", r[iN])
    name = name.replace("void ", "").replace("mgs::", "").replace("<unnamed>::", "")[:100]
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"# {sum(cnt.values())} kernel launches, sum of durations {s:.1f} us "
      "(cold-cache, serialised under ncu: compare SHARES, not absolutes)")
print("#         us  share count  kernel")
for name, v in tot.most_common():
    print(f"{v:12.1f} {100 * v / s:5.1f}% x{cnt[name]:4d}  {name}")
