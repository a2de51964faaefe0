"""SASS evidence table: per kernel of libmgs.so, how many tcgen05 / TMEM / TMA / bulk-copy instructions it contains
(`cuobjdump -sass`, mnemonics per /opt/skills/guides/B200_PROFILING.md).  Runs without a GPU.

    python tools/sass_table.py [> profiles/round2_sass_table.txt]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "m_gat_graphsage_b200" / "libmgs.so"
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF",
             "SYNCS", "LDGSTS", "HMMA", "FFMA", "FFMA2", "FADD2", "LDS", "STS", "LDG", "STG", "REDG", "ATOMG", "MUFU"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            per[cur][op] += 1
            per[cur]["_total"] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    cols = [m for m in MNEMONICS if any(c[m] for c in per.values())]
    print(f"# cuobjdump -sass {LIB.name}: instruction counts per kernel (static SASS; sm_100a)")
    print("# " + " ".join(f"{c:>8s}" for c in ["total"] + cols) + "  kernel")
    tot = collections.Counter()
    for (name, c), dn in zip(per.items(), demangled):
        dn = re.sub(r"\(.*", "", dn).replace("mgs::", "").replace("(anonymous namespace)::", "")
        print("  " + " ".join(f"{c[m]:8d}" for m in ["_total"] + cols) + "  " + dn[:110])
        tot.update(c)
    print("# " + " ".join(f"{tot[m]:8d}" for m in ["_total"] + cols) + "  ALL KERNELS")


if __name__ == "__main__":
    main()
