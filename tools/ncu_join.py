"""Join an ``ncu --page raw --csv`` export of ONE bench step with the ordered C-ABI call log of the same step
(``bench.py --ncu-steps 1 --call-log``) -> per-call ncu numbers (``profiles/round2_ncu_calls.json``, read by bench.py).

    python tools/ncu_join.py full_raw.csv calls.json out.json

The launch list is in program order; kernels that are not ours (``at::``, cub, cutlass, nccl, ...: PyTorch's own) are
skipped, the rest are dealt to the calls in order using the per-call kernel counts of the log (taken from
``mgs_launch_count()``).  Per call: kernels, summed duration, DRAM bytes read + written, and for the longest kernel of
the call the tensor-pipe / issue / occupancy / DRAM-throughput counters.
"""
import csv
import json
import re
import sys
import time

FOREIGN = re.compile(r"^(void )?(at::|at_cuda_detail|cub::|cutlass|nccl|ncclDevKernel|void cutlass|std::|thrust::|c10::|"
                     r"cudnn|void cudnn|void cask|sm\d+_xmma|ampere_|void gemv|void gemm|void dot|void splitK|cublas|void cublas)")
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "warp_inst",
}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def to_us(v, unit):
    return {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1.0) * v


def to_bytes(v, unit):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1) * v


def main():
    raw, calls_path, out_path = sys.argv[1:4]
    rows = list(csv.reader(l for l in open(raw, newline="") if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col["Kernel Name"]
    launches = []
    for r in data:
        if len(r) != len(hdr):
            continue
        k = {"name": r[name_i]}
        for metric, short in WANT.items():
            if metric in col:
                v = num(r[col[metric]])
                u = units[col[metric]]
                if v is not None and short == "duration":
                    v = to_us(v, u)
                elif v is not None and short in ("dram_read", "dram_write", "l2_bytes"):
                    v = to_bytes(v, u)
                k[short] = v
        launches.append(k)
    ours = [k for k in launches if not FOREIGN.match(k["name"])]
    log = json.load(open(calls_path))
    want = sum(c["kernels"] for c in log["calls"])
    if want != len(ours):
        print(f"WARNING: call log expects {want} libmgs kernels, the capture has {len(ours)} non-PyTorch kernels "
              f"(of {len(launches)}); joining in order as far as it goes", file=sys.stderr)
    out, pos, seen = {}, 0, {}
    for c in log["calls"]:
        ks = ours[pos:pos + c["kernels"]]
        pos += c["kernels"]
        if not ks:
            continue
        seen[c["call"]] = seen.get(c["call"], 0) + 1
        key = c["call"] if seen[c["call"]] == 1 else f"{c['call']}#{seen[c['call']]}"
        top = max(ks, key=lambda k: k.get("duration") or 0.0)
        short_names = [re.sub(r"\(.*", "", k["name"]).replace("void ", "").replace("mgs::", "").replace("<unnamed>::", "")
                       for k in ks]
        out[key] = {
            "dram_bytes": int(sum((k.get("dram_read") or 0) + (k.get("dram_write") or 0) for k in ks)),
            "duration_us_under_ncu": round(sum(k.get("duration") or 0.0 for k in ks), 2),
            "kernels": short_names,
            "top_kernel": {s: top.get(s) for s in ("tensor_pipe_pct", "warps_active_pct", "issue_pct", "dram_throughput_pct",
                                                  "l1tex_pct", "regs", "warp_inst") if top.get(s) is not None},
        }
    total = sum(k.get("duration") or 0.0 for k in launches)
    mine = sum(k.get("duration") or 0.0 for k in ours)
    json.dump({"captured_on": time.strftime("%Y-%m-%d"), "source": raw, "launches": len(launches), "libmgs_launches": len(ours),
               "sum_duration_us_under_ncu": round(total, 1), "libmgs_share_of_duration": round(mine / max(total, 1e-9), 4),
               "note": "ncu --set full --clock-control none over one bench.py training step (batch 4096); durations are "
                       "cold-cache and serialised: compare shares, not absolutes", "calls": out},
              open(out_path, "w"), indent=1)
    print(f"{len(out)} calls joined from {len(ours)} kernels -> {out_path}")


if __name__ == "__main__":
    main()
