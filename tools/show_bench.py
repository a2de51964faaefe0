"""Pretty-print the last JSON line of a bench.py log."""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "n_gpus")}, "e2e", d["e2e"]["value"], "infer", d.get("inference", {}).get("value"))
print(d["roofline"])
print("cpu", d.get("cpu_baseline"))
tot = 0
for k in d.get("kernels", []):
    tot += k["ms"]
    print(f"{k['call']:42s} {k['ms']:.4f} {k['share_of_step']:.3f} {k['achieved']:9.1f} {k['unit']:8s} frac={k['frac']:.3f}")
print("sum of libmgs calls ms:", round(tot, 3))
