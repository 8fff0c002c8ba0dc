#!/usr/bin/env python3
"""Markdown table of the bench lines at N = 1, 2, 4, 8 (gpurun_out/<tag>_n<N>.json or profiles/<tag>_bench_n<N>.json)."""
import json, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02y"
rows = {}
for n in (1, 2, 4, 8):
    for path in (f"profiles/{tag}_bench_n{n}.json", f"gpurun_out/{tag}_n{n}.json"):
        if os.path.exists(path):
            try:
                rows[n] = json.loads(open(path).read().strip().splitlines()[-1])
            except Exception:
                pass
            break
def f(x, d=0):
    return "–" if x is None else f"{x:.{d}f}"
print("| workload | quantity | " + " | ".join(f"N={n}" for n in rows) + " |")
print("|---|---|" + "---|" * len(rows))
print("| cfg2 (configs[1]), 148 candidates/GPU x 3 epochs, weak | value (candidate-epochs/s) | " + " | ".join(f(r["value"]) for r in rows.values()) + " |")
print("| | e2e, device init | " + " | ".join(f(r["e2e"]["value"]) for r in rows.values()) + " |")
print("| | e2e, reference-compatible host init | " + " | ".join(f((r.get("e2e_host_init") or {}).get("value")) for r in rows.values()) + " |")
print("| | fused train step, fraction of the HBM roofline | " + " | ".join(f(r["roofline"]["frac"], 3) for r in rows.values()) + " |")
for key, label in (("search256", "256-candidate search iteration (inner_repr 16, L=2, 1 epoch), strong"), ("search32", "configs[2]: 32 candidates x 5 epochs (inner_repr 16), strong"),
                   ("mmimdb64", "configs[3]: MM-IMDB, 64 candidates x 3 epochs (inner_repr 256), strong")):
    ex = [r.get("extras", {}).get(key, {}) for r in rows.values()]
    print(f"| {label} | value | " + " | ".join(f(e.get("value")) for e in ex) + " |")
    print("| | e2e | " + " | ".join(f((e.get("e2e") or {}).get("value")) for e in ex) + " |")
    print("| | per-GPU fraction of the HBM roofline (value) | " + " | ".join(f(e.get("frac"), 3) for e in ex) + " |")
    v1 = ex[0].get("value")
    print("| | speed-up over N=1 (value / e2e) | " + " | ".join((f(e.get("value") / v1, 2) + " / " + f(e["e2e"]["value"] / ex[0]["e2e"]["value"], 2)) if e.get("value") else "–" for e in ex) + " |")
dp = [r.get("extras", {}).get("depth", {}) for r in rows.values()]
print("| configs[4]: depth sweep L 1..6 x inner_repr {64,128,256}, bs 128 (18 points sharded over the ranks) | roofline fraction min / mean / max | " +
      " | ".join((f(d.get("frac_min"), 2) + " / " + f(d.get("frac_mean"), 2) + " / " + f(d.get("frac_max"), 2)) if d else "–" for d in dp) + " |")
