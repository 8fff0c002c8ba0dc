#!/usr/bin/env python3
"""BASELINE.json configs[4]: NTU fusion-depth sweep L in 1..6 x inner_repr in {64, 128, 256}, bs = 128 -- roofline report.

For every (L, H): rows `conf4[l mod 4]` (SURVEY.md 8(d) "cfg5"), algorithmic bytes / flops of one train step
(`mfas_algorithmic_counts`), and -- on a GPU -- the CUDA-event time of a train step of M candidates, the achieved
algorithmic GB/s and its fraction of the measured HBM peak.

    python profiles/depth_sweep.py --dry            # host only: the algorithmic table (no GPU needed)
    python profiles/depth_sweep.py [--cands 64]     # one JSON line per (L, H) with the measured columns
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mfas_b200 import _lib  # noqa: E402
from mfas_b200.engine import algorithmic_counts, plan_layout  # noqa: E402

CONF4 = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]]            # /root/reference/main_found_ntu.py:181-182
B = 128


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dry", action="store_true")
    ap.add_argument("--cands", type=int, default=148, help="candidates per GPU (148 = one fused-chain CTA per SM)")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    peak = 6546.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if not a.dry:
        import torch
        from mfas_b200.cache import synthetic_ntu_cache
        from mfas_b200.engine import CandidateGroup
        dev = torch.device("cuda:0")
        cache = synthetic_ntu_cache(4096, 1).to(dev)
    for H in (64, 128, 256):
        for L in range(1, 7):
            conf = np.array([CONF4[l % 4] for l in range(L)])
            cnt = algorithmic_counts(plan_layout(conf, H, 60, _lib.FLAG_BN), B)
            row = {"L": L, "inner_repr": H, "batch": B, "train_MB_per_step": cnt["train_bytes"] / 1e6,
                   "eval_MB_per_step": cnt["eval_bytes"] / 1e6, "MFLOP_per_step": (cnt["fwd_flops"] + cnt["bwd_flops"]) / 1e6,
                   "flop_per_byte": (cnt["fwd_flops"] + cnt["bwd_flops"]) / cnt["train_bytes"]}
            if not a.dry:
                g = CandidateGroup([conf] * a.cands, H, 60, _lib.FLAG_BN, dev, batch_max=B)
                g.set_adam(0.9, 0.999, 1e-8, 1e-4)
                g.params.uniform_(-0.03, 0.03)
                g.bufs.fill_(1.0)
                gen = torch.Generator(device=dev).manual_seed(0)
                rows = [torch.randint(0, 4096, (a.cands, B), device=dev, generator=gen, dtype=torch.int32) for _ in range(a.steps + 5)]   # every candidate its own batch
                for i in range(5):
                    g.train_step(cache, rows[i], lr=1e-3)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for i in range(a.steps):
                    g.train_step(cache, rows[5 + i], lr=1e-3)
                e1.record()
                torch.cuda.synchronize()
                g.check()
                ms = e0.elapsed_time(e1) / a.steps
                gbs = a.cands * cnt["train_bytes"] / ms / 1e6
                row.update(engine=g.engine, candidates=a.cands, ms_per_step=ms, achieved_GBs=gbs, peak_GBs=peak, frac=gbs / peak)
                try:                                            # the three kernels of the step, CUDA events on the launching stream
                    g.set_profiling(True)
                    acc = [0.0, 0.0, 0.0]
                    for i in range(10):
                        g.train_step(cache, rows[5 + i], lr=1e-3)
                        acc = [x + y for x, y in zip(acc, g.last_step_ms())]
                    row["kernel_ms"] = dict(zip(("k_tc_fwd_ws", "k_chain_all", "k_tc_bwd_ws"), [x / 10 for x in acc]))
                except Exception as ex:
                    row["kernel_ms"] = str(ex)
                g.close()
            print(json.dumps(row))


if __name__ == "__main__":
    main()
