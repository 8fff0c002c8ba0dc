#!/usr/bin/env python3
"""Roofline of the cache builder's pooling kernel (k_global_pool): the NTU visual tap out_4 of one 64-clip batch
([64, 2048, 8, 7, 7] fp32 = 205 MB, larger than the 126 MB L2) pooled into its cache slice; CUDA events, 3 warm-up + 20
timed launches.  Algorithmic bytes per launch = 4 (S + 1) per (b, c) row.  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mfas_b200 import cache_builder as cb  # noqa: E402


def main():
    out = {}
    for name, shape in (("out_4 [64,2048,8,7,7]", (64, 2048, 8, 7, 7)), ("out_2 [64,512,8,28,28]", (64, 512, 8, 28, 28))):
        x = torch.rand(*shape, device="cuda:0")
        dst = torch.empty(shape[0], 5632, device="cuda:0")[:, 1536:1536 + shape[1]]
        for _ in range(3):
            cb.global_pool_into(x, dst)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            cb.global_pool_into(x, dst)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        rows, S = shape[0] * shape[1], x.numel() // (shape[0] * shape[1])
        out[name] = {"ms": ms, "GB/s": 4.0 * rows * (S + 1) / ms / 1e6}
    peak = 6546.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({"kernel": "k_global_pool", "bound": "hbm", "peak_GB/s": peak,
                      "taps": {k: dict(v, frac=v["GB/s"] / peak) for k, v in out.items()}}))


if __name__ == "__main__":
    main()
