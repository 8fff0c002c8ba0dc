#!/usr/bin/env python3
"""Latency floor of the fused train / eval step at small candidate counts and small inner_repr (the shapes the SMBO search
issues, BASELINE configs[2] and north_star's 256-candidate iteration sharded over 8 GPUs): per-kernel CUDA-event times."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mfas_b200 import _lib
from mfas_b200.cache import synthetic_ntu_cache
from mfas_b200.engine import CandidateGroup, algorithmic_counts
dev = torch.device("cuda:0")
cache = synthetic_ntu_cache(4096, 1).to(dev)
rows32 = [[i, j, k] for i in range(4) for j in range(4) for k in range(2)]
parents = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0], [2, 2, 0], [0, 0, 1], [3, 2, 0], [2, 3, 1]]
cases = [("search32/8gpu", [np.array([r]) for r in rows32][:4], 16), ("search32/1gpu", [np.array([r]) for r in rows32], 16),
         ("search256/8gpu", [np.array([p, r]) for p in parents for r in rows32][:32], 16), ("search256/2gpu", [np.array([p, r]) for p in parents for r in rows32][:128], 16),
         ("search256/1gpu", [np.array([p, r]) for p in parents for r in rows32], 16),
         ("cfg2 x4", [np.array([[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]])] * 4, 128), ("cfg2 x148", [np.array([[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]])] * 148, 128)]
B = 64
for name, confs, H in cases:
    g = CandidateGroup(confs, H, 60, _lib.FLAG_BN, dev, batch_max=B)
    g.set_adam(0.9, 0.999, 1e-8, 1e-4); g.params.uniform_(-0.03, 0.03); g.bufs.fill_(1.0)
    n = len(confs)
    gen = torch.Generator(device=dev).manual_seed(0)
    rws = [torch.randint(0, 4096, (n, B), device=dev, generator=gen, dtype=torch.int32) for _ in range(45)]
    for i in range(5): g.train_step(cache, rws[i], lr=1e-3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(40): g.train_step(cache, rws[5 + i], lr=1e-3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    g.set_profiling(True); acc = [0, 0, 0]
    for i in range(20):
        g.train_step(cache, rws[5 + i], lr=1e-3); acc = [x + y for x, y in zip(acc, g.last_step_ms())]
    g.set_profiling(False)
    perm = torch.stack([torch.randperm(4096, device=dev) for _ in range(n)]).to(torch.int32)
    g.eval_pass(cache, B, perm); torch.cuda.synchronize(); e0.record()
    for _ in range(3): g.eval_pass(cache, B, perm)
    e1.record(); torch.cuda.synchronize()
    ev = e0.elapsed_time(e1) / 3 / 32        # 4096 rows in 128-row steps
    byt = sum(algorithmic_counts(l, B)["train_bytes"] for l in g.layouts)
    print(json.dumps({"case": name, "n": n, "H": H, "train_step_us": ms * 1e3, "kernels_us": [round(x / 20 * 1e3, 1) for x in acc], "eval_step128_us": ev * 1e3,
                      "train_MB": byt / 1e6, "frac": byt / ms / 1e6 / 6550.1}))
    g.close()

# MM-IMDB configs[3] step (64 two-step candidates, inner_repr 256, bs 64): per-kernel times
import mfas_b200.mmimdb_searchable as mm
mcache = mm.synthetic_mmimdb_cache(4096, 1).to(dev)
rows_mm = mm.get_possible_layer_configurations(0)
rng = np.random.default_rng(0)
for n in (64, 8):
    confs = [np.array([rows_mm[i] for i in rng.integers(0, len(rows_mm), size=2)]) for _ in range(n)]
    g = CandidateGroup(confs, 256, 23, _lib.FLAG_BN | _lib.FLAG_MULTILABEL, dev, batch_max=64, widths=mm.WIDTHS)
    g.set_adam(0.9, 0.999, 1e-8, 1e-4); g.params.uniform_(-0.03, 0.03); g.bufs.fill_(1.0)
    gen = torch.Generator(device=dev).manual_seed(0)
    rws = [torch.randint(0, 4096, (n, 64), device=dev, generator=gen, dtype=torch.int32) for _ in range(45)]
    for i in range(5): g.train_step(mcache, rws[i], lr=1e-3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(40): g.train_step(mcache, rws[5 + i], lr=1e-3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    g.set_profiling(True); acc = [0, 0, 0]
    for i in range(20):
        g.train_step(mcache, rws[5 + i], lr=1e-3); acc = [x + y for x, y in zip(acc, g.last_step_ms())]
    byt = sum(algorithmic_counts(l, 64)["train_bytes"] for l in g.layouts)
    print(json.dumps({"case": f"mmimdb64/{64 // n}gpu", "n": n, "H": 256, "train_step_us": ms * 1e3, "kernels_us": [round(x / 20 * 1e3, 1) for x in acc], "eval_step128_us": 0.0,
                      "train_MB": byt / 1e6, "frac": byt / ms / 1e6 / 6550.1}))
    g.close()
