#!/usr/bin/env python3
"""Diagnostic (GPU): per-tensor gradient error against float64 ground truth at teacher-forced states along the cfg2-shaped
trajectory, for the tensor-core engine, the CUDA-core engine and the fp32 oracle (rel L2)."""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import FOUND_CONFS, init_states, split_np
from mfas_b200 import _lib
from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache
from mfas_b200.engine import CandidateGroup
from oracle import mfas_oracle as O
DEV = "cuda:0"
conf = FOUND_CONFS[4]; H, B, E, ntr = 128, 64, 3, 448
train = synthetic_ntu_cache(ntr, 5); trs = split_np(train); tc = train.to(DEV)
ltr = FeatureCacheLoader(train, B, True, 7)
init = init_states([conf], H, 60, True, 0.0, 1)[0]
steps = ntr // B
def rl2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
def mk(engine):
    os.environ["MFAS_ENGINE"] = engine
    g = CandidateGroup([conf], H, 60, _lib.FLAG_BN, DEV, batch_max=B, keep_grads=True); g.set_adam(0.9, 0.999, 1e-8, 1e-4)
    os.environ.pop("MFAS_ENGINE"); return g
gt, gf = mk("tc"), mk("ffma")
head = O.FusionHead(conf, H, 60, init); sch = O.CosineRestartLR(1e-3, 1e-6, 1, 2, ntr / B)
names = None
for e in range(E):
    order = ltr.order_for_pass(e).numpy()
    for s in range(steps):
        t = e * steps + s
        rows = order[s * B:(s + 1) * B]; lr = sch.step()
        st = {k: np.array(v, np.float32) for k, v in head.state.items()}
        ad = {k: (np.array(m, np.float32), np.array(v, np.float32)) for k, (m, v) in head.adam.items()}
        t0 = head.t
        sk, rg, y = O._taps_of(trs, rows)
        with O.precision(np.float64):
            h64 = O.FusionHead(conf, H, 60, st); lg, tape = h64.forward(sk, rg, train=True); g64 = h64.backward(lg, y, tape)
        _, _, g32 = head.train_step(sk, rg, y, lr)
        if t not in (0, 3, 7, 10, 12, 14, 16, 18, 20):
            continue
        res = {}
        for nm, g in (("tc", gt), ("ffma", gf)):
            g.load_state(0, st)
            for k, (m, v) in ad.items():
                g.view(0, k, "m").copy_(torch.from_numpy(m)); g.view(0, k, "v").copy_(torch.from_numpy(v))
            g.adam_t = t0
            g.train_step(tc, torch.from_numpy(rows), lr=lr)
            res[nm] = g.state(0, "g")
        print(f"--- step {t} (lr {lr:.2e}); rel-L2 gradient error vs float64: oracle32 | tc | ffma")
        for k in g32:
            if k.startswith("alphas"): continue
            print(f"  {k:34s} {rl2(g32[k], g64[k]):.2e} | {rl2(res['tc'][k], g64[k]):.2e} | {rl2(res['ffma'][k], g64[k]):.2e}")
