#!/usr/bin/env python3
"""Diagnostic (GPU): per-step train loss of the cfg2-shaped trajectory -- CUDA path free-running and teacher-forced -- against the oracle."""
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
conf = FOUND_CONFS[4]; H, B, E, ntr, ndv = 128, 64, 3, 448, 192
train, dev = synthetic_ntu_cache(ntr, 5), synthetic_ntu_cache(ndv, 6)
trs = split_np(train)
ltr = FeatureCacheLoader(train, B, True, 7)
init = init_states([conf], H, 60, True, 0.0, 1)[0]
steps = ntr // B
def oracle(dt):
    out = []
    with O.precision(dt):
        head = O.FusionHead(conf, H, 60, init); sch = O.CosineRestartLR(1e-3, 1e-6, 1, 2, ntr / B)
        states = []
        for e in range(E):
            order = ltr.order_for_pass(e).numpy()
            for s in range(steps):
                rows = order[s * B:(s + 1) * B]
                lr = sch.step()
                states.append((dict(state={k: np.array(v, np.float32) for k, v in head.state.items()},
                                    adam={k: (np.array(m, np.float32), np.array(v, np.float32)) for k, (m, v) in head.adam.items()}), head.t, lr, rows))
                sk, rg, y = O._taps_of(trs, rows)
                _, loss, _ = head.train_step(sk, rg, y, lr)
                out.append(float(loss))
    return np.array(out), states
L32, states = oracle(np.float32)
L64, _ = oracle(np.float64)
tc = train.to(DEV)
def mk():
    g = CandidateGroup([conf], H, 60, _lib.FLAG_BN, DEV, batch_max=B); g.set_adam(0.9, 0.999, 1e-8, 1e-4); return g
g = mk(); g.load_state(0, init)
Lg = []
for (st, t, lr, rows) in states:
    _, loss, _ = g.train_step(tc, torch.from_numpy(rows), lr=lr); Lg.append(float(loss[0]))
Lg = np.array(Lg)
g2 = mk(); Ltf = []
for (st, t, lr, rows) in states:
    g2.load_state(0, st["state"])
    for k, (m, v) in st["adam"].items():
        g2.view(0, k, "m").copy_(torch.from_numpy(m)); g2.view(0, k, "v").copy_(torch.from_numpy(v))
    g2.adam_t = t
    _, loss, _ = g2.train_step(tc, torch.from_numpy(rows), lr=lr); Ltf.append(float(loss[0]))
Ltf = np.array(Ltf)
os.environ["MFAS_ENGINE"] = "ffma"
g3 = mk(); g3.load_state(0, init); os.environ.pop("MFAS_ENGINE")
Lf = []
for (st, t, lr, rows) in states:
    _, loss, _ = g3.train_step(tc, torch.from_numpy(rows), lr=lr); Lf.append(float(loss[0]))
Lf = np.array(Lf)
np.set_printoptions(precision=3, linewidth=200)
print("engine", g.engine, g3.engine)
print("step  L64        (L32-L64)/L64  (Ltc-L64)/L64  (Ltc_teacher_forced-L32)/L32  (Lffma-L64)/L64")
for t in range(len(L64)):
    print(f"{t:3d}  {L64[t]:.6f}  {(L32[t]-L64[t])/L64[t]:+.2e}  {(Lg[t]-L64[t])/L64[t]:+.2e}  {(Ltf[t]-L32[t])/L32[t]:+.2e}  {(Lf[t]-L64[t])/L64[t]:+.2e}")
for e in range(E):
    sl = slice(e * steps, (e + 1) * steps)
    print(f"epoch {e}: sum rel: fp32 {abs(L32[sl].sum()-L64[sl].sum())/L64[sl].sum():.2e}  tc {abs(Lg[sl].sum()-L64[sl].sum())/L64[sl].sum():.2e}  ffma {abs(Lf[sl].sum()-L64[sl].sum())/L64[sl].sum():.2e}")
