"""Diagnostic (run on a GPU box): forward error of the tensor-core engines against the float64 oracle.
    python tests/cuda/fwd_error.py
Prints max-norm relative errors of the logits and of every gradient for MFAS_FWD=cta and the default."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import init_states, split_np
from mfas_b200 import _lib
from mfas_b200.cache import synthetic_ntu_cache
from mfas_b200.engine import CandidateGroup
from oracle import mfas_oracle as O

conf = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]]
H, B = 128, 64
train = synthetic_ntu_cache(256, 3)
for seed in (0, 1, 2):
    init = init_states([conf], H, 60, True, 0.0, seed)[0]
    rows = torch.arange(B) + 64 * (seed % 3)
    sk, rg, y = O._taps_of(split_np(train), rows.numpy())
    with O.precision(np.float64):
        h64 = O.FusionHead(conf, H, 60, init)
        l64, tape = h64.forward(sk, rg, train=True)
        g64 = h64.backward(l64, y, tape)
    h32 = O.FusionHead(conf, H, 60, init)
    l32, tape32 = h32.forward(sk, rg, train=True)
    print(f"seed {seed}: oracle fp32 logits err {np.abs(l32 - l64).max() / np.abs(l64).max():.2e}")
    for mode in ("cta", "ws"):
        os.environ["MFAS_FWD"] = mode
        g = CandidateGroup([conf], H, 60, _lib.FLAG_BN, "cuda:0", batch_max=B, keep_grads=True)
        g.set_adam(0.9, 0.999, 1e-8, 1e-4)
        g.load_state(0, init)
        logits, loss, _ = g.train_step(train.to("cuda:0"), rows, lr=1e-3)
        torch.cuda.synchronize()
        lg = logits[0].cpu().numpy()
        gg = g.state(0, "g")
        gerr = {k: np.abs(gg[k] - v).max() / max(np.abs(v).max(), 1e-12) for k, v in g64.items()}
        worst = max(gerr, key=gerr.get)
        print(f"  MFAS_FWD={mode}: logits err {np.abs(lg - l64).max() / np.abs(l64).max():.2e}  worst grad {worst} {gerr[worst]:.2e}"
              f"  W0 grad {gerr['fusion_layers.0.0.weight']:.2e}")
        g.close()
