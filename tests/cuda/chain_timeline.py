"""Diagnostic (GPU box): phase timeline of the fused chain kernel, 128 cfg2 candidates, in SM clocks."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["MFAS_CHAIN_TIMELINE"] = "1"
import numpy as np, torch
from mfas_b200 import _lib
from mfas_b200.cache import synthetic_ntu_cache
from mfas_b200.engine import CandidateGroup
conf = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]][:int(os.environ.get("TL_L", "4"))]
M, H, B = int(os.environ.get("TL_M", "128")), int(os.environ.get("TL_H", "128")), 64
train = synthetic_ntu_cache(4096, 3).to("cuda:0")
g = CandidateGroup([conf] * M, H, 60, _lib.FLAG_BN, "cuda:0", batch_max=B)
g.set_adam(0.9, 0.999, 1e-8, 1e-4)
for c in range(M):
    for name in g.names(c):
        v = g.view(c, name)
        if name.endswith("0.weight") or name.startswith("central"):
            v.uniform_(-0.02, 0.02)
        elif name.endswith("2.weight") or name.endswith("running_var"):
            v.fill_(1.0)
rows = torch.stack([torch.randperm(4096)[:B] for _ in range(M)]).to("cuda:0", torch.int32)
for it in range(6):
    g.train_step(train, rows, 1e-4)
out = np.zeros((M, 16), np.int64)
_lib.check(_lib.lib().mfas_group_chain_timeline(g._h, out.ctypes.data, M))
L = len(conf)
d = np.diff(out[:, :2 * L + 2], axis=1)
names = ["open+fwd0"] + [f"fwd{l}" for l in range(1, L)] + ["head"] + [f"bwd{l}" for l in range(L - 1, -1, -1)]
print("phase: median / max cycles over candidates (1 us ~ 1900 cycles)")
for i, n in enumerate(names):
    print(f"  {n:10s} {np.median(d[:, i]):9.0f} {d[:, i].max():9.0f}")
print("  total      %9.0f %9.0f" % (np.median(out[:, 2 * L + 1] - out[:, 0]), (out[:, 2 * L + 1] - out[:, 0]).max()))
inner = out[:, 10:16]
t0 = inner[:, 0]
if H <= 32 and os.environ.get("MFAS_CHAIN_SMALL", "1") != "0":      # k_chain_small stamps (cycles since the kernel was entered)
    k0 = out[:, 0]
    print("k_chain_small, cycles since kernel entry (median): forward stream complete (griddepcontrol.wait) %d | weights, vectors, partial sums staged %d | fwd layers done %d | logits %d | head rows (softmax-CE, dlogits) %d | head done %d | end %d" % (
        np.median(inner[:, 0] - k0), np.median(inner[:, 1] - k0), np.median(out[:, L] - k0), np.median(inner[:, 2] - k0), np.median(inner[:, 3] - k0),
        np.median(out[:, L + 1] - k0), np.median(out[:, 2 * L + 1] - k0)))
    if L > 1:
        print("  inside forward layer 1 (cycles since the layer was entered, median): partial sums + W_hid h + bias + activation %d | BatchNorm statistics %d | layer end %d" % (
            np.median(inner[:, 4] - out[:, 1]), np.median(inner[:, 5] - out[:, 1]), np.median(out[:, 2] - out[:, 1])))
        if L == 2:      # (slots 6..8 are free at L = 2: inside the BatchNorm statistics)
            print("    BatchNorm statistics: first column sum %d | squares %d | second column sum %d | invstd, running stats %d" % (
                np.median(out[:, 6] - inner[:, 4]), np.median(out[:, 7] - out[:, 6]), np.median(out[:, 8] - out[:, 7]), np.median(inner[:, 5] - out[:, 8])))
    sys.exit(0)
print("inside forward layer 1 (cycles since the layer was entered, median): raw operand tiles landed (cp.async) %d | lo tiles written %d | MMAs issued %d | z complete (partials, MMA, bias, act) %d | BN stats %d | (layer end %d)" % (
    np.median(inner[:, 1] - t0), np.median(inner[:, 2] - t0), np.median(inner[:, 3] - t0), np.median(inner[:, 4] - t0),
    np.median(inner[:, 5] - t0), np.median(out[:, 2] - t0)))
