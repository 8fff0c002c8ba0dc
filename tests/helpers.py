"""Shared test helpers (no reference import here -- this file travels to the GPU box)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# /root/reference/main_found_ntu.py:173-182
FOUND_CONFS = {
    0: [[2, 2, 0], [1, 0, 1], [3, 2, 0], [3, 1, 1]],
    1: [[3, 0, 0], [1, 3, 0], [1, 1, 1], [3, 3, 0]],
    2: [[3, 2, 0], [2, 3, 1], [0, 1, 1], [3, 0, 0]],
    3: [[1, 1, 1], [3, 2, 0], [0, 1, 1], [3, 0, 0]],
    4: [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]],
}

# Golden cases: name -> settings.  Sizes are tiny so the CPU suite stays fast; partial last
# batches are on purpose (drop_last=False in the reference).
GOLDEN_CASES = {
    "cfg1": dict(confs=[FOUND_CONFS[0]], H=32, B=8, n_train=60, n_dev=28, epochs=2, bn=True, drpt=0.0,
                 Ti=5, model_seed=0, data_seed=11),
    "cfg2": dict(confs=[FOUND_CONFS[4]], H=128, B=64, n_train=288, n_dev=160, epochs=3, bn=True, drpt=0.0,
                 Ti=1, model_seed=0, data_seed=21),
    "mixL": dict(confs=[[[0, 0, 0]], [[3, 1, 1], [2, 2, 2]], [[1, 3, 2], [0, 0, 1], [2, 1, 0]]],
                 H=16, B=32, n_train=128, n_dev=64, epochs=2, bn=True, drpt=0.0, Ti=1, model_seed=3,
                 data_seed=31),
    # modality gates (AlphaScalarMultiplication, aux_models.py:94-111): args.alphas=True, mixed depths
    "alph": dict(confs=[[[3, 1, 1], [1, 3, 0]], [[0, 2, 1]], [[2, 0, 0], [1, 1, 1], [3, 2, 2]]], H=32, B=16, n_train=80,
                 n_dev=40, epochs=2, bn=True, drpt=0.0, Ti=1, model_seed=5, data_seed=41, alphas=True),
}


# weight sharing (args.weightsharing, /root/reference/models/search/ntu_searchable.py:74-75,91-92,123-174): candidates are trained one
# after the other and chained through a dict keyed '<step>.L_<in>_<out>.A_<act>'.  c0 / c1 / c3 share step 0 ('0.L_1536_32.A_sigmoid'),
# c0 / c2 share step 1 ('1.L_2336_32.A_relu': 256 + 2048 + 32 both ways), c3 also shares its step 1 with c1.
WS_CASES = {
    "wsh": dict(confs=[[[3, 1, 1], [1, 3, 0]], [[3, 1, 1], [2, 2, 0]], [[0, 2, 0], [1, 3, 0]], [[3, 1, 1], [2, 2, 0], [0, 0, 1]]],
                H=32, B=16, n_train=112, n_dev=120, epochs=2, bn=True, drpt=0.0, Ti=1, model_seed=6, data_seed=71, weightsharing=True),
}

# AV-MNIST searchable fusion: the reference's own class + loop on cached taps, tests/golden/gen_golden_avmnist.py
AVMNIST_CASE = dict(confs=[[[4, 2, 0], [1, 0, 1]], [[0, 1, 1]], [[3, 2, 2], [2, 1, 0], [4, 0, 1]]], H=32, B=16, channels=32, n_train=112, n_dev=120,
                    epochs=2, Ti=1, alphas=True, model_seed=8, data_seed=81, loader_seed=500)

# main_found_ntu.py flow (multitask + alphas, two training stages, test pass): tests/golden/gen_golden_found.py
FOUND_MT_CASE = dict(conf=FOUND_CONFS[3][:3], H=32, B=16, n_train=96, n_dev=48, n_test=40, epochs=2, Ti=1, alphas=True,
                     model_seed=9, data_seed=51, loader_seed=300)


def make_args(H, B, epochs, bn=True, drpt=0.0, Ti=1, Tm=2, eta_max=1e-3, eta_min=1e-6, C=60,
              alphas=False, multitask=False, weightsharing=False, checkpointdir="", verbose=False):
    """Namespace with every field the hot path reads ('Config / flags')."""
    return argparse.Namespace(
        inner_representation_size=H, num_outputs=C, vid_len=(8, 32), drpt=drpt, batchnorm=bn,
        alphas=alphas, multitask=multitask, weightsharing=weightsharing, batchsize=B,
        checkpointdir=checkpointdir, ske_cp="ske", rgb_cp="rgb", eta_max=eta_max, eta_min=eta_min,
        Ti=Ti, Tm=Tm, use_dataparallel=False, verbose=verbose, epochs=epochs)


D_SKE = (128, 256, 1024, 512)
D_RGB = (512, 1024, 2048, 2048)


# MM-IMDB searchable fusion: the reference loop around a reference-style module, tests/golden/gen_golden_mmimdb_path.py
MMIMDB_CASE = dict(confs=[[[1, 2, 0], [0, 3, 1], [1, 0, 0]], [[0, 1, 1]]], H=64, B=32, n_train=320, n_dev=72, epochs=4, Ti=1, eta_max=1e-2,
                   model_seed=7, data_seed=61, loader_seed=400)
D_TEXT = (64, 128)
D_IMAGE = (512, 512, 512, 512)


def make_mmimdb_args(H, B, epochs, bn=True, drpt=0.0, Ti=1, Tm=2, eta_max=1e-3, eta_min=1e-6, C=23, alphas=False, verbose=False):
    return argparse.Namespace(inner_representation_size=H, num_outputs=C, drpt=drpt, batchnorm=bn, alphas=alphas,
                              multitask=False, weightsharing=False, batchsize=B, eta_max=eta_max, eta_min=eta_min, Ti=Ti, Tm=Tm,
                              use_dataparallel=False, verbose=verbose, epochs=epochs)


def split_np_mmimdb(cache):
    """text/image FeatureCache -> dict of numpy tap arrays for oracle/mmimdb_oracle.py."""
    return dict(text=[t.numpy() for t in cache.ske_taps()], image=[t.numpy() for t in cache.rgb_taps()],
                targets=cache.labels.numpy(), pos_weight=cache.pos_weight.numpy())


def init_states(confs, H, C, bn, drpt, seed, widths=None):
    """Initial state_dicts of the fusion heads for ``confs`` built back to back from one seed.

    Replays the reference constructor's RNG consumption with parameter-free backbones
    (/root/reference/models/search/ntu_searchable.py:179-204): per candidate, L x nn.Linear
    (+BatchNorm1d), the classifier nn.Linear, then normal_(alpha, 0, 0.1) per step.
    ``tests/golden/gen_golden.py`` asserts this equals the reference model's own init.
    """
    torch.manual_seed(seed)
    d0, d1 = (D_SKE, D_RGB) if widths is None else widths
    out = []
    for conf in confs:
        sd = {}
        for l, c in enumerate(conf):
            K = d0[c[0]] + d1[c[1]] + (H if l > 0 else 0)
            lin = nn.Linear(K, H)
            sd[f"fusion_layers.{l}.0.weight"] = lin.weight.detach().numpy().copy()
            sd[f"fusion_layers.{l}.0.bias"] = lin.bias.detach().numpy().copy()
            if bn:
                sd[f"fusion_layers.{l}.2.weight"] = np.ones(H, np.float32)
                sd[f"fusion_layers.{l}.2.bias"] = np.zeros(H, np.float32)
                sd[f"fusion_layers.{l}.2.running_mean"] = np.zeros(H, np.float32)
                sd[f"fusion_layers.{l}.2.running_var"] = np.ones(H, np.float32)
                sd[f"fusion_layers.{l}.2.num_batches_tracked"] = np.zeros((), np.int64)
        cls = nn.Linear(H, C)
        sd["central_classifier.weight"] = cls.weight.detach().numpy().copy()
        sd["central_classifier.bias"] = cls.bias.detach().numpy().copy()
        for l in range(len(conf)):
            a = torch.zeros(1)
            nn.init.normal_(a, 0.0, 0.1)
            sd[f"alphas.{l}.alpha_x"] = a.numpy().copy()
        out.append(sd)
    return out


def sample_tensor(t, stride=7):
    """Deterministic thin sample of a tensor + its fp64 sum / sum of squares."""
    a = np.asarray(t, dtype=np.float64).ravel()
    return dict(sample=a[::stride].astype(np.float32), s1=a.sum(), s2=(a * a).sum(), n=a.size,
                amax=np.abs(a).max() if a.size else 0.0)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def split_np(cache):
    """FeatureCache -> dict of numpy tap arrays for the oracle."""
    d = dict(ske=[t.numpy() for t in cache.ske_taps()], rgb=[t.numpy() for t in cache.rgb_taps()],
             labels=cache.labels.numpy())
    if cache.logit_rgb is not None:
        d.update(logit_rgb=cache.logit_rgb.numpy(), logit_ske=cache.logit_ske.numpy())
    return d


def report_traj(what, err, tol):
    """Print an achieved trajectory error with its bound (pytest -rP / -s), append it to gpurun_out/traj_errors.txt when that
    directory exists, and assert it."""
    line = f"TRAJ {what}: achieved {err:.3e} (bound {tol:.1e})"
    print(line)
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        with open(os.path.join(log, "traj_errors.txt"), "a") as f:
            f.write(line + "\n")
    assert err <= tol, line
