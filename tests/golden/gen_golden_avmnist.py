#!/usr/bin/env python3
"""Fixture of the AV-MNIST searchable fusion network by EXECUTING the unmodified reference (build container only).

    python tests/golden/gen_golden_avmnist.py          # writes tests/golden/avmnist.npz

What runs: /root/reference's own ``Searchable_Audio_Image_Net`` (models/search/avmnist_searchable.py:184-297), its
``train_sampled_models`` (:22-105) and the loop ``train_avmnist_track_acc`` (models/search/train_searchable/avmnist.py:14-85)
with ``torch.optim.Adam`` and ``LRCosineAnnealingScheduler``, on CPU, fed from this repo's synthetic AV-MNIST-shaped cache
through parameter-free stub backbones with the output structure of GP_LeNet / GP_LeNet_Deeper (models/central/avmnist.py:57,
:112: (logits, gp1, gp2, ...)).  Shims: matplotlib (absent) and the dangling ``models.aux`` import of the loop (avmnist.py:10).
"""
import argparse
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

for n in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[n] = types.ModuleType(n)
import models.auxiliary.scheduler as _sc  # noqa: E402

for n in ("models.aux", "models.train"):
    pkg = types.ModuleType(n)
    pkg.scheduler = _sc
    sys.modules[n] = pkg
    sys.modules[n + ".scheduler"] = _sc
import models.central.avmnist as central  # noqa: E402

from helpers import AVMNIST_CASE as CS, sample_tensor  # noqa: E402
from mfas_b200.avmnist_searchable import AudioImageCacheLoader, synthetic_avmnist_cache, tap_widths  # noqa: E402

AUD, IMG = tap_widths(CS["channels"])


class StubLeNet(nn.Module):            # GP_LeNet: (logits, gp1, gp2, gp3)
    def __init__(self, args, in_channels):
        super().__init__()

    def forward(self, x):
        return (None, *torch.split(x, list(IMG), 1))


class StubLeNetDeeper(nn.Module):      # GP_LeNet_Deeper: (logits, gp1 .. gp5)
    def __init__(self, args, in_channels):
        super().__init__()

    def forward(self, x):
        return (None, *torch.split(x, list(AUD), 1))


central.GP_LeNet, central.GP_LeNet_Deeper = StubLeNet, StubLeNetDeeper
import models.search.avmnist_searchable as av  # noqa: E402


def flat(prefix, d, out):
    for k, v in d.items():
        out[f"{prefix}/{k}"] = np.asarray(v)


def main():
    torch.set_num_threads(1)
    cs = CS
    tmp = tempfile.mkdtemp()
    torch.save({}, os.path.join(tmp, "aud"))
    torch.save({}, os.path.join(tmp, "rgb"))
    args = argparse.Namespace(inner_representation_size=cs["H"], num_outputs=10, channels=cs["channels"], drpt=0.0, batchnorm=False,
                              alphas=cs["alphas"], multitask=False, weightsharing=False, batchsize=cs["B"], checkpointdir=tmp, audio_cp="aud",
                              rgb_cp="rgb", eta_max=1e-3, eta_min=1e-6, Ti=cs["Ti"], Tm=2, use_dataparallel=False, verbose=False,
                              epochs=cs["epochs"])
    train = synthetic_avmnist_cache(cs["n_train"], cs["data_seed"], cs["channels"])
    dev = synthetic_avmnist_cache(cs["n_dev"], cs["data_seed"] + 1, cs["channels"])
    loaders = {"train": AudioImageCacheLoader(train, cs["B"], True, cs["loader_seed"]),
               "dev": AudioImageCacheLoader(dev, cs["B"], True, cs["loader_seed"] + 50000)}
    confs = [np.array(c) for c in cs["confs"]]
    out = {}
    # (1) initial weights + one manual step per candidate: logits, loss, autograd gradients
    torch.manual_seed(cs["model_seed"])
    E, B = cs["epochs"], cs["B"]
    for ci, conf in enumerate(confs):
        m = av.Searchable_Audio_Image_Net(args, conf)
        for k, v in m.state_dict().items():
            out[f"c{ci}/init/{k}"] = v.numpy().copy()
        m.train(True)
        rows = loaders["train"].order_for_pass(ci * E)[:B]
        logits = m((train.rgb_cat[rows], train.ske_cat[rows]))            # (image, sound), avmnist_searchable.py:207
        loss = torch.nn.CrossEntropyLoss()(logits, train.labels[rows])
        loss.backward()
        out[f"c{ci}/step0_logits"] = logits.detach().numpy()
        out[f"c{ci}/step0_loss"] = np.float32(loss.item())
        for k, p in m.named_parameters():
            if p.grad is not None:
                flat(f"c{ci}/grad/{k}", sample_tensor(p.grad.numpy()), out)
    # (2) the reference's own train_sampled_models over all candidates
    torch.manual_seed(cs["model_seed"])
    accs, models = av.train_sampled_models(confs, av.Searchable_Audio_Image_Net, loaders, args, torch.device("cpu"),
                                           return_model=list(range(len(confs))))
    for ci in range(len(confs)):
        out[f"c{ci}/best_acc"] = np.float64(float(accs[ci]))
        for k, v in models[ci].state_dict().items():
            flat(f"c{ci}/final/{k}", sample_tensor(v.numpy()), out)
    out["meta/torch"] = np.array(torch.__version__)
    path = os.path.join(os.environ.get("MFAS_GOLDEN_OUT", HERE), "avmnist.npz")
    np.savez_compressed(path, **out)
    print("avmnist ->", path, os.path.getsize(path) // 1024, "KiB; best accs", [float(a) for a in accs])


if __name__ == "__main__":
    main()
