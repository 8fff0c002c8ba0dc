#!/usr/bin/env python3
"""Fixture for the MM-IMDB searchable-fusion path, produced by EXECUTING the reference's own
training loop ``train_mmimdb_track_f1`` (/root/reference/models/search/train_searchable/mmimdb.py:14-136, with its sklearn
``f1_score``), loss ``WeightedCrossEntropyWithLogits`` (models/auxiliary/aux_models.py:129-147), ``torch.optim.Adam`` and
``LRCosineAnnealingScheduler`` (models/auxiliary/scheduler.py:12-46) around a torch module written in the reference's
style.  The reference has no MM-IMDB *searchable* network; ``RefStyleTextImageNet`` below is
``Searchable_Skeleton_Image_Net`` (models/search/ntu_searchable.py:178-301) with the MM-IMDB tap widths and the
``model(text, image)`` call of the MM-IMDB loop -- plain nn.Linear / BatchNorm1d / autograd, nothing of this repo.

    python tests/golden/gen_golden_mmimdb_path.py        # writes tests/golden/mmimdb_path.npz
"""
import contextlib
import io
import os
import re
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.optim as op

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("MFAS_GOLDEN_OUT", HERE)          # tests regenerate into a scratch directory and compare
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")
for n in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[n] = types.ModuleType(n)
import models.auxiliary.scheduler as _sc  # noqa: E402  (real reference module)

for n in ("models.aux", "models.train"):
    pkg = types.ModuleType(n)
    pkg.scheduler = _sc
    sys.modules[n] = pkg
    sys.modules[n + ".scheduler"] = _sc
from models.auxiliary.aux_models import WeightedCrossEntropyWithLogits  # noqa: E402
from models.search.train_searchable.mmimdb import train_mmimdb_track_f1  # noqa: E402

from helpers import D_IMAGE, D_TEXT, MMIMDB_CASE, init_states, make_mmimdb_args, sample_tensor  # noqa: E402
from mfas_b200.mmimdb_searchable import TextImageCacheLoader, synthetic_mmimdb_cache  # noqa: E402


class RefStyleTextImageNet(nn.Module):
    """ntu_searchable.py:178-301 with text / image taps (no alphas, batchnorm recipe)."""

    def __init__(self, args, conf):
        super().__init__()
        self.conf, self.args = conf, args
        H = args.inner_representation_size
        layers = []
        for i, c in enumerate(conf):
            in_size = D_TEXT[c[0]] + D_IMAGE[c[1]] + (H if i > 0 else 0)
            nl = [nn.ReLU(), nn.Sigmoid(), nn.LeakyReLU()][c[2]]
            layers.append(nn.Sequential(nn.Linear(in_size, H), nl, nn.BatchNorm1d(H)))
        self.fusion_layers = nn.ModuleList(layers)
        self.central_classifier = nn.Linear(H, args.num_outputs)

    def forward(self, text, image):
        tt, it = torch.split(text, D_TEXT, 1), torch.split(image, D_IMAGE, 1)
        for l, c in enumerate(self.conf):
            parts = (tt[c[0]], it[c[1]]) if l == 0 else (tt[c[0]], it[c[1]], out)
            out = self.fusion_layers[l](torch.cat(parts, 1))
        return self.central_classifier(out)


def main():
    torch.set_num_threads(1)
    cs = MMIMDB_CASE
    H, B, E = cs["H"], cs["B"], cs["epochs"]
    args = make_mmimdb_args(H, B, E, Ti=cs["Ti"], eta_max=cs["eta_max"])
    train = synthetic_mmimdb_cache(cs["n_train"], cs["data_seed"])
    dev = synthetic_mmimdb_cache(cs["n_dev"], cs["data_seed"] + 1)
    out = {"pos_weight": train.pos_weight.numpy()}
    inits = init_states(cs["confs"], H, 23, True, 0.0, cs["model_seed"], widths=(D_TEXT, D_IMAGE))
    for ci, conf in enumerate(cs["confs"]):
        def fresh():
            m = RefStyleTextImageNet(args, conf)
            m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in inits[ci].items() if not k.startswith("alphas")})
            return m
        loaders = {"train": TextImageCacheLoader(train, B, True, cs["loader_seed"] + ci),
                   "dev": TextImageCacheLoader(dev, B, True, cs["loader_seed"] + 50000 + ci)}
        crit = WeightedCrossEntropyWithLogits(train.pos_weight.numpy())
        # ---- one manual step: logits, loss, autograd gradients
        m = fresh()
        m.train(True)
        rows = loaders["train"].order_for_pass(0)[:B]
        logits = m(train.ske_cat[rows], train.rgb_cat[rows])
        loss = crit(logits, train.labels[rows])
        loss.backward()
        out[f"c{ci}/step/logits"] = logits.detach().numpy()
        out[f"c{ci}/step/loss"] = np.float32(loss.item())
        for k, p in m.named_parameters():
            out[f"c{ci}/step/grad/{k}"] = p.grad.numpy().copy()
        # ---- the reference loop
        m = fresh()
        opt = op.Adam(m.parameters(), lr=args.eta_max, weight_decay=1e-4)
        sched = _sc.LRCosineAnnealingScheduler(args.eta_max, args.eta_min, args.Ti, args.Tm, cs["n_train"] / B)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            best = train_mmimdb_track_f1(m, crit, opt, sched, loaders, {"train": cs["n_train"], "dev": cs["n_dev"]},
                                         device=torch.device("cpu"), num_epochs=E, verbose=True)
        f1s = [float(x) for x in re.findall(r"dev F1: ([0-9.]+)", buf.getvalue())]
        assert len(f1s) == E, buf.getvalue()
        out[f"c{ci}/best_f1"] = np.float64(best)
        out[f"c{ci}/epoch_dev_f1"] = np.array(f1s)
        for k, v in m.state_dict().items():
            for kk, vv in sample_tensor(v.numpy()).items():
                out[f"c{ci}/final/{k}/{kk}"] = np.asarray(vv)
        print(ci, conf, "best F1", best, "per epoch", f1s)
    np.savez_compressed(os.path.join(OUT, "mmimdb_path.npz"), **out)


if __name__ == "__main__":
    main()
