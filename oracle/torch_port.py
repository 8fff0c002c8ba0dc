"""PyTorch-CPU port of the reference candidate-training path (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference *is* eager PyTorch; this file restates its model, loss, optimiser and loop with the same
torch modules in the same order, fed from cached taps, so that ``bench.py --impl reference`` and the
``cpu_baseline`` leg time the reference's own arithmetic stack (ATen CPU kernels, torch.optim.Adam) on
the GPU box's host cores, where /root/reference itself does not exist.  Never imported by mfas_b200.

Parity status: PINNED -- tests/test_oracle_golden.py::test_torch_port_matches_reference_fixture checks
it against fixtures produced by executing the unmodified reference.

Restates: models/search/ntu_searchable.py:178-301 (model), :23-102 (candidate loop),
models/search/train_searchable/ntu.py:14-89 (epoch loop), models/auxiliary/scheduler.py:12-46.
"""
import copy

import numpy as np
import torch
import torch.nn as nn

from .mfas_oracle import D_RGB, CosineRestartLR, d_ske


class FusionHeadTorch(nn.Module):
    def __init__(self, conf, H, C, batchnorm=True, drpt=0.0, vid_len_ske=32):
        super().__init__()
        self.conf = np.asarray(conf).reshape(-1, 3)
        ds = d_ske(vid_len_ske)
        self.ds = ds
        layers = []
        for l, (i, j, a) in enumerate(self.conf):
            K = ds[i] + D_RGB[j] + (H if l > 0 else 0)
            nl = {0: nn.ReLU, 1: nn.Sigmoid, 2: nn.LeakyReLU}[int(a)]()
            mods = [nn.Linear(K, H), nl]
            if batchnorm:
                mods.append(nn.BatchNorm1d(H))
            if drpt > 1e-10:
                mods.append(nn.Dropout(drpt))
            if not batchnorm and drpt < 1e-10:
                raise UnboundLocalError("no layer recipe for drpt<1e-10 and batchnorm=False")
            layers.append(nn.Sequential(*mods))
        self.fusion_layers = nn.ModuleList(layers)
        self.central_classifier = nn.Linear(H, C)

    def forward(self, ske_cat, rgb_cat):
        so = np.concatenate([[0], np.cumsum(self.ds)])
        ro = np.concatenate([[0], np.cumsum(D_RGB)])
        out = None
        for l, (i, j, _) in enumerate(self.conf):
            s = ske_cat[:, so[i]:so[i + 1]]
            r = rgb_cat[:, ro[j]:ro[j + 1]]
            fused = torch.cat((s, r), 1) if l == 0 else torch.cat((s, r, out), 1)
            out = self.fusion_layers[l](fused)
        return self.central_classifier(out)


def train_candidate(model, train, dev, orders, batch, epochs, eta_max=1e-3, eta_min=1e-6, Ti=1, Tm=2,
                    max_train_steps=None, max_dev_steps=None):
    """train / dev: (ske_cat, rgb_cat, labels) tensors.  orders(phase, epoch) -> LongTensor.
    Returns (best_acc, per-epoch stats).  max_*_steps bound the work for the timed CPU sample."""
    opt = torch.optim.Adam(model.parameters(), lr=eta_max, weight_decay=1e-4)
    sched = CosineRestartLR(eta_max, eta_min, Ti, Tm, train[2].shape[0] / batch)
    crit = nn.CrossEntropyLoss()
    best_sd, best_acc, stats = copy.deepcopy(model.state_dict()), 0.0, []
    for e in range(epochs):
        row = {}
        for phase, (ske, rgb, lab) in (("train", train), ("dev", dev)):
            model.train(phase == "train")
            order = orders(phase, e)
            n = len(order)
            rl, rc, nsteps = 0.0, 0, 0
            for s0 in range(0, n, batch):
                lim = max_train_steps if phase == "train" else max_dev_steps
                if lim is not None and nsteps >= lim:
                    break
                rows = order[s0:s0 + batch]
                x_s, x_r, y = ske[rows], rgb[rows], lab[rows]
                opt.zero_grad()
                with torch.set_grad_enabled(phase == "train"):
                    out = model(x_s, x_r)
                    loss = crit(out, y)
                    if phase == "train":
                        lr = sched.step()
                        for g in opt.param_groups:
                            g["lr"] = lr
                        loss.backward()
                        opt.step()
                rl += loss.item() * len(rows)
                rc += int((out.argmax(1) == y).sum())
                nsteps += 1
            row[phase + "_loss"], row[phase + "_acc"], row[phase + "_steps"] = rl / n, rc / n, nsteps
            if phase == "dev" and row["dev_acc"] > best_acc:
                best_acc, best_sd = row["dev_acc"], copy.deepcopy(model.state_dict())
        stats.append(row)
    model.load_state_dict(best_sd)
    model.train(False)
    return best_acc, stats
