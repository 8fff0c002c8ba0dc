#!/usr/bin/env python3
"""Generate parity fixtures by EXECUTING the unmodified reference (run in the build container).

    python tests/golden/gen_golden.py          # writes tests/golden/<case>.npz

What runs: /root/reference's own ``train_sampled_models`` -> ``train_ntu_track_acc`` ->
``Searchable_Skeleton_Image_Net.forward`` + ``torch.optim.Adam`` + ``LRCosineAnnealingScheduler``
(models/search/ntu_searchable.py:23-102, models/search/train_searchable/ntu.py:14-89,
models/auxiliary/scheduler.py:12-46), on CPU, fed from this repo's synthetic feature cache
through parameter-free stub backbones.  Nothing in the reference is
edited; three ``sys.modules`` shims make it importable here (matplotlib is absent, and
``models.aux`` / ``models.train`` are dangling imports in files we never call).

The GPU box has no /root/reference: tests read only the .npz files written here.
"""
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
REF = "/root/reference"
sys.path.insert(0, REF)

for n in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[n] = types.ModuleType(n)
import models.auxiliary.scheduler as _sc  # noqa: E402  (real reference module)

for n in ("models.aux", "models.train"):
    pkg = types.ModuleType(n)
    pkg.scheduler = _sc
    sys.modules[n] = pkg
    sys.modules[n + ".scheduler"] = _sc
import models.central.ntu as central  # noqa: E402

SKE, RGB = [128, 256, 1024, 512], [512, 1024, 2048, 2048]


class StubVisual(nn.Module):
    """Mimics Visual.forward's 6-tuple (models/central/ntu.py:50) on cached taps."""

    def __init__(self, args):
        super().__init__()

    def forward(self, x):
        return (None, *torch.split(x, RGB, 1), None)


class StubSkel(nn.Module):
    """Mimics Skeleton.forward's (hiddens, logits) (models/central/ntu.py:183)."""

    def __init__(self, args):
        super().__init__()

    def forward(self, x):
        return [None] * 4 + list(torch.split(x, SKE, 1)), None


central.Visual, central.Skeleton = StubVisual, StubSkel
import models.search.ntu_searchable as ntu  # noqa: E402

from helpers import GOLDEN_CASES, WS_CASES, init_states, make_args, sample_tensor  # noqa: E402
from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache  # noqa: E402

LOADER_SEED = 100


def flat(prefix, d, out):
    for k, v in d.items():
        out[f"{prefix}/{k}"] = np.asarray(v)


def run_case(name, cs):
    torch.set_num_threads(1)            # deterministic summation order for the fixtures
    confs = [np.array(c) for c in cs["confs"]]
    H, B, E = cs["H"], cs["B"], cs["epochs"]
    tmp = tempfile.mkdtemp()
    torch.save({}, os.path.join(tmp, "ske"))
    torch.save({}, os.path.join(tmp, "rgb"))
    args = make_args(H, B, E, bn=cs["bn"], drpt=cs["drpt"], Ti=cs["Ti"], checkpointdir=tmp, alphas=cs.get("alphas", False),
                     weightsharing=cs.get("weightsharing", False))
    train = synthetic_ntu_cache(cs["n_train"], cs["data_seed"])
    dev = synthetic_ntu_cache(cs["n_dev"], cs["data_seed"] + 1)
    loaders = {"train": FeatureCacheLoader(train, B, True, LOADER_SEED),
               "dev": FeatureCacheLoader(dev, B, True, LOADER_SEED + 50000)}

    out = {}
    # ---- (1) init parity of tests/helpers.init_states with the reference constructor
    inits = init_states(cs["confs"], H, 60, cs["bn"], cs["drpt"], cs["model_seed"])
    torch.manual_seed(cs["model_seed"])
    for ci, conf in enumerate(confs):
        m = ntu.Searchable_Skeleton_Image_Net(args, conf)
        sd = m.state_dict()
        assert set(sd.keys()) == set(inits[ci].keys()), (sorted(sd.keys()), sorted(inits[ci].keys()))
        for k, v in sd.items():
            assert np.array_equal(v.numpy(), inits[ci][k]), k

    # ---- (2) one manual step per candidate: logits, loss, autograd grads (pins the derivation)
    torch.manual_seed(cs["model_seed"])
    for ci, conf in enumerate(confs):
        m = ntu.Searchable_Skeleton_Image_Net(args, conf)
        m.train(True)
        rows = loaders["train"].order_for_pass(ci * E)[:B]
        batch = (train.rgb_cat[rows], train.ske_cat[rows])
        logits = m(batch)
        loss = torch.nn.CrossEntropyLoss()(logits, train.labels[rows])
        loss.backward()
        out[f"c{ci}/step0_logits"] = logits.detach().numpy()
        out[f"c{ci}/step0_loss"] = np.float32(loss.item())
        for k, p in m.named_parameters():
            if p.grad is not None:
                flat(f"c{ci}/grad/{k}", sample_tensor(p.grad.numpy()), out)

    # ---- (3) the real thing: unmodified train_sampled_models over all candidates
    rec = {"logits": [], "adam": []}

    class RecAdam(torch.optim.Adam):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            rec["adam"].append(self)

    hooks_model = []
    orig_init = ntu.Searchable_Skeleton_Image_Net.__init__

    def patched_init(self, *a, **k):
        orig_init(self, *a, **k)
        self.register_forward_hook(lambda mod, inp, outp: rec["logits"].append(
            (len(hooks_model) - 1, mod.training, outp.detach().numpy().copy())))
        hooks_model.append(self)

    ntu.Searchable_Skeleton_Image_Net.__init__ = patched_init
    real_adam = torch.optim.Adam
    torch.optim.Adam = RecAdam
    ntu.op.Adam = RecAdam
    try:
        torch.manual_seed(cs["model_seed"])
        shared = {}                     # args.weightsharing: the dict the reference chains the candidates through (:27,:74-75,:91-92)
        accs, models = ntu.train_sampled_models(confs, ntu.Searchable_Skeleton_Image_Net, loaders, args,
                                                torch.device("cpu"), return_model=list(range(len(confs))), state_dict=shared)
    finally:
        torch.optim.Adam = real_adam
        ntu.op.Adam = real_adam
        ntu.Searchable_Skeleton_Image_Net.__init__ = orig_init

    n_tb = (cs["n_train"] + B - 1) // B
    n_db = (cs["n_dev"] + B - 1) // B
    for ci, conf in enumerate(confs):
        out[f"c{ci}/best_acc"] = np.float64(float(accs[ci]))
        calls = [(tr, lg) for (mi, tr, lg) in rec["logits"] if mi == ci]
        assert len(calls) == E * (n_tb + n_db), (len(calls), E, n_tb, n_db)
        losses_tr, losses_dv, corr_tr, corr_dv = [], [], [], []
        k = 0
        for e in range(E):
            otr = loaders["train"].order_for_pass(ci * E + e)
            odv = loaders["dev"].order_for_pass(ci * E + e)
            for phase, order, nb, split in (("train", otr, n_tb, train), ("dev", odv, n_db, dev)):
                for bi in range(nb):
                    tr, lg = calls[k]
                    k += 1
                    assert tr == (phase == "train")
                    y = split.labels[order[bi * B:(bi + 1) * B]]
                    lgt = torch.from_numpy(lg)
                    loss = torch.nn.functional.cross_entropy(lgt, y).item()
                    corr = int((lgt.argmax(1) == y).sum())
                    (losses_tr if phase == "train" else losses_dv).append(loss)
                    (corr_tr if phase == "train" else corr_dv).append(corr)
                    if phase == "train" and e == 0 and bi < 3:
                        out[f"c{ci}/train_logits_e0_b{bi}"] = lg
                    if phase == "dev" and e == E - 1 and bi == nb - 1:
                        out[f"c{ci}/dev_logits_last"] = lg
        out[f"c{ci}/train_loss"] = np.array(losses_tr, np.float32).reshape(E, n_tb)
        out[f"c{ci}/dev_loss"] = np.array(losses_dv, np.float32).reshape(E, n_db)
        out[f"c{ci}/train_correct"] = np.array(corr_tr, np.int64).reshape(E, n_tb)
        out[f"c{ci}/dev_correct"] = np.array(corr_dv, np.int64).reshape(E, n_db)
        # final (rolled back to best dev epoch) weights + the optimiser's last moments
        for kname, v in models[ci].state_dict().items():
            flat(f"c{ci}/final/{kname}", sample_tensor(v.numpy()), out)
        opt = rec["adam"][ci]
        named = {id(p): n for n, p in models[ci].named_parameters()}
        for p, st in opt.state.items():
            flat(f"c{ci}/adam_m/{named[id(p)]}", sample_tensor(st["exp_avg"].numpy()), out)
            flat(f"c{ci}/adam_v/{named[id(p)]}", sample_tensor(st["exp_avg_sq"].numpy()), out)
            out[f"c{ci}/adam_t"] = np.int64(int(st["step"]))
    if cs.get("weightsharing", False):
        out["meta/shared_keys"] = np.array(sorted(shared.keys()))
        for key, sd in shared.items():                      # what the dict holds after the call (the last writer's layer)
            for kname, v in sd.items():
                flat(f"shared/{key}/{kname}", sample_tensor(v.numpy()), out)
    out["meta/torch"] = np.array(torch.__version__)
    out["meta/loader_seed"] = np.int64(LOADER_SEED)
    path = os.path.join(os.environ.get("MFAS_GOLDEN_OUT", HERE), name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; best accs",
          [float(a) for a in accs])


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, cs in {**GOLDEN_CASES, **WS_CASES}.items():
        if not only or name in only:
            run_case(name, cs)
