#!/usr/bin/env python3
"""Fixture for the main_found_ntu.py flow (multitask 3-head loss, two-stage training, test pass), produced by
EXECUTING the unmodified reference ``main_found_ntu.train_model`` (/root/reference/main_found_ntu.py:94-157) ->
``train_ntu_track_acc(..., multitask=True)`` / ``test_ntu_track_acc`` (models/search/train_searchable/ntu.py) on CPU,
fed from this repo's synthetic cache WITH cached backbone logits through parameter-free stub backbones that hand the
logits through.  Run in the build container:

    python tests/golden/gen_golden_found.py        # writes tests/golden/found_mt.npz
"""
import contextlib
import io
import os
import re
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
import models.auxiliary.scheduler as _sc  # noqa: E402

for n in ("models.aux", "models.train"):
    pkg = types.ModuleType(n)
    pkg.scheduler = _sc
    sys.modules[n] = pkg
    sys.modules[n + ".scheduler"] = _sc
import models.central.ntu as central  # noqa: E402

SKE, RGB = [128, 256, 1024, 512], [512, 1024, 2048, 2048]


class StubVisual(nn.Module):
    """Visual.forward's 6-tuple (models/central/ntu.py:50): taps + the cached logits riding behind them."""

    def __init__(self, args):
        super().__init__()

    def forward(self, x):
        return (None, *torch.split(x[:, :sum(RGB)], RGB, 1), x[:, sum(RGB):])


class StubSkel(nn.Module):
    def __init__(self, args):
        super().__init__()

    def forward(self, x):
        return [None] * 4 + list(torch.split(x[:, :sum(SKE)], SKE, 1)), x[:, sum(SKE):]


central.Visual, central.Skeleton = StubVisual, StubSkel
import models.search.ntu_searchable as ntu  # noqa: E402
import main_found_ntu as found  # noqa: E402

from helpers import FOUND_MT_CASE, make_args, sample_tensor  # noqa: E402
from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache  # noqa: E402


def main():
    cs = FOUND_MT_CASE
    torch.set_num_threads(1)
    tmp = tempfile.mkdtemp()
    torch.save({}, os.path.join(tmp, "ske"))
    torch.save({}, os.path.join(tmp, "rgb"))
    args = make_args(cs["H"], cs["B"], cs["epochs"], bn=True, drpt=0.0, Ti=cs["Ti"], checkpointdir=tmp,
                     alphas=cs["alphas"], multitask=True, verbose=True)
    args.test_cp = ''
    splits = {k: synthetic_ntu_cache(n, cs["data_seed"] + i, with_backbone_logits=True)
              for i, (k, n) in enumerate((("train", cs["n_train"]), ("dev", cs["n_dev"]), ("test", cs["n_test"])))}
    loaders = {k: FeatureCacheLoader(v, cs["B"], True, cs["loader_seed"] + 1000 * i) for i, (k, v) in enumerate(splits.items())}
    torch.manual_seed(cs["model_seed"])
    conf = np.array(cs["conf"])
    rmode = ntu.Searchable_Skeleton_Image_Net(args, conf)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        test_acc = found.train_model(rmode, conf, loaders, args, torch.device("cpu"))
    text = buf.getvalue()
    out = {"test_acc": np.float64(float(test_acc))}
    rows = re.findall(r"(train|dev) Loss: ([0-9.]+) Acc: ([0-9.]+)", text)
    out["epoch_phase"] = np.array([r[0] for r in rows])
    out["epoch_loss"] = np.array([float(r[1]) for r in rows])
    out["epoch_acc"] = np.array([float(r[2]) for r in rows])
    m = re.search(r"Intermediate val accuracy: tensor\(([0-9.]+)", text)
    out["interm_acc"] = np.float64(float(m.group(1)))
    m = re.search(r"Final val accuracy: tensor\(([0-9.]+)", text)
    out["final_acc"] = np.float64(float(m.group(1)))
    for k, v in rmode.state_dict().items():
        for kk, vv in sample_tensor(v.numpy()).items():
            out[f"final/{k}/{kk}"] = np.asarray(vv)
    out["meta/torch"] = np.array(torch.__version__)
    path = os.path.join(os.environ.get("MFAS_GOLDEN_OUT", HERE), "found_mt.npz")
    np.savez_compressed(path, **out)
    print(text[-600:])
    print("->", path, os.path.getsize(path) // 1024, "KiB; test acc", float(test_acc))


if __name__ == "__main__":
    main()
