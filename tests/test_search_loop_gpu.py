"""The candidate trainer inside a search loop, on the GPU, with the REAL trainer (tests/test_driver_dropin.py runs the reference's
own `_epnas` on the CPU with the trainer stubbed; the reference tree does not travel to the GPU box).

The loop below issues exactly the call sequence of `ModelSearcher._epnas` (/root/reference/models/searchable.py:48-137) for one
search iteration of two progression levels -- every unfolded one-step configuration, then K sampled two-step configurations
built by `merge_unfolded_with_sampled` (models/search/tools.py:68-93: previous top-K x the 32 unfolded rows) -- through
`train_sampled_models(confs, model_type, dataloaders, args, device, state_dict=shared)` and consumes the results the way the
driver does (`np.array(accs)`, probabilities from accuracies).  Checked: return types and order, determinism under a seed, and
that the single-process multi-device fan-out returns the identical lists.
"""
import numpy as np
import pytest
import torch

from helpers import make_args
from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _search(args, train, dev, seed):
    import mfas_b200.ntu_searchable as ntu
    rng = np.random.RandomState(seed)                  # the driver's np.random.choice (tools.py:53), seeded here
    torch.manual_seed(seed)
    loaders = {"train": FeatureCacheLoader(train, args.batchsize, True, seed), "dev": FeatureCacheLoader(dev, args.batchsize, True, seed + 50000)}
    shared = dict()
    rows = ntu.get_possible_layer_configurations(0)
    level0 = [np.expand_dims(np.array(r), 0) for r in rows]                                            # tools.py:80-81
    accs0 = ntu.train_sampled_models(level0, ntu.Searchable_Skeleton_Image_Net, loaders, args, torch.device(DEV), state_dict=shared)
    a0 = np.array(accs0)                                                                               # tools.py:47
    assert a0.shape == (32,) and a0.dtype == np.float64 and np.isfinite(a0).all() and (a0 >= 0).all() and a0.max() > 1.5 / 60
    p = a0 / a0.sum()
    top = [level0[i] for i in rng.choice(len(level0), args.num_samples, replace=False, p=p)]           # tools.py:46-56
    level1 = [np.concatenate([t, np.expand_dims(np.array(r), 0)], 0) for t in top for r in rows]      # tools.py:84-91
    sampled = [level1[i] for i in rng.choice(len(level1), args.num_samples, replace=False)]
    accs1 = ntu.train_sampled_models(sampled, ntu.Searchable_Skeleton_Image_Net, loaders, args, torch.device(DEV), state_dict=shared)
    assert len(accs1) == args.num_samples and all(a.dtype == torch.float64 and a.dim() == 0 and a.device.type == "cpu" for a in accs1)
    assert shared == {}                                                                                 # args.weightsharing is off
    return a0, np.array(accs1), [s.tolist() for s in sampled]


def test_one_search_iteration_with_the_real_trainer(monkeypatch):
    args = make_args(16, 64, 2, bn=True, drpt=0.0, Ti=1)                  # main_searchable_ntu.py defaults: inner_repr 16
    args.num_samples = 4
    train, dev = synthetic_ntu_cache(1024, 1), synthetic_ntu_cache(512, 2)
    a0, a1, confs = _search(args, train, dev, 7)
    b0, b1, confs_b = _search(args, train, dev, 7)
    assert confs == confs_b and np.array_equal(a0, b0) and np.array_equal(a1, b1)                     # one seed, one result
    # the same search with every call fanned out over the devices of this process (3 host threads on a 1-GPU box)
    import mfas_b200.ntu_searchable as ntu
    args.init_on_device = True
    c0, c1, confs_c = _search(args, train, dev, 7)
    args.fanout_gpus = "all" if torch.cuda.device_count() >= 2 else 3
    if torch.cuda.device_count() < 2:
        monkeypatch.setattr(ntu, "fanout_devices", lambda a, d: [torch.device(DEV)] * 3)
    d0, d1, confs_d = _search(args, train, dev, 7)
    assert confs_c == confs_d and np.array_equal(c0, d0) and np.array_equal(c1, d1)
