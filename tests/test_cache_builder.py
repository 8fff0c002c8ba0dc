"""Feature-cache builder: pooling oracle against the executed reference module (CPU), the
CUDA pooling kernel and the builder against the oracle (``-m gpu``)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import GOLDEN_DIR, make_args
from oracle import pooling as P

DEV = "cuda:0"
SHAPES = ("visual5d", "ske4d", "vector", "odd", "long")
D_SKE, D_RGB = (128, 256, 1024, 512), (512, 1024, 2048, 2048)


class StubVisual(nn.Module):
    """Output structure of Visual.forward (/root/reference/models/central/ntu.py:50): four feature maps [B, C, T, W, H],
    the pooled vector and the logits; parameter-free, deterministic in the input."""

    def forward(self, x):                                       # x: [B, 6]
        B = x.shape[0]
        maps = []
        for i, (c, t, w) in enumerate(((256, 2, 3), (512, 2, 3), (1024, 1, 2), (2048, 1, 2))):
            base = torch.arange(c * t * w * w, device=x.device, dtype=torch.float32).reshape(1, c, t, w, w)
            maps.append(torch.sin(base * 0.01 * (i + 1) + x[:, :1, None, None, None]).abs())
        out5 = maps[3].reshape(B, 2048, -1).mean(2)
        return maps[0], maps[1], maps[2], maps[3], out5, x[:, :4] * 2.0


class StubSkel(nn.Module):
    """Output structure of Skeleton.forward (central/ntu.py:183): (list of 8 hidden maps, logits)."""

    def forward(self, x):                                       # x: [B, 6]
        B = x.shape[0]
        hid = [torch.zeros(B, 1, device=x.device)] * 4
        for i, (c, s) in enumerate(((128, (5, 3)), (256, (3, 2)), (1024, ()), (512, ()))):
            base = torch.arange(int(c * np.prod(s, dtype=np.int64)), device=x.device, dtype=torch.float32).reshape((1, c) + s)
            hid.append(torch.cos(base * 0.02 * (i + 1) + x[:, 1:2].reshape((B, 1) + (1,) * len(s))).abs())
        return hid, x[:, 2:6] * 0.5


class RawLoader:
    def __init__(self, n, bs):
        g = torch.Generator().manual_seed(3)
        self.x = torch.rand(n, 6, generator=g)
        self.y = torch.randint(0, 60, (n,), generator=g)
        self.dataset = range(n)
        self.bs = bs

    def __iter__(self):
        for s in range(0, len(self.x), self.bs):
            yield {'rgb': self.x[s:s + self.bs], 'ske': self.x[s:s + self.bs], 'label': self.y[s:s + self.bs]}


def test_pooling_oracle_matches_reference_module():
    fx = np.load(os.path.join(GOLDEN_DIR, "pooling.npz"))
    for name in SHAPES:
        got, ref = P.global_pool(fx[name + "_x"]), fx[name + "_y"]
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max(), name


def test_tap_selection_and_errors():
    from mfas_b200 import cache_builder as cb
    x = torch.rand(3, 6)
    ske, vis, ske_logits, vis_logits = cb.ntu_taps(StubVisual()(x), StubSkel()(x))
    assert [t.shape[1] for t in ske] == list(D_SKE) and [t.shape[1] for t in vis] == list(D_RGB)      # out_2..out_5: out_1 is dropped
    assert torch.equal(vis_logits, x[:, :4] * 2.0) and torch.equal(ske_logits, x[:, 2:6] * 0.5)
    with pytest.raises(RuntimeError):
        cb.build_feature_cache(StubVisual(), StubSkel(), RawLoader(4, 2), "cpu")
    with pytest.raises(RuntimeError):
        cb.global_pool_into(torch.rand(2, 4, 3), torch.empty(2, 4))


def test_builder_host_logic_with_the_oracle_pooling():
    """The builder loop (tap selection, column slices, short last batch, labels / logits, mode restore) on the CPU, with the
    CUDA pooling call replaced by the oracle -- the product entry point itself refuses a CPU device."""
    from mfas_b200 import cache_builder as cb
    vis, ske, loader = StubVisual(), StubSkel(), RawLoader(37, 16)
    pool = lambda t, out: out.copy_(torch.from_numpy(P.global_pool(t.numpy())))
    with torch.no_grad():
        c = cb._build(vis, ske, loader, torch.device("cpu"), pool, 32, True, cb.ntu_taps, None, ('rgb', 'ske', 'label'), None, None)
    assert len(c) == 37 and c.rgb_cat.shape == (37, 5632) and c.ske_cat.shape == (37, 1920) and vis.training
    assert torch.equal(c.labels, loader.y) and torch.equal(c.logit_rgb, loader.x[:, :4] * 2.0)
    vo, (hid, _) = vis(loader.x), ske(loader.x)
    for t, got in zip(list(vo[-5:-1]) + hid[-4:], c.rgb_taps() + c.ske_taps()):
        assert np.array_equal(got.numpy(), P.global_pool(t.numpy()))
    with pytest.raises(ValueError):
        cb._build(vis, ske, loader, torch.device("cpu"), pool, 32, False, cb.ntu_taps, None, ('rgb', 'ske', 'label'), 30, None)


@pytest.mark.gpu
def test_gpu_pooling_kernel_vs_fixture_and_oracle():
    from mfas_b200 import cache_builder as cb
    fx = np.load(os.path.join(GOLDEN_DIR, "pooling.npz"))
    for name in SHAPES:
        x = torch.from_numpy(fx[name + "_x"]).to(DEV)
        wide = torch.full((x.shape[0], x.shape[1] + 24), -7.0, device=DEV)
        cb.global_pool_into(x, wide[:, 8:8 + x.shape[1]])                 # a column slice of a wider cache matrix
        torch.cuda.synchronize()
        got = wide[:, 8:8 + x.shape[1]].cpu().numpy()
        ref = fx[name + "_y"]
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max(), (name, np.abs(got - ref).max())
        assert np.abs(got - P.global_pool(fx[name + "_x"])).max() <= 2e-6 * np.abs(ref).max(), name
        assert (wide[:, :8] == -7.0).all() and (wide[:, 8 + x.shape[1]:] == -7.0).all(), "wrote outside its slice"
    # a large tap: more rows than resident warps, unaligned row length
    x = torch.rand(64, 512, 3, 7, 7, device=DEV)
    out = torch.empty(64, 512, device=DEV)
    cb.global_pool_into(x, out)
    ref = x.double().reshape(64, 512, -1).mean(2)
    assert float((out.double() - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
    with pytest.raises(ValueError):
        cb.global_pool_into(x, torch.empty(64, 500, device=DEV))


@pytest.mark.gpu
def test_gpu_build_cache_and_train_on_it():
    """Backbones -> pooled taps -> FeatureCache -> the hot path: the built cache equals the oracle's pooling of the same
    backbone outputs, carries labels / backbone logits, and trains through train_sampled_models."""
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200 import cache_builder as cb
    from mfas_b200.cache import FeatureCacheLoader
    vis, ske = StubVisual(), StubSkel()
    vis.train(True)                                             # the builder switches to eval() and restores the mode
    loader = RawLoader(70, 16)                                   # a short last batch
    cache = cb.build_feature_cache(vis, ske, loader, DEV, with_logits=True)
    assert vis.training and len(cache) == 70 and cache.device.type == "cuda"
    assert torch.equal(cache.labels.cpu(), loader.y)
    hid, sl = ske(loader.x)
    vo = vis(loader.x)
    for t, got in zip(hid[-4:], cache.ske_taps()):
        ref = P.global_pool(t.numpy())
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
    for t, got in zip(vo[-5:-1], cache.rgb_taps()):
        ref = P.global_pool(t.numpy())
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
    assert torch.equal(cache.logit_rgb.cpu(), vo[-1]) and torch.equal(cache.logit_ske.cpu(), sl)
    # end to end: train two candidates for an epoch on a built cache
    ntu_cache = cb.build_feature_cache(vis, ske, RawLoader(96, 32), DEV)
    assert ntu_cache.rgb_cat.shape == (96, 5632) and ntu_cache.ske_cat.shape == (96, 1920)
    args = make_args(64, 32, 1)
    loaders = {"train": FeatureCacheLoader(ntu_cache, 32, True, 1), "dev": FeatureCacheLoader(ntu_cache, 32, True, 2)}
    torch.manual_seed(0)
    accs = ntu.train_sampled_models([np.array([[3, 1, 1], [0, 2, 0]]), np.array([[1, 0, 0]])], ntu.Searchable_Skeleton_Image_Net,
                                    loaders, args, torch.device(DEV))
    st = ntu.train_sampled_models.last_stats.numpy()
    assert len(accs) == 2 and np.isfinite(st).all() and (st[:, 0, 0] > 0).all()
