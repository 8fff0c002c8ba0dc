"""Drop-in for /root/reference/models/search/avmnist_searchable.py + train_searchable/avmnist.py on cached backbone taps
(the AV-MNIST twin of the NTU searchable fusion network).

Same module attributes as the reference file: ``train_sampled_models``, ``get_possible_layer_configurations`` (5 x 3 x 2),
``get_central_states`` / ``set_central_states``, ``Searchable_Audio_Image_Net``; plus the loops of
``train_searchable/avmnist.py`` (``train_avmnist_track_acc``, ``test_avmnist_track_acc``).  What differs from the NTU network
(and is carried by the same CUDA kernels):

  * taps: 5 audio taps ``channels * {1, 2, 4, 8, 16}`` (GP_LeNet_Deeper, models/central/avmnist.py:60-112) and 3 image taps
    ``channels * {1, 2, 4}`` (GP_LeNet, :18-57), already globally pooled by the backbones -- ``args.channels`` must be a
    multiple of 32 here (the kernels walk the concatenated input in 32-column k-blocks);
  * recipe: Linear -> activation [-> Dropout], never a BatchNorm (avmnist_searchable.py:276-285: the BatchNorm branches are
    commented out there), MFAS_FLAG_PLAIN in the C ABI;
  * the loop prints '<phase> Acc: ...' only and passes ``multitask`` through (train_searchable/avmnist.py:14-85).

Upstream this path has no entry point and its loop does not import (``import models.aux.scheduler``, avmnist.py:10);
``mfas_b200.install.install()`` provides the shim.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .cache import FeatureCache, FeatureCacheLoader
from .engine import flags_from_args
from .ntu_searchable import (AlphaScalarMultiplication, CachedTaps, Searchable_Skeleton_Image_Net, TrainerSpec, _activation,
                             _feature_cache_of, get_central_states, set_central_states, train_sampled)
from .train_ntu import _check_multitask, _train_track_acc

NUM_OUTPUTS = 10


def tap_widths(channels):
    """(audio tap widths, image tap widths), avmnist_searchable.py:288-292."""
    c = int(channels)
    return (c, 2 * c, 4 * c, 8 * c, 16 * c), (c, 2 * c, 4 * c)


def _check_channels(channels):
    if int(channels) < 32 or int(channels) % 32:
        raise ValueError(f"args.channels={channels}: the cached-tap kernels need tap widths that are multiples of 32 "
                         "(channels a multiple of 32); pad the backbone's channels or use the reference for other widths")


class Searchable_Audio_Image_Net(Searchable_Skeleton_Image_Net):
    """Searchable fusion head over AV-MNIST taps (avmnist_searchable.py:184-297).  ``forward(tensor_tuple)`` with
    ``tensor_tuple = (image, sound)`` -- the reference's order, :207 -- takes the concatenated cached taps (image
    [B, 7 channels], sound [B, 31 channels]) and returns [B, num_outputs] logits (the 3-tuple when ``args.multitask``)."""

    _extra_flags = _lib.FLAG_PLAIN

    @staticmethod
    def _tap_widths(args):
        _check_channels(args.channels)
        return tap_widths(args.channels)

    def __init__(self, args, conf):
        cf = np.asarray(conf).reshape(-1, 3)
        if cf[:, 0].max() >= 5 or cf[:, 1].max() >= 3 or cf[:, :2].min() < 0:
            raise ValueError(f"conf rows are [audio tap < 5, image tap < 3, activation]: {cf.tolist()}")
        super().__init__(args, conf)
        self._widths_kw = tap_widths(args.channels)
        ds, dr = self._widths_kw
        self.rgbnet = CachedTaps(dr, "lenet")           # GP_LeNet stands here in the reference (:200)
        self.skenet = CachedTaps(ds, "lenet")           # GP_LeNet_Deeper (:201)

    @property
    def audnet(self):
        return self.skenet

    def _create_fc_layers(self, cf):
        """Linear -> activation [-> Dropout]: avmnist_searchable.py:258-285 (no BatchNorm branch)."""
        H, drpt = self.args.inner_representation_size, self.args.drpt
        layers = []
        for i, c in enumerate(cf):
            in_size = self.alphas[i].size_alpha_x + self.alphas[i].size_alpha_y + (H if i > 0 else 0)
            nl = _activation(c[2])
            layers.append(nn.Sequential(nn.Linear(in_size, H), nl, nn.Dropout(drpt)) if drpt > 1e-10 else nn.Sequential(nn.Linear(in_size, H), nl))
        return nn.ModuleList(layers)


def get_possible_layer_configurations(progression_index):
    """All [audio tap, image tap, activation] rows of one fusion step: 5 x 3 x 2 = 30 (avmnist_searchable.py:108-122)."""
    return [[t, v, n] for t in range(5) for v in range(3) for n in range(2)]


def _flags(args):
    f = flags_from_args(args) & ~_lib.FLAG_BN           # args.batchnorm has no effect on this network (:276-285)
    return f | _lib.FLAG_PLAIN


class AudioImageCacheLoader(FeatureCacheLoader):
    """FeatureCacheLoader yielding the reference's AV-MNIST batch dict {'image','audio','label'} (train_searchable/avmnist.py:36)."""

    def __iter__(self):
        order = self.order_for_pass(self.take_passes(1))
        c = self.dataset
        for s in range(0, len(order), self.batch_size):
            rows = order[s:s + self.batch_size].to(c.device)
            img, aud = c.rgb_cat.index_select(0, rows), c.ske_cat.index_select(0, rows)
            if c.logit_rgb is not None:              # multitask: the cached backbone logits ride behind the taps of their modality
                img = torch.cat((img, c.logit_rgb.index_select(0, rows)), 1)
                aud = torch.cat((aud, c.logit_ske.index_select(0, rows)), 1)
            yield {'image': img, 'audio': aud, 'label': c.labels.index_select(0, rows)}


def audio_image_cache(audio_cat, image_cat, labels, channels, logit_image=None, logit_audio=None) -> FeatureCache:
    """audio_cat [N, 31 channels] / image_cat [N, 7 channels] fp32 cached (pooled) taps, labels [N] int64."""
    _check_channels(channels)
    return FeatureCache(audio_cat, image_cat, labels, logit_rgb=logit_image, logit_ske=logit_audio, widths=tap_widths(channels))


def synthetic_avmnist_cache(n_rows: int, seed: int, channels: int = 32, num_outputs: int = NUM_OUTPUTS, signal: float = 2.0,
                            with_backbone_logits: bool = False) -> FeatureCache:
    """AV-MNIST-shaped synthetic split: |N(0,1)| taps (post-ReLU, pooled), ``signal`` * onehot(label) added to the first
    ``num_outputs`` columns of audio tap 4 and image tap 2 so that accuracy is learnable."""
    g = torch.Generator().manual_seed(int(seed))
    da, di = tap_widths(channels)
    aud = torch.randn(n_rows, sum(da), generator=g).abs_()
    img = torch.randn(n_rows, sum(di), generator=g).abs_()
    labels = torch.randint(0, num_outputs, (n_rows,), generator=g, dtype=torch.int64)
    onehot = torch.zeros(n_rows, num_outputs).scatter_(1, labels[:, None], float(signal))
    aud[:, sum(da[:4]):sum(da[:4]) + num_outputs] += onehot
    img[:, sum(di[:2]):sum(di[:2]) + num_outputs] += onehot
    li = la = None
    if with_backbone_logits:
        li = torch.randn(n_rows, num_outputs, generator=g) + onehot * 0.5
        la = torch.randn(n_rows, num_outputs, generator=g) + onehot * 0.5
    return audio_image_cache(aud, img, labels, channels, li, la)


class _AVMNISTSpec(TrainerSpec):
    def flags(self, args):
        return _flags(args)

    def check(self, args, flags, preaccuracies):
        _check_channels(args.channels)                  # (the reference's loop takes multitask and no init_f1: nothing to reject)

    def load_backbones(self, rmode, args):
        pass                                            # cached taps: there are no backbones to load (avmnist_searchable.py:44-60)

    def log_epochs(self, stats, n_train, n_dev):
        for e in range(stats.shape[0]):                 # train_searchable/avmnist.py:79
            print('{} Acc: {:.4f}'.format('train', stats[e, 1] / n_train))
            print('{} Acc: {:.4f}'.format('dev', stats[e, 3] / n_dev))

    def vid_len(self, args):
        return 32


def train_sampled_models(sampled_configurations, searchable_type, dataloaders,
                         args, device,
                         return_model=[], premodels=[], preaccuracies=[],
                         train_only_central_params=True,
                         state_dict=dict()):
    """Train every sampled configuration and return its best dev accuracy, in input order: signature and semantics of
    /root/reference/models/search/avmnist_searchable.py:22-105 (Adam(lr=eta_max, weight_decay=1e-4), per-batch cosine LR,
    ``args.epochs`` x (train pass, dev pass), strict-'>' best-dev tracking and rollback, ``multitask`` passed through); the
    body is the NTU trainer's (mfas_b200.ntu_searchable.train_sampled)."""
    spec = _AVMNISTSpec()
    spec.own_class = Searchable_Audio_Image_Net
    spec.widths = tap_widths(args.channels)
    return train_sampled(spec, train_sampled_models, sampled_configurations, searchable_type, dataloaders, args, device,
                         return_model, premodels, preaccuracies, state_dict)


def train_avmnist_track_acc(model, criteria, optimizer, scheduler, dataloaders, dataset_sizes,
                            device=None, num_epochs=200, verbose=False, multitask=False):
    """num_epochs x (train pass, dev pass) with best-dev rollback; returns the best dev accuracy
    (/root/reference/models/search/train_searchable/avmnist.py:14-85)."""
    return _train_track_acc(train_avmnist_track_acc, model, optimizer, scheduler, dataloaders, device, num_epochs, multitask, with_loss=False)


def test_avmnist_track_acc(model, dataloaders, dataset_sizes, device=None, multitask=False):
    """Eval-mode accuracy over dataloaders['test'] (train_searchable/avmnist.py:88-125)."""
    net = model.module if isinstance(model, torch.nn.DataParallel) else model
    _check_multitask(net, multitask, dataloaders, ('test',))
    model.train(False)
    g = net.native(device)
    test_c = _feature_cache_of(dataloaders['test'], 'test').to(g.device)
    B = int(getattr(dataloaders['test'], 'batch_size', None) or net.args.batchsize)
    if B > g.batch_max:
        g = net.native(device, batch_max=B)
    out = g.eval_pass(test_c, B).cpu()
    return (out[0, 1] / dataset_sizes['test']).clone()


test_avmnist_track_acc.__test__ = False       # (not a pytest test)
