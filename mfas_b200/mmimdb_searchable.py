"""MM-IMDB searchable fusion on cached text / image taps (BASELINE.json configs[3]).

The reference has MM-IMDB backbones (/root/reference/models/central/mm_imdb.py), a training loop
(/root/reference/models/search/train_searchable/mmimdb.py:14-136) and the loss
(/root/reference/models/auxiliary/aux_models.py:129-147) but no *searchable* network for them.  This module
supplies it by analogy with ``models/search/ntu_searchable.py`` -- same attribute names, same signatures:

    Searchable_Text_Image_Net(args, conf)      conf rows [text tap, image tap, activation]
    train_sampled_models(...)                  -> list of best dev F1-samples, 0-dim float64 CPU tensors
    train_mmimdb_track_f1(...)                 the reference loop's signature (train_searchable/mmimdb.py:14-15)
    get_possible_layer_configurations(i)       2 x 4 x 2 = 16 rows
    get_central_states / set_central_states    shared with the NTU module (keys carry the layer shapes)

Taps: text = the two hidden layers of ``MaxOut_MLP`` (central/mm_imdb.py:176-196: 64, 128 wide), image = the four pooled
``GP_VGG`` blocks (central/mm_imdb.py:41-52: 512 wide each).  Concat order ``[text | image | hidden]``.  The fusion steps
run in the same CUDA kernels as the NTU head (the tap set is a (pointer, width) table); the head is the multi-label one:
``WeightedCrossEntropyWithLogits`` and per-sample F1 of ``sigmoid > 0.3`` (MFAS_FLAG_MULTILABEL, kernels_ffma.cuh
``head_rows_ml``).  With inner_representation_size a multiple of 64 and batches of at most 64 rows the group runs on the
tensor-core engine (forward / backward streaming kernels + per-layer chain kernels + the CUDA-core head kernel; the backward
tile list is cut at the concat-source boundaries, so the 64-wide text tap is a short tile); otherwise on the CUDA-core
engine.  No CPU path: a non-CUDA device raises.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from .cache import FeatureCache, FeatureCacheLoader
from .engine import CandidateGroup, GroupLayout, flags_from_args
from .ntu_searchable import (Searchable_Skeleton_Image_Net, TrainerSpec, _feature_cache_of, _reserve_passes, cosine_lrs,
                             get_central_states, pass_orders, set_central_states, train_sampled)
from .scheduler import is_per_batch_cosine

D_TEXT = (64, 128)
D_IMAGE = (512, 512, 512, 512)
WIDTHS = (D_TEXT, D_IMAGE)
NUM_OUTPUTS = 23                 # /root/reference/datasets/mm_imdb.py (23 genres)
TH_FSCORE = 0.3                  # train_searchable/mmimdb.py:15: the only threshold the kernels implement


class WeightedCrossEntropyWithLogits(torch.nn.Module):
    """Criterion object with the reference's constructor (aux_models.py:129-134); the arithmetic runs in the CUDA head."""

    def __init__(self, pos_weight):
        super().__init__()
        self.w = pos_weight

    def forward(self, logits, targets):
        raise RuntimeError("the loss is computed inside the CUDA library (train_mmimdb_track_f1); there is no eager path")


class TextImageCacheLoader(FeatureCacheLoader):
    """FeatureCacheLoader yielding the reference's MM-IMDB batch dict {'image','text','label'}
    (train_searchable/mmimdb.py:55)."""

    def __iter__(self):
        order = self.order_for_pass(self.take_passes(1))
        c = self.dataset
        for s in range(0, len(order), self.batch_size):
            rows = order[s:s + self.batch_size].to(c.device)
            yield {'image': c.rgb_cat.index_select(0, rows), 'text': c.ske_cat.index_select(0, rows),
                   'label': c.labels.index_select(0, rows)}


def text_image_cache(text_cat, image_cat, targets, pos_weight) -> FeatureCache:
    """text_cat [N, 192] / image_cat [N, 2048] fp32 cached taps, targets [N, C] multi-hot fp32, pos_weight [C]."""
    return FeatureCache(text_cat, image_cat, targets, widths=WIDTHS,
                        pos_weight=torch.as_tensor(np.asarray(pos_weight, np.float32)))


def synthetic_mmimdb_cache(n_rows: int, seed: int, num_outputs: int = NUM_OUTPUTS, signal: float = 2.0) -> FeatureCache:
    """MM-IMDB-shaped synthetic split: |N(0,1)| taps, ~2.5 genres per sample, ``signal`` * targets added to the first
    ``num_outputs`` columns of text tap 1 and image tap 0 so that F1 is learnable; pos_weight = (1 - f) / f of the genre
    frequencies of a fixed prior (the usual choice for this loss)."""
    g = torch.Generator().manual_seed(int(seed))
    text = torch.randn(n_rows, sum(D_TEXT), generator=g).abs_()
    image = torch.randn(n_rows, sum(D_IMAGE), generator=g).abs_()
    freq = torch.linspace(0.30, 0.03, num_outputs)
    targets = (torch.rand(n_rows, num_outputs, generator=g) < freq).float()
    text[:, D_TEXT[0]:D_TEXT[0] + num_outputs] += signal * targets
    image[:, :num_outputs] += signal * targets
    return text_image_cache(text, image, targets, ((1 - freq) / freq).numpy())


class Searchable_Text_Image_Net(Searchable_Skeleton_Image_Net):
    """Searchable fusion head over MM-IMDB taps.  ``forward(text, image)`` -- the call of the reference loop
    (train_searchable/mmimdb.py:68) -- takes the concatenated cached taps (text [B, 192], image [B, 2048]) and returns
    [B, num_outputs] logits computed by the CUDA library.  state_dict keys are those of the NTU network."""

    _widths_kw = WIDTHS
    _extra_flags = _lib.FLAG_MULTILABEL

    @staticmethod
    def _tap_widths(args):
        return D_TEXT, D_IMAGE

    def __init__(self, args, conf):
        cf = np.asarray(conf).reshape(-1, 3)
        if cf[:, 0].max() >= len(D_TEXT) or cf[:, 1].max() >= len(D_IMAGE) or cf[:, :2].min() < 0:
            raise ValueError(f"conf rows are [text tap < {len(D_TEXT)}, image tap < {len(D_IMAGE)}, activation]: {cf.tolist()}")
        if getattr(args, "multitask", False):
            raise ValueError("the MM-IMDB head has no multitask variant")
        super().__init__(args, conf)

    @property
    def textnet(self):
        return self.skenet

    @property
    def imagenet(self):
        return self.rgbnet

    def forward(self, text, image=None):
        if image is None:                                  # also accept the NTU-style tuple
            text, image = text
        nt, ni = sum(D_TEXT), sum(D_IMAGE)
        if text.dim() != 2 or image.dim() != 2 or text.shape[1] != nt or image.shape[1] != ni:
            raise ValueError("expected cached taps: text [B, %d], image [B, %d]" % (nt, ni))
        g = self.native(text.device)
        B, Cn = text.shape[0], self.args.num_outputs
        if B > _lib.MAX_BATCH:
            raise ValueError(f"batch of {B} rows exceeds MFAS_MAX_BATCH={_lib.MAX_BATCH}")
        if B > g.batch_max:                                # the group was sized by a training loop with a smaller batch
            g = self.native(text.device, batch_max=_lib.MAX_BATCH)
        cache = FeatureCache(text.contiguous(), image.contiguous(), torch.zeros(B, Cn, device=text.device), widths=WIDTHS,
                             pos_weight=torch.ones(Cn, device=text.device))
        rows = torch.arange(B, dtype=torch.int32, device=text.device)
        logits, _, _ = g.forward(cache, rows, train=self.training, step=self._fwd_steps)
        if self.training:
            self._fwd_steps += 1
        return logits[self._slot].clone()


def get_possible_layer_configurations(progression_index):
    """All [text tap, image tap, activation] rows of one fusion step: 2 x 4 x 2 = 16 (by analogy with
    /root/reference/models/search/ntu_searchable.py:105-119)."""
    return [[t, v, n] for t in range(len(D_TEXT)) for v in range(len(D_IMAGE)) for n in range(2)]


def _flags(args):
    return flags_from_args(args) | _lib.FLAG_MULTILABEL


def _with_pos_weight(cache: FeatureCache, pos_weight, device) -> FeatureCache:
    """Device-resident view of ``cache`` whose pos_weight is the criterion's (the loss weights belong to the criterion in
    the reference, aux_models.py:131-133)."""
    c = cache.to(device)
    if pos_weight is None:
        return c
    w = torch.as_tensor(np.asarray(pos_weight, np.float32)).to(device)
    if w.shape != c.pos_weight.shape:
        raise ValueError(f"pos_weight has {tuple(w.shape)} entries, the targets have {c.labels.shape[1]} classes")
    return FeatureCache(c.ske_cat, c.rgb_cat, c.labels, widths=WIDTHS, pos_weight=w)


def _multilabel_cache_of(loader, what):
    c = _feature_cache_of(loader, what)
    if not c.multilabel or c.widths != WIDTHS:
        raise TypeError(f"dataloaders['{what}'].dataset must be a text/image cache with multi-hot targets "
                        "(mfas_b200.mmimdb_searchable.text_image_cache)")
    return c


def _final_f1(best):
    """Host-side epilogue of train_mmimdb_track_f1 for one candidate (train_searchable/mmimdb.py:131-136).  The device loop
    tracks the strict-'>' best dev F1 starting from init_f1 (mfas_run_args.best_acc_init) and snapshots only epochs that beat
    it, so the rolled-back weights and the returned score always belong together; a NaN train loss makes the reference return
    the best F1 seen BEFORE that epoch (:105-109) -- after a NaN every later dev F1 is 0 (sigmoid(nan) > 0.3 is False), so the
    device-side best is already that value.  A NaN best (only possible through a NaN init_f1) is reported as 0.0 (:133-134)."""
    b = float(best)
    return 0.0 if b != b else b


class _MMIMDBSpec(TrainerSpec):
    widths = WIDTHS

    def cache_of(self, loader, what):
        return _multilabel_cache_of(loader, what)

    def flags(self, args):
        return _flags(args)

    def check(self, args, flags, preaccuracies):
        if flags & _lib.FLAG_MULTITASK:
            raise ValueError("the MM-IMDB head has no multitask variant")

    def best_init(self, preaccuracies, idx):
        return float(preaccuracies[idx]) if preaccuracies else 0.0          # init_f1=preaccuracies[idx], ntu_searchable.py:88

    def result(self, best):
        return _final_f1(best)

    def load_backbones(self, rmode, args):
        pass                                                               # cached taps: there are no backbones to load

    def log_epochs(self, stats, n_train, n_dev):
        for e in range(stats.shape[0]):                                    # train_searchable/mmimdb.py:102-103
            print('epoch #{} {} F1: {:.4f} '.format(e, 'dev', float(stats[e, 3]) / n_dev))

    def vid_len(self, args):
        return 32


def train_sampled_models(sampled_configurations, searchable_type, dataloaders,
                         args, device,
                         return_model=[], premodels=[], preaccuracies=[],
                         train_only_central_params=True,
                         state_dict=dict()):
    """Train every sampled configuration; returns its best dev F1-samples, in input order (and the models when
    ``return_model``).  Signature of /root/reference/models/search/ntu_searchable.py:23-27; recipe of that function with
    the MM-IMDB loop: Adam(lr=eta_max, weight_decay=1e-4), per-batch cosine LR, ``args.epochs`` x (train pass, dev pass),
    strict-'>' best-dev tracking (starting from ``preaccuracies[idx]`` = init_f1) and rollback.  The body is the one of the NTU
    trainer (mfas_b200.ntu_searchable.train_sampled): all candidates of the call train concurrently (one at a time only when
    ``args.weightsharing`` chains them through ``state_dict``), sharded over torch.distributed ranks or, with
    ``args.fanout_gpus``, over the devices of this process."""
    spec = _MMIMDBSpec()
    spec.own_class = Searchable_Text_Image_Net
    return train_sampled(spec, train_sampled_models, sampled_configurations, searchable_type, dataloaders, args, device,
                         return_model, premodels, preaccuracies, state_dict)


def train_mmimdb_track_f1(model, criterion, optimizer, scheduler, dataloaders, dataset_sizes,
                          device=None, num_epochs=200, verbose=False, init_f1=0.0, th_fscore=0.3):
    """num_epochs x (train pass, dev pass) with best-dev rollback; returns the best dev F1-samples as a python float
    (/root/reference/models/search/train_searchable/mmimdb.py:14-136).  ``criterion`` is a
    ``WeightedCrossEntropyWithLogits`` (its ``.w`` are the positive-class weights); the LR of every batch comes from
    ``scheduler`` as in the reference."""
    from .train_ntu import _adam_hparams
    if abs(float(th_fscore) - TH_FSCORE) > 1e-12:
        raise NotImplementedError(f"th_fscore={th_fscore}: the CUDA head thresholds at {TH_FSCORE} (the reference default)")
    if not hasattr(criterion, "w"):
        raise TypeError("criterion must be a WeightedCrossEntropyWithLogits (carrying the pos_weight vector as .w)")
    net = model.module if isinstance(model, torch.nn.DataParallel) else model
    B = int(getattr(dataloaders['train'], 'batch_size', None) or net.args.batchsize)
    g = net.native(device, batch_max=B)                    # batch <= 64 puts inner_repr % 64 == 0 on the tensor-core engine
    b1, b2, eps, wd, _ = _adam_hparams(optimizer)
    g.set_adam(b1, b2, eps, wd)
    train_c = _with_pos_weight(_multilabel_cache_of(dataloaders['train'], 'train'), criterion.w, g.device)
    dev_c = _with_pos_weight(_multilabel_cache_of(dataloaders['dev'], 'dev'), criterion.w, g.device)
    n_train, n_dev = len(train_c), len(dev_c)
    steps = math.ceil(n_train / B)
    slot = net._slot
    named = dict(net.named_parameters())
    t0 = 0
    for name, p in named.items():                          # optimiser state in -> arenas (a fresh Adam has none)
        if name not in g.slots[slot]:
            continue
        st = optimizer.state.get(p, None)
        if st:
            g.view(slot, name, "m").copy_(st["exp_avg"].reshape(g.view(slot, name, "m").shape))
            g.view(slot, name, "v").copy_(st["exp_avg_sq"].reshape(g.view(slot, name, "v").shape))
            t0 = int(st["step"])
        else:
            g.view(slot, name, "m").zero_()
            g.view(slot, name, "v").zero_()
    g.adam_t = t0
    # a NaN init_f1 is never beaten; with num_epochs == 1 the reference then trains one more epoch before giving up (:22-24,
    # :123-129: 'Recording a NaN F1, training for one more epoch.') and ends on the incoming weights with a score of 0.0
    nan_init = float(init_f1) != float(init_f1)
    if nan_init and num_epochs == 1:
        print('Recording a NaN F1, training for one more epoch.')
        num_epochs = 2
    lrs = []
    for _ in range(num_epochs):
        if not is_per_batch_cosine(scheduler):
            scheduler.step()                               # mmimdb.py:37-38 (epoch-level schedulers)
            lrs += [optimizer.param_groups[0]["lr"]] * steps
        else:
            for _ in range(steps):                         # mmimdb.py:76-78
                lrs.append(scheduler.step())
    if is_per_batch_cosine(scheduler) and lrs:
        scheduler.update_optimizer(optimizer)
    k_tr = _reserve_passes(dataloaders['train'], num_epochs)
    k_dv = _reserve_passes(dataloaders['dev'], num_epochs)
    ptr = pass_orders(dataloaders['train'], k_tr, num_epochs, n_train)[None]
    pdv = pass_orders(dataloaders['dev'], k_dv, num_epochs, n_dev)[None]
    stats, best, _ = g.train_run(train_c, dev_c, ptr, pdv, lrs, num_epochs, B, b1, b2, best_init=[float(init_f1)])
    stats, best = stats.cpu(), best.cpu()
    g.check()
    if verbose:
        for e in range(num_epochs):                        # mmimdb.py:102-103
            print('epoch #{} {} F1: {:.4f} '.format(e, 'dev', float(stats[0, e, 3]) / n_dev))
    for name, p in named.items():                          # arenas -> optimiser state
        if name in g.slots[slot] and p.requires_grad and any(p is q for grp in optimizer.param_groups for q in grp["params"]):
            if name.startswith("alphas") and not getattr(net.args, "alphas", False):
                continue
            optimizer.state[p] = {"step": torch.tensor(float(g.adam_t)),
                                  "exp_avg": g.view(slot, name, "m").clone().reshape(p.shape),
                                  "exp_avg_sq": g.view(slot, name, "v").clone().reshape(p.shape)}
    model.train(False)
    train_mmimdb_track_f1.last_stats = stats[0]
    return _final_f1(best[0])
