"""Feature-cache builder: the step before the hot path.

The reference runs both frozen backbones on every batch of every candidate and pools the selected taps on the fly
(/root/reference/models/search/ntu_searchable.py:206-225).  Here they run ONCE over a split: the 4 + 4 taps every
configuration can select are globally pooled by a CUDA kernel (``mfas_global_pool``, csrc/kernels_pool.cuh) straight into
their column slices of the cache matrices, and the two backbone classifier outputs are kept for ``args.multitask``.

Semantic difference to the reference: the backbones run in ``eval()`` mode under ``no_grad`` -- the reference
leaves them in whatever mode ``model.train(phase == 'train')`` sets while never updating their weights, so its BatchNorm /
Dropout layers inside the backbones behave differently between the train and dev phases; a cache holds one deterministic
feature per sample.

    cache = build_feature_cache(rgbnet, skenet, loader, "cuda:0", vid_len_ske=args.vid_len[1])

``loader`` yields the reference's batch dicts {'rgb', 'ske', 'label'} (/root/reference/datasets/ntu.py:84-87) in dataset
order (``shuffle=False``); ``rgbnet`` / ``skenet`` are the reference's ``Visual`` / ``Skeleton`` modules
(/root/reference/models/central/ntu.py:17-50, :55-183) or anything with the same output structure.  Other tap sets pass
their own ``taps`` selector and ``widths`` (e.g. ``mmimdb_taps`` with ``mmimdb_searchable.WIDTHS``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .cache import D_RGB, FeatureCache, ske_widths


def ntu_taps(rgb_out, ske_out):
    """The tensors ``Searchable_Skeleton_Image_Net.forward`` keeps from the backbone outputs (ntu_searchable.py:211-217):
    -> (4 skeleton taps, 4 visual taps, skeleton logits, visual logits)."""
    visual_logits = rgb_out[-1]
    visual = list(rgb_out[-5:-1])
    ske_feats, ske_logits = ske_out
    return list(ske_feats[-4:]), visual, ske_logits, visual_logits


def mmimdb_taps(image_out, text_out):
    """MM-IMDB: text taps = the two Maxout hidden layers of ``MaxOut_MLP.forward`` -> (o1, o3, o5)
    (/root/reference/models/central/mm_imdb.py:189-196); image taps = every output of the pooled VGG but its logits (last)."""
    return [text_out[0], text_out[1]], list(image_out[:-1])[-4:], text_out[2], image_out[-1]


def global_pool_into(x: torch.Tensor, out2d: torch.Tensor):
    """out2d[b, c] = mean of x[b, c, ...] (GlobalPooling2D, aux_models.py:58-64) by the CUDA kernel; ``out2d`` may be a
    column slice of a wider matrix (unit column stride)."""
    if x.device.type != "cuda" or out2d.device != x.device:
        raise RuntimeError("global_pool_into runs on CUDA tensors only (no CPU fallback)")
    if x.dim() < 2 or out2d.dim() != 2 or out2d.shape != (x.shape[0], x.shape[1]) or out2d.stride(1) != 1:
        raise ValueError(f"tap {tuple(x.shape)} does not fit the cache slice {tuple(out2d.shape)} (strides {out2d.stride()})")
    if x.dtype != torch.float32 or out2d.dtype != torch.float32:
        raise TypeError("taps and caches are fp32")
    x = x.contiguous()
    B, Cn = x.shape[0], x.shape[1]
    S = x.numel() // max(B * Cn, 1)
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(_lib.lib().mfas_global_pool(x.device.index or 0, x.data_ptr(), B, Cn, S, out2d.data_ptr(), out2d.stride(0), stream))
    return out2d


@torch.no_grad()
def build_feature_cache(rgbnet, skenet, loader, device, vid_len_ske=32, with_logits=False, taps=ntu_taps, widths=None,
                        keys=('rgb', 'ske', 'label'), n_rows=None, pos_weight=None) -> FeatureCache:
    """Run the two backbones once over ``loader`` and return the device-resident FeatureCache of the split.

    ``keys`` names the (second-modality input, first-modality input, label) entries of a batch dict; the networks are
    called as ``rgbnet(batch[keys[0]])`` and ``skenet(batch[keys[1]])``.  Labels may be int64 class ids or fp32 multi-hot
    rows (then pass ``pos_weight``)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("the feature cache is built on a CUDA device (no CPU fallback)")
    return _build(rgbnet, skenet, loader, device, global_pool_into, vid_len_ske, with_logits, taps, widths, keys, n_rows, pos_weight)


def _build(rgbnet, skenet, loader, device, pool, vid_len_ske, with_logits, taps, widths, keys, n_rows, pos_weight):
    """The builder's host logic; ``pool(tap, cache_slice)`` is the CUDA kernel (tests substitute the oracle to exercise
    this loop without a GPU)."""
    d0, d1 = (ske_widths(vid_len_ske), D_RGB) if widths is None else (tuple(widths[0]), tuple(widths[1]))
    N = int(n_rows if n_rows is not None else len(loader.dataset))
    first = torch.empty(N, sum(d0), dtype=torch.float32, device=device)
    second = torch.empty(N, sum(d1), dtype=torch.float32, device=device)
    labels, logit_first, logit_second = None, None, None
    was = (rgbnet.training, skenet.training)
    rgbnet.eval(); skenet.eval()
    row = 0
    try:
        for batch in loader:
            x1, x0, y = (batch[k].to(device, non_blocking=True) for k in keys)
            t0, t1, l0, l1 = taps(rgbnet(x1), skenet(x0))
            B = y.shape[0]
            if row + B > N:
                raise ValueError(f"the loader yields more than the {N} rows the cache was sized for")
            if len(t0) != len(d0) or len(t1) != len(d1):
                raise ValueError(f"expected {len(d0)} + {len(d1)} taps, the backbones returned {len(t0)} + {len(t1)}")
            for taps_m, widths_m, cat in ((t0, d0, first), (t1, d1, second)):
                off = 0
                for t, w in zip(taps_m, widths_m):
                    if t.shape[0] != B or t.shape[1] != w:
                        raise ValueError(f"tap of shape {tuple(t.shape)} where [{B}, {w}, ...] was expected")
                    pool(t.float(), cat[row:row + B, off:off + w])
                    off += w
            if labels is None:
                labels = torch.empty((N,) + tuple(y.shape[1:]), dtype=y.dtype, device=device)
                if with_logits:
                    logit_first = torch.empty(N, l0.shape[1], dtype=torch.float32, device=device)
                    logit_second = torch.empty(N, l1.shape[1], dtype=torch.float32, device=device)
            labels[row:row + B] = y
            if with_logits:
                logit_first[row:row + B] = l0
                logit_second[row:row + B] = l1
            row += B
    finally:
        rgbnet.train(was[0]); skenet.train(was[1])
    if row != N:
        raise ValueError(f"the loader yielded {row} rows, the cache was sized for {N}")
    if pos_weight is not None:
        pos_weight = torch.as_tensor(pos_weight, dtype=torch.float32).to(device)
    return FeatureCache(first, second, labels, vid_len_ske, logit_second, logit_first, widths=widths, pos_weight=pos_weight)
