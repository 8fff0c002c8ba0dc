"""Multi-GPU: candidates are independent, so the only traffic is (1) one broadcast of the feature
cache and (2) one gather of the accuracies per train_sampled_models call.
One process per GPU (torchrun); NCCL over NVLink/NVSwitch on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as td

from .cache import FeatureCache


def world():
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def shard(n_items: int, rank: int, world_size: int):
    """Round-robin ownership: item j belongs to rank j % world_size.  Placement never changes a
    result (batch orders depend on (seed, candidate, epoch) only)."""
    return [j for j in range(n_items) if j % world_size == rank]


def my_share(n_items: int):
    r, w = world()
    return shard(n_items, r, w)


def _comm_device():
    if td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def sync_call_inputs(sampled_configurations, seed: int):
    """Make every rank train the SAME call: rank 0's configuration list and seed win.

    The reference's driver samples with unseeded numpy (models/search/tools.py:53 ``np.random.choice``), so the ranks of an
    unmodified ``main_searchable_ntu.py`` under torchrun draw different lists; sharding "candidate j -> rank j % world" and
    summing disjoint result vectors is only meaningful when the lists agree.  One small object broadcast per call.
    Returns (configurations, seed) as rank 0 holds them."""
    r, w = world()
    if w == 1:
        return sampled_configurations, seed
    import numpy as np
    box = [([np.asarray(c).tolist() for c in sampled_configurations], int(seed)) if r == 0 else None]
    td.broadcast_object_list(box, src=0)
    confs, seed0 = box[0]
    if r != 0:
        mine = [np.asarray(c).tolist() for c in sampled_configurations]
        if mine != confs:
            import warnings
            warnings.warn(f"mfas_b200: rank {r} sampled a different configuration list than rank 0 (unseeded driver RNG); "
                          "training rank 0's list", RuntimeWarning)
    return [np.asarray(c) for c in confs], seed0


def gather_results(values: torch.Tensor, n_items: int) -> torch.Tensor:
    """values: [n_items, ...] with zeros in the slots other ranks own -> the full tensor on every
    rank, in input order.  (A sum all-reduce of disjointly-filled vectors is an order-preserving
    all-gather for any ragged split.)"""
    r, w = world()
    if w == 1:
        return values
    buf = values.to(_comm_device()).contiguous()
    td.all_reduce(buf, op=td.ReduceOp.SUM)
    return buf.cpu()


def broadcast_cache(cache, device, src: int = 0) -> FeatureCache:
    """Rank ``src`` holds ``cache`` (host or device); every rank returns a device-resident copy.
    The payload (NTU: 0.46 GB for train+dev) crosses NVLink once and is reused by every later call."""
    r, w = world()
    device = torch.device(device)
    if w == 1:
        return cache.to(device)
    cdev = device if td.get_backend() == "nccl" else torch.device("cpu")
    meta = [None]
    if r == src:       # labels: int64 [N] class ids, or fp32 [N, C] multi-hot targets + pos_weight [C] (MM-IMDB)
        meta = [(len(cache), cache.ske_cat.shape[1], cache.rgb_cat.shape[1], cache.vid_len_ske, tuple(cache.labels.shape),
                 cache.widths, None if cache.logit_rgb is None else int(cache.logit_rgb.shape[1]))]
    td.broadcast_object_list(meta, src=src)
    n, ws, wr, vl, lshape, widths, n_logit = meta[0]
    multilabel = len(lshape) == 2
    if r == src:
        c = cache.to(cdev)
        ske, rgb, lab, pw, lrgb, lske = c.ske_cat, c.rgb_cat, c.labels, c.pos_weight, c.logit_rgb, c.logit_ske
    else:
        lrgb = torch.empty(n, n_logit, dtype=torch.float32, device=cdev) if n_logit else None      # cached backbone logits (multitask)
        lske = torch.empty(n, n_logit, dtype=torch.float32, device=cdev) if n_logit else None
        ske = torch.empty(n, ws, dtype=torch.float32, device=cdev)
        rgb = torch.empty(n, wr, dtype=torch.float32, device=cdev)
        lab = torch.empty(lshape, dtype=torch.float32 if multilabel else torch.int64, device=cdev)
        pw = torch.empty(lshape[1], dtype=torch.float32, device=cdev) if multilabel else None
    for t in (ske, rgb, lab) + ((pw,) if multilabel else ()) + ((lrgb, lske) if n_logit else ()):
        td.broadcast(t, src=src)
    return FeatureCache(ske.to(device), rgb.to(device), lab.to(device), vl,
                        lrgb.to(device) if n_logit else None, lske.to(device) if n_logit else None, widths=widths,
                        pos_weight=pw.to(device) if multilabel else None)
