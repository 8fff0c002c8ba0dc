"""Drop-in for /root/reference/models/search/train_searchable/ntu.py (single-model entry points used
by main_found_ntu.py): same signatures and return values, the loops run inside the CUDA library.
"""
from __future__ import annotations

import math

import torch

from .ntu_searchable import _feature_cache_of, _reserve_passes, pass_orders
from .scheduler import is_per_batch_cosine


def _adam_hparams(optimizer):
    if not isinstance(optimizer, torch.optim.Adam):
        raise TypeError("the native loop implements torch.optim.Adam (the optimiser the reference uses); got "
                        + type(optimizer).__name__)
    g = optimizer.param_groups[0]
    if g.get("amsgrad", False) or g.get("maximize", False):
        raise NotImplementedError("amsgrad / maximize are not used by the reference and not built")
    return g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], g["lr"]


def _check_multitask(net, multitask, dataloaders, splits):
    """The model returns the 3-tuple iff args.multitask (ntu_searchable.py:244-247) and the loop indexes it iff
    ``multitask`` (train_searchable/ntu.py:53-61): a mismatch fails in the reference too."""
    if bool(getattr(net.args, "multitask", False)) != bool(multitask):
        raise TypeError(f"multitask={bool(multitask)} but the model was built with args.multitask="
                        f"{bool(getattr(net.args, 'multitask', False))} (the reference loop would index / reduce the wrong type)")
    if multitask:
        for sp in splits:
            if _feature_cache_of(dataloaders[sp], sp).logit_rgb is None:
                raise ValueError(f"multitask needs cached backbone logits in dataloaders['{sp}'].dataset "
                                 "(FeatureCache(logit_rgb=, logit_ske=))")


def train_ntu_track_acc(model, criteria, optimizer, scheduler, dataloaders, dataset_sizes,
                        device=None, num_epochs=200, verbose=False, multitask=False):
    """num_epochs x (train pass, dev pass) with best-dev rollback; returns the best dev accuracy
    (/root/reference/models/search/train_searchable/ntu.py:14-89).  ``criteria`` must be cross-entropy
    (what every caller passes); the LR of every batch comes from ``scheduler`` exactly as in the
    reference (per-batch for LRCosineAnnealingScheduler, otherwise stepped once per epoch)."""
    return _train_track_acc(train_ntu_track_acc, model, optimizer, scheduler, dataloaders, device, num_epochs, multitask, with_loss=True)


def _train_track_acc(entry, model, optimizer, scheduler, dataloaders, device, num_epochs, multitask, with_loss):
    """Body shared with train_avmnist_track_acc (whose loop prints '<phase> Acc: ...' only, train_searchable/avmnist.py:79)."""
    net = model.module if isinstance(model, torch.nn.DataParallel) else model
    _check_multitask(net, multitask, dataloaders, ('train', 'dev'))
    B = int(getattr(dataloaders['train'], 'batch_size', None) or net.args.batchsize)
    # size the group by the loop's batch: batches of at most 64 rows get the persistent streaming kernels, the fused chain
    # and the wide eval pass (a group sized for MFAS_MAX_BATCH rows runs the grid-indexed kernels)
    g = net.native(device, batch_max=B)
    b1, b2, eps, wd, lr0 = _adam_hparams(optimizer)
    g.set_adam(b1, b2, eps, wd)
    train_c = _feature_cache_of(dataloaders['train'], 'train').to(g.device)
    dev_c = _feature_cache_of(dataloaders['dev'], 'dev').to(g.device)
    n_train, n_dev = len(train_c), len(dev_c)
    steps = math.ceil(n_train / B)

    # optimiser state in -> arenas (a fresh Adam has none)
    named = dict(net.named_parameters())
    slot = net._slot
    t0 = 0
    for name, p in named.items():
        if name not in g.slots[slot]:
            continue
        st = optimizer.state.get(p, None)
        if st:
            g.view(slot, name, "m").copy_(st["exp_avg"].reshape(g.view(slot, name, "m").shape))
            g.view(slot, name, "v").copy_(st["exp_avg_sq"].reshape(g.view(slot, name, "v").shape))
            t0 = int(st["step"])
        else:                                                 # a fresh optimiser starts from zero moments, whatever an
            g.view(slot, name, "m").zero_()                   # earlier stage left in the arenas (main_found_ntu.py:133)
            g.view(slot, name, "v").zero_()
    g.adam_t = t0

    lrs = []
    for _ in range(num_epochs):
        if not is_per_batch_cosine(scheduler):
            scheduler.step()                                  # ntu.py:25-26 (epoch-level schedulers)
            lrs += [optimizer.param_groups[0]["lr"]] * steps
        else:
            for _ in range(steps):                            # ntu.py:65-67
                lrs.append(scheduler.step())
    if is_per_batch_cosine(scheduler) and lrs:
        scheduler.update_optimizer(optimizer)

    k_tr = _reserve_passes(dataloaders['train'], num_epochs)
    k_dv = _reserve_passes(dataloaders['dev'], num_epochs)
    ptr = pass_orders(dataloaders['train'], k_tr, num_epochs, n_train)[None]
    pdv = pass_orders(dataloaders['dev'], k_dv, num_epochs, n_dev)[None]
    stats, best, _ = g.train_run(train_c, dev_c, ptr, pdv, lrs, num_epochs, B, b1, b2)
    stats, best = stats.cpu(), best.cpu()
    for e in range(num_epochs):                               # ntu.py:78-79 / avmnist.py:79 print unconditionally
        if with_loss:
            print('{} Loss: {:.4f} Acc: {:.4f}'.format('train', stats[0, e, 0] / n_train, stats[0, e, 1] / n_train))
            print('{} Loss: {:.4f} Acc: {:.4f}'.format('dev', stats[0, e, 2] / n_dev, stats[0, e, 3] / n_dev))
        else:
            print('{} Acc: {:.4f}'.format('train', stats[0, e, 1] / n_train))
            print('{} Acc: {:.4f}'.format('dev', stats[0, e, 3] / n_dev))

    # arenas -> optimiser state, so a caller inspecting / reusing the optimiser sees torch's layout
    for name, p in named.items():
        if name in g.slots[slot] and p.requires_grad and any(p is q for grp in optimizer.param_groups for q in grp["params"]):
            if name.startswith("alphas") and not getattr(net.args, "alphas", False):
                continue                                       # no grad -> Adam never creates state
            optimizer.state[p] = {"step": torch.tensor(float(g.adam_t)),
                                  "exp_avg": g.view(slot, name, "m").clone().reshape(p.shape),
                                  "exp_avg_sq": g.view(slot, name, "v").clone().reshape(p.shape)}
    model.train(False)
    entry.last_stats = stats[0]
    return best[0].clone()


def test_ntu_track_acc(model, dataloaders, dataset_sizes, device=None, multitask=False):
    """Eval-mode accuracy over dataloaders['test'] (train_searchable/ntu.py:92-125)."""
    net = model.module if isinstance(model, torch.nn.DataParallel) else model
    _check_multitask(net, multitask, dataloaders, ('test',))
    model.train(False)
    g = net.native(device)
    test_c = _feature_cache_of(dataloaders['test'], 'test').to(g.device)
    B = int(getattr(dataloaders['test'], 'batch_size', None) or net.args.batchsize)
    if B > g.batch_max:                                     # the group was sized by a training loop with a smaller batch
        g = net.native(device, batch_max=B)
    out = g.eval_pass(test_c, B).cpu()
    return (out[0, 1] / dataset_sizes['test']).clone()
