"""CPU oracle for the MFAS candidate-training hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy/fp32 restatement of the reference algorithm.  It is the checker
for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``mfas_b200``) never does -- it fails loudly when the CUDA library is missing.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this restatement against
fixtures under ``tests/golden/`` that were produced by executing the unmodified reference
(``/root/reference``, commit 1e9a715) with ``tests/golden/gen_golden.py``.

Reference lines restated (all relative to /root/reference):

* feature widths ............ models/search/ntu_searchable.py:288-296
* fusion forward ............ models/search/ntu_searchable.py:206-247, :258-286
* alpha gate ................ models/auxiliary/aux_models.py:94-111
* loss / preds .............. models/search/train_searchable/ntu.py:53-61
* epoch loop / best rollback  models/search/train_searchable/ntu.py:14-89
* test loop ................. models/search/train_searchable/ntu.py:92-125
* cosine LR w/ restarts ..... models/auxiliary/scheduler.py:12-46
* Adam(+L2) ................. models/search/ntu_searchable.py:65 (torch.optim.Adam, torch 2.11
                              ``_single_tensor_adam``: lerp for exp_avg, addcmul for exp_avg_sq,
                              denom = sqrt(v)/sqrt(bc2) + eps, step = lr/bc1)
* candidate loop ............ models/search/ntu_searchable.py:23-102
* weight sharing ............ models/search/ntu_searchable.py:74-75, :91-92, :123-174

The backward pass is hand derived (the reference uses autograd); it is the derivation the
CUDA kernels implement, so the oracle validates the derivation as well as the arithmetic.
"""
from __future__ import annotations

import contextlib
import copy
import math

import numpy as np

F32 = np.float32            # working precision of the restatement (see precision() below)


@contextlib.contextmanager
def precision(dtype):
    """Run the restatement in another working precision (tests use float64 as ground truth to measure
    how much of a gradient is fp32 rounding noise: BatchNorm over a nearly constant unit multiplies
    rounding errors by 1/sqrt(var+eps), up to 316x, twice, whatever the implementation)."""
    global F32
    old, F32 = F32, dtype
    try:
        yield
    finally:
        F32 = old
GEMM_SPLITS = 1             # see summation_order() below


@contextlib.contextmanager
def summation_order(splits):
    """Evaluate the fusion-step products x W^T as ``splits`` partial products over column ranges, added in order (what any
    split-K GEMM does) -- the same arithmetic in another, equally valid, fp32 summation order.  tests/golden/noise_floor.py
    uses it to measure how far two correct fp32 implementations of the reference algorithm may drift apart along a trajectory
    (the rounding of a pre-activation decides e.g. on which side of the ReLU kink it falls)."""
    global GEMM_SPLITS
    old, GEMM_SPLITS = GEMM_SPLITS, int(splits)
    try:
        yield
    finally:
        GEMM_SPLITS = old


def _xwt(x, W):
    if GEMM_SPLITS <= 1:
        return x @ W.T
    K = x.shape[1]
    edges = [K * i // GEMM_SPLITS for i in range(GEMM_SPLITS + 1)]
    acc = (x[:, edges[0]:edges[1]] @ W[:, edges[0]:edges[1]].T).astype(F32)
    for a, b in zip(edges[1:-1], edges[2:]):
        acc = (acc + (x[:, a:b] @ W[:, a:b].T).astype(F32)).astype(F32)
    return acc


D_RGB = (512, 1024, 2048, 2048)            # ntu_searchable.py:292
BN_EPS = 1e-5                              # torch.nn.BatchNorm1d default
BN_MOMENTUM = 0.1
LEAKY_SLOPE = 0.01                         # torch.nn.LeakyReLU default (ntu_searchable.py:272)

ACT_RELU, ACT_SIGMOID, ACT_LRELU = 0, 1, 2


def d_ske(vid_len_ske: int = 32):
    """ntu_searchable.py:291 -- widths of the last four skeleton taps."""
    return (128, 256, 32 * vid_len_ske, 512)


def layer_in_features(conf, H, vid_len_ske=32):
    """K_l = D_ske[i] + D_rgb[j] + (l>0)*H  (ntu_searchable.py:261-264)."""
    ds = d_ske(vid_len_ske)
    return [ds[int(c[0])] + D_RGB[int(c[1])] + (H if l > 0 else 0) for l, c in enumerate(conf)]


# --------------------------------------------------------------------------------------
# scheduler (models/auxiliary/scheduler.py:12-46)
# --------------------------------------------------------------------------------------
class CosineRestartLR:
    """Warm-restart cosine schedule evaluated once per *batch* (scheduler.py:29-40).

    Note the reference quirk: ``Tcur`` is a float ``iteration/nbpe`` and a restart fires only
    when eta falls to within 1e-10 of eta_min, i.e. when Tcur/Ti hits an odd integer exactly.
    """

    def __init__(self, eta_max, eta_min, Ti, Tm, nbpe):
        self.eta_max, self.eta_min = eta_max, eta_min
        self.Ti, self.Tm, self.nbpe = Ti, Tm, nbpe
        self.Tcur = 0.0
        self.it = 0.0
        self.eta = eta_max

    def step(self):
        self.Tcur = self.it / self.nbpe
        self.it += 1.0
        self.eta = self.eta_min + 0.5 * (self.eta_max - self.eta_min) * (
            1 + np.cos(np.pi * self.Tcur / self.Ti))
        eta = self.eta
        if eta <= self.eta_min + 1e-10:
            self.Tcur = 0
            self.Ti = self.Ti * self.Tm
            self.it = 0
        return eta


# --------------------------------------------------------------------------------------
# counter-based dropout mask shared by oracle and CUDA path (own design; the reference
# uses torch's Philox stream which cannot be reproduced -- see DESIGN.md "Dropout").
# --------------------------------------------------------------------------------------
def _mix32(x):
    x = np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def dropout_keep_mask(seed, cand, step, layer, nrows, H, p):
    """keep[b, h] = u(b,h) >= p with u from a 2-round 32-bit mixer over
    (seed, candidate, adam step, layer, b*H+h).  Mirrors ``mfas_dropout_keep`` in
    mfas_b200/csrc/common.cuh bit for bit."""
    idx = (np.arange(nrows, dtype=np.uint64)[:, None] * np.uint64(H)
           + np.arange(H, dtype=np.uint64)[None, :])
    k = _mix32(np.uint64(seed & 0xFFFFFFFF) ^ np.uint64(0x9E3779B9))
    k = _mix32(k + np.uint64(cand) * np.uint64(0x85EBCA6B) & np.uint64(0xFFFFFFFF))
    k = _mix32((k + np.uint64(step)) & np.uint64(0xFFFFFFFF))
    k = _mix32((k + np.uint64(layer) * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF))
    r = _mix32((k + idx) & np.uint64(0xFFFFFFFF))
    u = (r >> np.uint64(8)).astype(np.float32) * F32(1.0 / 16777216.0)
    return u >= F32(p)


# --------------------------------------------------------------------------------------
# the fusion head
# --------------------------------------------------------------------------------------
def _act(z, kind):
    if kind == ACT_RELU:
        return np.maximum(z, F32(0))
    if kind == ACT_SIGMOID:
        return (F32(1) / (F32(1) + np.exp(-z, dtype=F32))).astype(F32)
    if kind == ACT_LRELU:
        return np.where(z > 0, z, F32(LEAKY_SLOPE) * z).astype(F32)
    raise ValueError(kind)


def _act_grad(a, kind):
    """phi'(z) expressed through the activated value a = phi(z)."""
    if kind == ACT_RELU:
        return (a > 0).astype(F32)
    if kind == ACT_SIGMOID:
        return (a * (F32(1) - a)).astype(F32)
    if kind == ACT_LRELU:
        return np.where(a > 0, F32(1), F32(LEAKY_SLOPE)).astype(F32)
    raise ValueError(kind)


class FusionHead:
    """State and arithmetic of one candidate (Searchable_Skeleton_Image_Net minus backbones).

    ``state`` uses the reference state_dict key names:
    ``fusion_layers.{l}.0.{weight,bias}``, ``fusion_layers.{l}.2.{weight,bias,running_mean,
    running_var,num_batches_tracked}``, ``central_classifier.{weight,bias}``,
    ``alphas.{l}.alpha_x``.
    """

    def __init__(self, conf, H, C, state, batchnorm=True, drpt=0.0, alphas=False,
                 vid_len_ske=32, dropout_seed=0, cand_index=0, plain=False, widths=None):
        self.conf = np.asarray(conf, dtype=np.int64).reshape(-1, 3)
        self.L = len(self.conf)
        self.H, self.C = int(H), int(C)
        self.bn = bool(batchnorm)
        self.drpt = float(drpt)
        self.use_alphas = bool(alphas)
        self.ds, self.dr = widths if widths is not None else (d_ske(vid_len_ske), D_RGB)   # tap widths of the two modalities
        self.K = [self.ds[int(c[0])] + self.dr[int(c[1])] + (self.H if l > 0 else 0) for l, c in enumerate(self.conf)]
        self.dropout_seed, self.cand_index = dropout_seed, cand_index
        if self.drpt < 1e-10 and not self.bn and not plain:
            # ntu_searchable.py:274-284 has no branch for this combination (plain: the AV-MNIST recipe Linear -> act,
            # avmnist_searchable.py:276-285, oracle/avmnist_oracle.py)
            raise UnboundLocalError("no layer recipe for drpt<1e-10 and batchnorm=False")
        self.state = {k: np.array(v, copy=True) for k, v in state.items()}
        self.adam = {}       # name -> (m, v)
        self.t = 0           # Adam step counter

    # -- parameter bookkeeping -----------------------------------------------------
    def trainable_names(self):
        names = []
        for l in range(self.L):
            names += [f"fusion_layers.{l}.0.weight", f"fusion_layers.{l}.0.bias"]
            if self.bn:
                names += [f"fusion_layers.{l}.2.weight", f"fusion_layers.{l}.2.bias"]
        names += ["central_classifier.weight", "central_classifier.bias"]
        if self.use_alphas:      # alphas get a grad only when the gate is in the graph
            names += [f"alphas.{l}.alpha_x" for l in range(self.L)]
        return names

    # -- forward ---------------------------------------------------------------------
    def forward(self, ske_taps, rgb_taps, train, step_for_dropout=None):
        """ske_taps / rgb_taps: lists of 4 arrays [B, D].  Returns logits and a tape."""
        s = self.state
        B = ske_taps[0].shape[0]
        tape = []
        h = None
        for l, (i, j, act) in enumerate(self.conf):
            xs, xr = ske_taps[i].astype(F32), rgb_taps[j].astype(F32)
            gate = None
            if self.use_alphas:                               # aux_models.py:103-111
                alpha = s[f"alphas.{l}.alpha_x"].astype(F32)[0]
                sg = F32(1) / (F32(1) + np.exp(-alpha, dtype=F32))
                gate = sg
                xs_in, xr_in = xs * sg, xr * (F32(1) - sg)
            else:
                xs_in, xr_in = xs, xr
            parts = (xs_in, xr_in) if l == 0 else (xs_in, xr_in, h)   # ntu_searchable.py:235-239
            x = np.concatenate(parts, axis=1)
            W, b = s[f"fusion_layers.{l}.0.weight"], s[f"fusion_layers.{l}.0.bias"]
            z = (_xwt(x, W) + b).astype(F32)
            a = _act(z, act)
            rec = dict(x=x, z=z, a=a, xs=xs, xr=xr, gate=gate)
            out = a
            if self.bn:
                g_, be = s[f"fusion_layers.{l}.2.weight"], s[f"fusion_layers.{l}.2.bias"]
                if train:
                    if B <= 1:
                        raise ValueError("Expected more than 1 value per channel when training")
                    mu = a.mean(axis=0, dtype=F32)
                    var = ((a - mu) ** 2).mean(axis=0, dtype=F32)
                    rm, rv = s[f"fusion_layers.{l}.2.running_mean"], s[f"fusion_layers.{l}.2.running_var"]
                    s[f"fusion_layers.{l}.2.running_mean"] = (
                        F32(1 - BN_MOMENTUM) * rm + F32(BN_MOMENTUM) * mu).astype(F32)
                    s[f"fusion_layers.{l}.2.running_var"] = (
                        F32(1 - BN_MOMENTUM) * rv + F32(BN_MOMENTUM) * (var * F32(B / (B - 1.0)))).astype(F32)
                    s[f"fusion_layers.{l}.2.num_batches_tracked"] = (
                        s[f"fusion_layers.{l}.2.num_batches_tracked"] + 1)
                else:
                    mu = s[f"fusion_layers.{l}.2.running_mean"]
                    var = s[f"fusion_layers.{l}.2.running_var"]
                invstd = (F32(1) / np.sqrt(var + F32(BN_EPS))).astype(F32)
                ahat = ((a - mu) * invstd).astype(F32)
                out = (ahat * g_ + be).astype(F32)
                rec.update(ahat=ahat, invstd=invstd)
            if self.drpt > 1e-10 and train:
                keep = dropout_keep_mask(self.dropout_seed, self.cand_index,
                                         self.t if step_for_dropout is None else step_for_dropout,
                                         l, B, self.H, self.drpt)
                scale = F32(1.0 / (1.0 - self.drpt))
                out = np.where(keep, out * scale, F32(0)).astype(F32)
                rec.update(keep=keep, scale=scale)
            h = out
            tape.append(rec)
        Wc, bc = s["central_classifier.weight"], s["central_classifier.bias"]
        logits = (h @ Wc.T + bc).astype(F32)
        return logits, dict(layers=tape, h_last=h)

    # -- loss (train_searchable/ntu.py:53-58) ---------------------------------------------
    @staticmethod
    def ce_loss(logits, labels):
        mx = logits.max(axis=1, keepdims=True)
        sh = logits - mx
        lse = np.log(np.exp(sh, dtype=F32).sum(axis=1, keepdims=True, dtype=F32))
        logp = (sh - lse).astype(F32)
        B = logits.shape[0]
        loss = F32(-logp[np.arange(B), labels].mean(dtype=F32))
        return loss, logp

    # -- hand-derived backward -------------------------------
    def loss_and_dlogits(self, logits, labels):
        """mean cross-entropy and dlogits = (softmax - onehot) / B (a subclass swaps the head: oracle/mmimdb_oracle.py)"""
        B = logits.shape[0]
        loss, logp = self.ce_loss(logits, labels)
        dlog = np.exp(logp, dtype=F32)
        dlog[np.arange(B), labels] -= F32(1)
        return loss, (dlog / F32(B)).astype(F32)

    def backward(self, logits, labels, tape):
        s = self.state
        _, dlog = self.loss_and_dlogits(logits, labels)
        g = {}
        h_last = tape["h_last"]
        g["central_classifier.weight"] = (dlog.T @ h_last).astype(F32)
        g["central_classifier.bias"] = dlog.sum(axis=0, dtype=F32)
        dh = (dlog @ s["central_classifier.weight"]).astype(F32)
        for l in range(self.L - 1, -1, -1):
            i, j, act = self.conf[l]
            rec = tape["layers"][l]
            if "keep" in rec:
                dh = np.where(rec["keep"], dh * rec["scale"], F32(0)).astype(F32)
            if self.bn:
                ahat, invstd = rec["ahat"], rec["invstd"]
                gam = s[f"fusion_layers.{l}.2.weight"]
                g[f"fusion_layers.{l}.2.weight"] = (dh * ahat).sum(axis=0, dtype=F32)
                g[f"fusion_layers.{l}.2.bias"] = dh.sum(axis=0, dtype=F32)
                m1 = dh.mean(axis=0, dtype=F32)
                m2 = (dh * ahat).mean(axis=0, dtype=F32)
                da = (gam * invstd * (dh - m1 - ahat * m2)).astype(F32)
            else:
                da = dh
            dz = (da * _act_grad(rec["a"], act)).astype(F32)
            W = s[f"fusion_layers.{l}.0.weight"]
            g[f"fusion_layers.{l}.0.weight"] = (dz.T @ rec["x"]).astype(F32)
            g[f"fusion_layers.{l}.0.bias"] = dz.sum(axis=0, dtype=F32)
            Ds, Dr = self.ds[i], self.dr[j]
            if self.use_alphas:
                sg = rec["gate"]
                dxs = dz @ W[:, :Ds]
                dxr = dz @ W[:, Ds:Ds + Dr]
                dsg = (dxs * rec["xs"]).sum(dtype=F32) - (dxr * rec["xr"]).sum(dtype=F32)
                g[f"alphas.{l}.alpha_x"] = np.array([dsg * sg * (F32(1) - sg)], dtype=F32)
            if l > 0:
                dh = (dz @ W[:, Ds + Dr:]).astype(F32)       # features need no dX
        return g

    # -- Adam with coupled L2 (torch 2.11 _single_tensor_adam) ----------------------------
    def adam_step(self, grads, lr, weight_decay=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
        self.t += 1
        bc1 = 1.0 - beta1 ** self.t
        bc2 = 1.0 - beta2 ** self.t
        step_size = F32(lr / bc1)
        bc2_sqrt = F32(math.sqrt(bc2))
        for name in self.trainable_names():
            if name not in grads:
                continue
            p = self.state[name]
            gr = (grads[name] + F32(weight_decay) * p).astype(F32)
            if name not in self.adam:
                self.adam[name] = (np.zeros_like(p), np.zeros_like(p))
            m, v = self.adam[name]
            m = (m + F32(1 - beta1) * (gr - m)).astype(F32)                  # lerp_
            v = (v * F32(beta2) + F32(1 - beta2) * gr * gr).astype(F32)      # mul_.addcmul_
            denom = (np.sqrt(v) / bc2_sqrt + F32(eps)).astype(F32)
            self.state[name] = (p - step_size * (m / denom)).astype(F32)     # addcdiv_
            self.adam[name] = (m, v)

    def kink_margin(self):
        """Smallest |z| / max|z| over the piecewise-linear activations of the last train_step: when this
        is ~1e-6 a different (equally valid) fp32 summation order flips a ReLU derivative, so gradients
        of two correct implementations legitimately differ in that unit's row."""
        m = 1.0
        for l, rec in enumerate(self.last_tape["layers"]):
            if int(self.conf[l][2]) != ACT_SIGMOID:
                z = np.abs(rec["z"])
                m = min(m, float(z.min() / max(z.max(), 1e-30)))
        return m

    def train_step(self, ske_taps, rgb_taps, labels, lr):
        logits, tape = self.forward(ske_taps, rgb_taps, train=True)
        self.last_tape = tape
        loss, _ = self.loss_and_dlogits(logits, labels)
        grads = self.backward(logits, labels, tape)
        self.adam_step(grads, lr)
        return logits, loss, grads


# --------------------------------------------------------------------------------------
# epoch loop (train_searchable/ntu.py:14-89) and candidate loop (ntu_searchable.py:23-102)
# --------------------------------------------------------------------------------------
def _taps_of(split, rows):
    return [t[rows] for t in split["ske"]], [t[rows] for t in split["rgb"]], split["labels"][rows]


def multitask_loss_preds(logits, labels, aux):
    """train_searchable/ntu.py:59-61: loss = CE(out) + CE(visual) + CE(skeleton), preds = argmax of the sum.
    ``aux`` = (rgb backbone logits, ske backbone logits) of the batch rows, or None (single task, :53-58)."""
    loss, _ = FusionHead.ce_loss(logits, labels)
    if aux is None:
        return loss, logits.argmax(axis=1)
    lr, ls = (a.astype(F32) for a in aux)
    loss = F32(F32(loss + FusionHead.ce_loss(lr, labels)[0]) + FusionHead.ce_loss(ls, labels)[0])
    return loss, ((logits + lr).astype(F32) + ls).astype(F32).argmax(axis=1)


def _aux_of(split, rows, multitask):
    return (split["logit_rgb"][rows], split["logit_ske"][rows]) if multitask else None


def train_track_acc(head, sched, train_split, dev_split, batch, orders, num_epochs, log=None, multitask=False):
    """orders: callable (phase, epoch) -> row-index array for that pass.

    Returns (best_acc float64, per-epoch stats list).  Rolls ``head.state`` back to the best
    dev epoch, strict '>' (train_searchable/ntu.py:82-86).
    """
    best_state = copy.deepcopy(head.state)
    best_acc = 0.0
    stats = []
    for epoch in range(num_epochs):
        row = {}
        for phase, split in (("train", train_split), ("dev", dev_split)):
            order = np.asarray(orders(phase, epoch))
            n = len(order)
            run_loss, run_correct = 0.0, 0
            for s0 in range(0, n, batch):
                rows = order[s0:s0 + batch]
                sk, rg, y = _taps_of(split, rows)
                if phase == "train":
                    lr = sched.step()
                    logits, loss, _ = head.train_step(sk, rg, y, lr)      # backbone logits are constants: same gradients
                else:
                    logits, _ = head.forward(sk, rg, train=False)
                loss, preds = multitask_loss_preds(logits, y, _aux_of(split, rows, multitask))
                run_loss += float(loss) * len(rows)
                run_correct += int((preds == y).sum())
            row[phase + "_loss"] = run_loss / n
            row[phase + "_acc"] = run_correct / n
            if log:
                log('{} Loss: {:.4f} Acc: {:.4f}'.format(phase, row[phase + "_loss"], row[phase + "_acc"]))
            if phase == "dev" and row["dev_acc"] > best_acc:
                best_acc = row["dev_acc"]
                best_state = copy.deepcopy(head.state)
        stats.append(row)
    head.state = best_state
    return np.float64(best_acc), stats


def test_track_acc(head, split, batch, order, multitask=False):
    order = np.asarray(order)
    correct = 0
    for s0 in range(0, len(order), batch):
        rows = order[s0:s0 + batch]
        sk, rg, y = _taps_of(split, rows)
        logits, _ = head.forward(sk, rg, train=False)
        _, preds = multitask_loss_preds(logits, y, _aux_of(split, rows, multitask))
        correct += int((preds == y).sum())
    return np.float64(correct / len(order))


# --------------------------------------------------------------------------------------
# weight sharing between the candidates of one call (args.weightsharing)
# --------------------------------------------------------------------------------------
_ACT_TAG = {ACT_RELU: ".A_relu", ACT_SIGMOID: ".A_sigmoid", ACT_LRELU: ".A_lrelu"}
_LAYER_KEYS = ("0.weight", "0.bias", "2.weight", "2.bias", "2.running_mean", "2.running_var", "2.num_batches_tracked")


def shared_key(head, l):
    """'<step>.L_<in_features>_<out_features>.A_<act>' (ntu_searchable.py:131-140 / :161-170)."""
    return f"{l}.L_{head.K[l]}_{head.H}" + _ACT_TAG.get(int(head.conf[l][2]), "")


def get_central_states(head, shared):
    """ntu_searchable.py:123-149: every fusion step of the trained candidate is stored under its key (created or overwritten)."""
    for l in range(head.L):
        shared[shared_key(head, l)] = {k: np.array(head.state[f"fusion_layers.{l}.{k}"], copy=True)
                                       for k in _LAYER_KEYS if f"fusion_layers.{l}.{k}" in head.state}
    return shared


def set_central_states(head, shared):
    """ntu_searchable.py:152-174: a fusion step whose key is in the dict starts from the stored layer (weights AND BatchNorm
    buffers: ``layer.load_state_dict``); the classifier, the alphas and the optimiser state are never shared."""
    for l in range(head.L):
        key = shared_key(head, l)
        if key in shared:
            for k, v in shared[key].items():
                head.state[f"fusion_layers.{l}.{k}"] = np.array(v, copy=True)


def train_sampled_heads(heads, scheds, train_split, dev_split, batch, orders, num_epochs, weightsharing=False, shared=None):
    """The candidate loop of train_sampled_models (ntu_searchable.py:36-97) over already-constructed heads: candidate ``ci``
    sees passes ``orders(phase, ci, epoch)``; with ``weightsharing`` the candidates are chained through ``shared`` in order."""
    shared = {} if shared is None else shared
    accs, all_stats = [], []
    for ci, (head, sched) in enumerate(zip(heads, scheds)):
        if weightsharing:
            set_central_states(head, shared)
        best, stats = train_track_acc(head, sched, train_split, dev_split, batch, lambda ph, e, ci=ci: orders(ph, ci, e), num_epochs)
        if weightsharing:
            get_central_states(head, shared)
        accs.append(best)
        all_stats.append(stats)
    return accs, all_stats


# --------------------------------------------------------------------------------------
# algorithmic bytes / flops -- used by bench.py for the roofline
# --------------------------------------------------------------------------------------
def algorithmic_counts(conf, H, C, B, bn=True, vid_len_ske=32):
    conf = np.asarray(conf).reshape(-1, 3)
    ds = d_ske(vid_len_ske)
    L = len(conf)
    F_sel = sum(ds[int(c[0])] + D_RGB[int(c[1])] for c in conf)
    K = layer_in_features(conf, H, vid_len_ske)
    P = sum(k * H + H for k in K) + (2 * H * L if bn else 0) + C * H + C
    train_bytes = 4 * (B * F_sel + 6 * P) + 8 * B
    eval_bytes = 4 * (B * F_sel + P) + 8 * B
    fwd_flops = 2 * B * (sum(K) * H + H * C)
    bwd_flops = fwd_flops + 2 * B * H * H * (L - 1) + 2 * B * H * C
    return dict(F_sel=F_sel, K=K, P=P, train_bytes=train_bytes, eval_bytes=eval_bytes,
                fwd_flops=fwd_flops, bwd_flops=bwd_flops)
