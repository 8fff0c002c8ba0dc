"""Host-side handle over the C ABI: a group of candidate fusion heads trained together on one GPU.

PyTorch is plumbing here (device memory, streams); all arithmetic runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .cache import D_RGB, FeatureCache, ske_widths


def flags_from_args(args) -> int:
    """Layer recipe flags (/root/reference/models/search/ntu_searchable.py:274-284)."""
    f = 0
    if getattr(args, "batchnorm", False):
        f |= _lib.FLAG_BN
    if getattr(args, "drpt", 0.0) > 1e-10:
        f |= _lib.FLAG_DROPOUT
    if getattr(args, "alphas", False):
        f |= _lib.FLAG_ALPHAS
    if getattr(args, "multitask", False):
        f |= _lib.FLAG_MULTITASK
    return f


def _four(widths):
    """The ABI carries MFAS_NUM_TAPS = 8 tap slots per modality; a shorter tap set (NTU 4 + 4, MM-IMDB 2 + 4, AV-MNIST 5 + 3) is
    repeated to fill the unused slots -- ``plan_layout`` rejects a conf row that points at one of them."""
    w = [int(x) for x in widths]
    if not 1 <= len(w) <= _lib.NUM_TAPS:
        raise ValueError(f"a modality has 1..{_lib.NUM_TAPS} taps, got {len(w)}")
    return [w[t % len(w)] for t in range(_lib.NUM_TAPS)]


def plan_layout(conf, H, C_out, flags, vid_len_ske=32, widths=None) -> _lib.Layout:
    """``widths`` = (first-modality tap widths, second-modality tap widths); default: the NTU taps."""
    conf = np.ascontiguousarray(np.asarray(conf, dtype=np.int32).reshape(-1, 3))
    lay = _lib.Layout()
    d0, d1 = (ske_widths(vid_len_ske), D_RGB) if widths is None else widths
    for col, d in ((0, d0), (1, d1)):           # slots a shorter tap set leaves unused (the library checks the range [0, 4))
        if conf.shape[0] and len(d) <= conf[:, col].max() < _lib.NUM_TAPS:
            raise ValueError(f"conf tap index out of range for a ({len(d0)}, {len(d1)})-tap set: {conf.tolist()}")
    ds = (C.c_int32 * _lib.NUM_TAPS)(*_four(d0))
    dr = (C.c_int32 * _lib.NUM_TAPS)(*_four(d1))
    _lib.check(_lib.lib().mfas_plan_layout(conf.shape[0], conf.ctypes.data_as(C.POINTER(C.c_int32)), int(H),
                                           int(C_out), int(flags), ds, dr, C.byref(lay)))
    return lay


def algorithmic_counts(lay: _lib.Layout, batch: int):
    out = (C.c_double * 4)()
    _lib.check(_lib.lib().mfas_algorithmic_counts(C.byref(lay), int(batch), out))
    return dict(train_bytes=out[0], eval_bytes=out[1], fwd_flops=out[2], bwd_flops=out[3])


def tensor_slots(lay: _lib.Layout):
    """state-dict name -> (arena, offset, shape), names as in the reference state_dict."""
    s = {}
    H, Cn = lay.H, lay.C
    for l in range(lay.L):
        s[f"fusion_layers.{l}.0.weight"] = ("p", lay.off_W[l], (H, lay.K[l]))
        s[f"fusion_layers.{l}.0.bias"] = ("p", lay.off_b[l], (H,))
        if lay.flags & _lib.FLAG_BN:
            s[f"fusion_layers.{l}.2.weight"] = ("p", lay.off_gamma[l], (H,))
            s[f"fusion_layers.{l}.2.bias"] = ("p", lay.off_beta[l], (H,))
            s[f"fusion_layers.{l}.2.running_mean"] = ("b", lay.off_rm[l], (H,))
            s[f"fusion_layers.{l}.2.running_var"] = ("b", lay.off_rv[l], (H,))
            s[f"fusion_layers.{l}.2.num_batches_tracked"] = ("n", l, ())
        s[f"alphas.{l}.alpha_x"] = ("p", lay.off_alpha[l], (1,))
    s["central_classifier.weight"] = ("p", lay.off_Wc, (Cn, H))
    s["central_classifier.bias"] = ("p", lay.off_bc, (Cn,))
    return s


def cache_desc(cache: FeatureCache) -> _lib.CacheDesc:
    if cache.device.type != "cuda":
        raise RuntimeError("feature cache must be resident on a CUDA device (FeatureCache.to('cuda'))")
    d = _lib.CacheDesc()
    d.n_rows = len(cache)
    for widths, cat, ptr, ld, dw in ((cache.d_ske, cache.ske_cat, d.ske, d.ske_ld, d.d_ske),
                                     (cache.d_rgb, cache.rgb_cat, d.rgb, d.rgb_ld, d.d_rgb)):
        offs = np.concatenate([[0], np.cumsum(widths)])
        for t in range(_lib.NUM_TAPS):             # a shorter tap set repeats (see _four)
            u = t % len(widths)
            ptr[t] = cat.data_ptr() + 4 * int(offs[u])
            ld[t] = cat.stride(0)
            dw[t] = widths[u]
    if getattr(cache, "multilabel", False):
        d.labels = None
        d.targets = cache.labels.data_ptr()
        d.pos_weight = cache.pos_weight.data_ptr()
    else:
        d.labels = cache.labels.data_ptr()
    d.logit_rgb = cache.logit_rgb.data_ptr() if cache.logit_rgb is not None else None
    d.logit_ske = cache.logit_ske.data_ptr() if cache.logit_ske is not None else None
    return d


def adam_schedule(lrs, t0, beta1=0.9, beta2=0.999):
    """Per-step scalars of torch's Adam, formed in fp64 as torch does: lr/(1-b1^t), sqrt(1-b2^t)."""
    lrs = np.asarray(lrs, dtype=np.float64)
    t = t0 + 1 + np.arange(len(lrs), dtype=np.float64)
    step_size = (lrs / (1.0 - beta1 ** t)).astype(np.float32)
    bc2_sqrt = np.sqrt(1.0 - beta2 ** t).astype(np.float32)
    return np.ascontiguousarray(step_size), np.ascontiguousarray(bc2_sqrt)


class GroupLayout:
    """Where every tensor of every candidate of a group lives in the flat arenas (host-only, no GPU)."""

    def __init__(self, confs, H, C_out, flags, vid_len_ske=32, widths=None):
        self.n = len(confs)
        self.H, self.C, self.flags = int(H), int(C_out), int(flags)
        self.layouts = [plan_layout(c, H, C_out, flags, vid_len_ske, widths) for c in confs]
        self.slots = [tensor_slots(l) for l in self.layouts]
        np_ = [int(l.n_params) for l in self.layouts]
        nb_ = [int(l.n_bufs) for l in self.layouts]
        self.p_off = np.concatenate([[0], np.cumsum(np_)]).astype(np.int64)
        self.b_off = np.concatenate([[0], np.cumsum(nb_)]).astype(np.int64)


class CandidateGroup(GroupLayout):
    """n candidates with their parameter / Adam / BN arenas resident on one CUDA device."""

    @staticmethod
    def plan(confs, H, C_out, flags, vid_len_ske=32, widths=None):
        """The layouts of ``confs`` without creating a group (host only)."""
        return [plan_layout(c, H, C_out, flags, vid_len_ske, widths) for c in confs]

    def __init__(self, confs, H, C_out, flags, device, batch_max, drop_p=0.0, drop_seed=0, cand_ids=None,
                 vid_len_ske=32, keep_grads=False, widths=None):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("mfas_b200 runs on CUDA devices only (no CPU fallback)")
        super().__init__(confs, H, C_out, flags, vid_len_ske, widths)
        self.device = device
        self.batch_max = int(batch_max)
        z = lambda n, dt=torch.float32: torch.zeros(int(n), dtype=dt, device=device)
        self.params, self.adam_m, self.adam_v = z(self.p_off[-1]), z(self.p_off[-1]), z(self.p_off[-1])
        self.grads = z(self.p_off[-1]) if keep_grads else None
        self.bufs = z(self.b_off[-1])
        self.nbt = z(self.n * _lib.MAX_LAYERS, torch.int64)
        self.adam_t = 0
        lay_arr = (_lib.Layout * self.n)(*self.layouts)
        ids = None
        if cand_ids is not None:
            ids = (C.c_int32 * self.n)(*[int(i) for i in cand_ids])
        h = C.c_void_p()
        _lib.check(_lib.lib().mfas_group_create(device.index or 0, self.n, lay_arr, self.batch_max, float(drop_p),
                                                int(drop_seed) & 0xFFFFFFFF, ids, C.byref(h)))
        self._h = h
        for c in range(self.n):
            a = _lib.Arenas()
            a.params = self.params.data_ptr() + 4 * int(self.p_off[c])
            a.adam_m = self.adam_m.data_ptr() + 4 * int(self.p_off[c])
            a.adam_v = self.adam_v.data_ptr() + 4 * int(self.p_off[c])
            a.grad = (self.grads.data_ptr() + 4 * int(self.p_off[c])) if keep_grads else None
            a.bufs = self.bufs.data_ptr() + 4 * int(self.b_off[c])
            a.nbt = self.nbt.data_ptr() + 8 * c * _lib.MAX_LAYERS
            _lib.check(_lib.lib().mfas_group_bind(self._h, c, C.byref(a)))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().mfas_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tensors ------------------------------------------------------------------------------
    def view(self, c, name, arena=None):
        kind, off, shape = self.slots[c][name]
        if kind == "n":
            return self.nbt[c * _lib.MAX_LAYERS + off]
        if kind == "b":
            base, o = self.bufs, int(self.b_off[c]) + int(off)
        else:
            base = {None: self.params, "p": self.params, "m": self.adam_m, "v": self.adam_v, "g": self.grads}[arena]
            o = int(self.p_off[c]) + int(off)
        n = int(np.prod(shape)) if shape else 1
        return base[o:o + n].view(shape)

    def names(self, c):
        return list(self.slots[c].keys())

    def load_state(self, c, state):
        """Copy a reference-style state_dict (tensors or numpy arrays) into candidate c's arenas."""
        with torch.no_grad():
            for name in self.names(c):
                if name not in state:
                    continue
                src = state[name]
                src = torch.from_numpy(np.asarray(src)) if not torch.is_tensor(src) else src
                self.view(c, name).copy_(src.to(self.device).reshape(self.view(c, name).shape))

    def state(self, c, arena=None):
        out = {}
        for name in self.names(c):
            kind = self.slots[c][name][0]
            if arena in ("m", "v", "g") and kind != "p":
                continue
            out[name] = self.view(c, name, arena).detach().cpu().numpy().copy()
        return out

    def init_params(self, seed):
        """Initial weights of every candidate in one launch, keyed by (seed, candidate id) -- see mfas_group_init_params."""
        _lib.check(_lib.lib().mfas_group_init_params(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, self._stream()))
        self.adam_t = 0

    def set_adam(self, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-4):
        hp = _lib.AdamHParams(beta1, beta2, eps, weight_decay)
        self.betas = (beta1, beta2)
        _lib.check(_lib.lib().mfas_group_set_adam(self._h, C.byref(hp)))

    @property
    def engine(self):
        e = C.c_int32()
        _lib.check(_lib.lib().mfas_group_engine(self._h, C.byref(e)))
        return {0: "ffma", 1: "tc"}[e.value]

    def set_profiling(self, on=True):
        _lib.check(_lib.lib().mfas_group_set_profiling(self._h, 1 if on else 0))

    def last_step_ms(self):
        """Device time (ms) of the forward streaming / fused chain / backward streaming kernel of the last train step."""
        ms = (C.c_float * 3)()
        _lib.check(_lib.lib().mfas_group_last_step_ms(self._h, ms))
        return [float(x) for x in ms]

    def check(self):
        """Synchronise and raise if a kernel reported a failure."""
        _lib.check(_lib.lib().mfas_group_status(self._h))

    @property
    def launches(self):
        n = C.c_int64()
        _lib.check(_lib.lib().mfas_group_num_launches(self._h, C.byref(n)))
        return n.value

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _rows(self, rows):
        """rows: int tensor [n_rows] (shared) or [n, n_rows] (per candidate) -> (tensor, stride, n_rows)."""
        rows = rows.to(device=self.device, dtype=torch.int32).contiguous()
        if rows.dim() == 1:
            return rows, 0, rows.shape[0]
        assert rows.shape[0] == self.n
        return rows, rows.shape[1], rows.shape[1]

    # ---- single-batch ops (op-level parity, model.forward) ----------------------------------------
    def forward(self, cache, rows, train=False, step=0):
        rows, stride, n = self._rows(rows)
        logits = torch.empty(self.n, self.batch_max, self.C, device=self.device)
        loss = torch.empty(self.n, device=self.device)
        correct = torch.empty(self.n, dtype=torch.int32, device=self.device)
        d = cache_desc(cache)
        _lib.check(_lib.lib().mfas_forward(self._h, C.byref(d), rows.data_ptr(), stride, n, int(train), int(step),
                                           logits.data_ptr(), loss.data_ptr(), correct.data_ptr(), self._stream()))
        return logits[:, :n], loss, correct

    def train_step(self, cache, rows, lr, beta1=0.9, beta2=0.999):
        rows, stride, n = self._rows(rows)
        ss, b2 = adam_schedule([lr], self.adam_t, beta1, beta2)
        logits = torch.empty(self.n, self.batch_max, self.C, device=self.device)
        loss = torch.empty(self.n, device=self.device)
        correct = torch.empty(self.n, dtype=torch.int32, device=self.device)
        d = cache_desc(cache)
        _lib.check(_lib.lib().mfas_train_step(self._h, C.byref(d), rows.data_ptr(), stride, n, float(ss[0]), float(b2[0]),
                                              self.adam_t, logits.data_ptr(), loss.data_ptr(), correct.data_ptr(),
                                              self._stream()))
        self.adam_t += 1
        return logits[:, :n], loss, correct

    # ---- the production path ---------------------------------------------------------------------
    def train_run(self, train_cache, dev_cache, perm_train, perm_dev, lrs, epochs, batch, beta1=0.9, beta2=0.999, best_init=None):
        """epochs x (train pass + dev pass) for all candidates; returns device tensors
        (stats [n, epochs, 4] f64, best_acc [n] f64, best_epoch [n] i32) without synchronising.
        ``best_init``: per-candidate value the best-dev tracking starts from (default 0; init_f1 of the MM-IMDB loop)."""
        n_tr, n_dv = len(train_cache), len(dev_cache)
        steps = math.ceil(n_tr / batch)
        assert len(lrs) == epochs * steps, (len(lrs), epochs, steps)
        perm_train = perm_train.to(device=self.device, dtype=torch.int32).contiguous()
        assert perm_train.shape == (self.n, epochs, n_tr), perm_train.shape
        if perm_dev is not None:
            perm_dev = perm_dev.to(device=self.device, dtype=torch.int32).contiguous()
            assert perm_dev.shape == (self.n, epochs, n_dv), perm_dev.shape
        ss, b2 = adam_schedule(lrs, self.adam_t, beta1, beta2)
        stats = torch.empty(self.n, max(epochs, 1), 4, dtype=torch.float64, device=self.device)
        best_acc = torch.empty(self.n, dtype=torch.float64, device=self.device)
        best_epoch = torch.empty(self.n, dtype=torch.int32, device=self.device)
        a = _lib.RunArgs()
        a.n_epochs, a.batch = int(epochs), int(batch)
        a.perm_train = perm_train.data_ptr()
        a.perm_dev = perm_dev.data_ptr() if perm_dev is not None else None
        a.step_size, a.bc2_sqrt = ss.ctypes.data, b2.ctypes.data
        a.adam_t0 = self.adam_t
        a.stats, a.best_acc, a.best_epoch = stats.data_ptr(), best_acc.data_ptr(), best_epoch.data_ptr()
        a.best_acc_init = None
        if best_init is not None:
            bi = np.ascontiguousarray(np.broadcast_to(np.asarray(best_init, dtype=np.float64), (self.n,)))
            a.best_acc_init = bi.ctypes.data
        dtr, ddv = cache_desc(train_cache), cache_desc(dev_cache)
        _lib.check(_lib.lib().mfas_train_run(self._h, C.byref(dtr), C.byref(ddv), C.byref(a), self._stream()))
        self.adam_t += epochs * steps
        self._keepalive = (perm_train, perm_dev)      # until the stream has consumed them
        return stats, best_acc, best_epoch

    def eval_pass(self, cache, batch, perm=None):
        out = torch.empty(self.n, 2, dtype=torch.float64, device=self.device)
        if perm is not None:
            perm = perm.to(device=self.device, dtype=torch.int32).contiguous()
            assert perm.shape == (self.n, len(cache))
        d = cache_desc(cache)
        _lib.check(_lib.lib().mfas_eval_pass(self._h, C.byref(d), perm.data_ptr() if perm is not None else None,
                                             int(batch), out.data_ptr(), self._stream()))
        self._keepalive = perm
        return out
