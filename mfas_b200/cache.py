"""Feature cache: the data format on the input side of the hot path.

The reference re-runs both frozen backbones on every batch
(/root/reference/models/search/ntu_searchable.py:206-225).  north_star replaces that with
pre-extracted, globally pooled backbone taps held resident in HBM:

    ske_cat : [N, sum(D_ske)] fp32   the 4 skeleton taps  skenet(x)[0][-4:]  side by side
    rgb_cat : [N, sum(D_rgb)] fp32   the 4 visual taps    rgbnet(x)[-5:-1]   side by side
    labels  : [N] int64

Tap ``i`` of a modality is the column slice ``[off_i, off_i + D_i)`` with leading dimension
``sum(D)``; the CUDA kernels take (base pointer, ld) per tap, so no copy is ever made.

``FeatureCacheLoader`` is the DataLoader stand-in the reference driver iterates: it has
``.dataset`` (for ``len``) and yields ``{'rgb','ske','label'}`` dict batches
(/root/reference/datasets/ntu.py:84-87), so the *unmodified* reference loop can consume the
same cache through parameter-free stub backbones -- that is how parity fixtures are produced.
"""
from __future__ import annotations

import torch

D_RGB = (512, 1024, 2048, 2048)      # /root/reference/models/search/ntu_searchable.py:292


def ske_widths(vid_len_ske: int = 32):
    """/root/reference/models/search/ntu_searchable.py:291."""
    return (128, 256, 32 * int(vid_len_ske), 512)


def _offsets(widths):
    out, o = [], 0
    for w in widths:
        out.append(o)
        o += w
    return out


class FeatureCache:
    """One split (train / dev / test) of cached backbone taps."""

    def __init__(self, ske_cat, rgb_cat, labels, vid_len_ske=32, logit_rgb=None, logit_ske=None, widths=None,
                 pos_weight=None):
        """``widths`` = (first-modality tap widths, second-modality tap widths) for a tap set other than NTU's (the
        MM-IMDB text / image taps, mfas_b200.mmimdb_searchable); the first modality rides in ``ske_cat``, the second in
        ``rgb_cat``.  ``labels`` is int64 [N] class ids, or fp32 [N, C] multi-hot targets together with ``pos_weight``
        [C] for the multi-label head."""
        self.d_ske, self.d_rgb = (ske_widths(vid_len_ske), D_RGB) if widths is None else (tuple(widths[0]), tuple(widths[1]))
        self.widths = widths if widths is None else (self.d_ske, self.d_rgb)
        assert ske_cat.dtype == torch.float32 and rgb_cat.dtype == torch.float32
        assert ske_cat.shape[1] == sum(self.d_ske), ske_cat.shape
        assert rgb_cat.shape[1] == sum(self.d_rgb), rgb_cat.shape
        assert labels.shape[0] == ske_cat.shape[0] == rgb_cat.shape[0]
        self.multilabel = labels.dim() == 2
        if self.multilabel:
            assert labels.dtype == torch.float32 and pos_weight is not None and pos_weight.dtype == torch.float32
            assert pos_weight.shape == (labels.shape[1],), (pos_weight.shape, labels.shape)
        else:
            assert labels.dtype == torch.int64
        self.ske_cat = ske_cat.contiguous()
        self.rgb_cat = rgb_cat.contiguous()
        self.labels = labels.contiguous()
        self.pos_weight = None if pos_weight is None else pos_weight.contiguous()
        self.logit_rgb = logit_rgb       # [N, C] backbone logits, only for multitask
        self.logit_ske = logit_ske
        self.vid_len_ske = vid_len_ske
        self._device_copies = {}

    def __len__(self):
        return self.labels.shape[0]

    @property
    def device(self):
        return self.ske_cat.device

    def nbytes(self):
        n = self.ske_cat.numel() * 4 + self.rgb_cat.numel() * 4 + self.labels.numel() * self.labels.element_size()
        for t in (self.logit_rgb, self.logit_ske):
            if t is not None:
                n += t.numel() * 4
        return n

    def ske_taps(self):
        return [self.ske_cat[:, o:o + w] for o, w in zip(_offsets(self.d_ske), self.d_ske)]

    def rgb_taps(self):
        return [self.rgb_cat[:, o:o + w] for o, w in zip(_offsets(self.d_rgb), self.d_rgb)]

    def __getitem__(self, i):
        """Sample dict with the reference's keys (/root/reference/datasets/ntu.py:84-87)."""
        if self.logit_rgb is not None:
            return {'rgb': torch.cat((self.rgb_cat[i], self.logit_rgb[i])), 'ske': torch.cat((self.ske_cat[i], self.logit_ske[i])),
                    'label': self.labels[i]}
        return {'rgb': self.rgb_cat[i], 'ske': self.ske_cat[i], 'label': self.labels[i]}

    def pin(self):
        if self.device.type == 'cpu' and torch.cuda.is_available() and not self.ske_cat.is_pinned():
            self.ske_cat, self.rgb_cat, self.labels = (self.ske_cat.pin_memory(), self.rgb_cat.pin_memory(),
                                                       self.labels.pin_memory())
            if self.pos_weight is not None:
                self.pos_weight = self.pos_weight.pin_memory()
            if self.logit_rgb is not None:
                self.logit_rgb, self.logit_ske = self.logit_rgb.pin_memory(), self.logit_ske.pin_memory()
        return self

    def to(self, device, non_blocking=True, memoize=True):
        """Upload (H2D) the split once; later calls return the resident copy."""
        device = torch.device(device)
        if device == self.device:
            return self
        key = str(device)
        if memoize and key in self._device_copies:
            return self._device_copies[key]
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        out = FeatureCache(mv(self.ske_cat), mv(self.rgb_cat), mv(self.labels), self.vid_len_ske,
                           mv(self.logit_rgb), mv(self.logit_ske), self.widths, mv(self.pos_weight))
        if memoize:
            self._device_copies[key] = out
        return out

    def drop_device_copies(self):
        self._device_copies.clear()


_M32 = 0xFFFFFFFF


def _mix32(x):
    """32-bit integer mixer on int64 tensors (identical results on CPU and CUDA)."""
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    return x ^ (x >> 16)


def hashed_orders_of(seed: int, pass_ids, n: int, device="cpu") -> torch.Tensor:
    """``hashed_orders`` for an arbitrary list of pass numbers (a rank's share of a sharded call): int64 [len(ids), n]."""
    dev = torch.device(device)
    k = torch.as_tensor(list(pass_ids), dtype=torch.int64, device=dev)[:, None]
    i = torch.arange(n, dtype=torch.int64, device=dev)[None, :]
    key = _mix32(_mix32((int(seed) & _M32) * 0x9E3779B1 + k * 0x85EBCA6B) + i)
    return torch.sort(key, dim=1, stable=True).indices


def hashed_orders(seed: int, first_pass: int, count: int, n: int, device="cpu") -> torch.Tensor:
    """Row orders of ``count`` consecutive passes as an int64 [count, n] tensor: pass k visits rows in
    the (stable) argsort of hash(seed, k, row).  Pure integer arithmetic, so the CPU (reference loop,
    oracle) and the GPU (one batched sort for all candidates x epochs of a call) produce the same
    permutations bit for bit."""
    dev = torch.device(device)
    k = torch.arange(first_pass, first_pass + count, dtype=torch.int64, device=dev)[:, None]
    i = torch.arange(n, dtype=torch.int64, device=dev)[None, :]
    key = _mix32(_mix32((int(seed) & _M32) * 0x9E3779B1 + k * 0x85EBCA6B) + i)
    return torch.sort(key, dim=1, stable=True).indices


class FeatureCacheLoader:
    """Deterministic DataLoader stand-in over a FeatureCache.

    The k-th pass (``__iter__`` call) over the loader visits rows in ``order_for_pass(k)``: a hashed
    permutation keyed by (seed, k) when ``shuffle`` (the reference shuffles both 'train' and 'dev',
    /root/reference/models/searchable.py:247-250), identity otherwise.  The order depends on (seed, k)
    only, so candidate ``idx`` / epoch ``e`` of a ``train_sampled_models`` call sees pass number
    ``base + idx*epochs + e`` no matter which GPU trains it -- 1-GPU and 8-GPU runs return identical
    accuracies.  ``drop_last`` is False as in the reference.
    """

    def __init__(self, cache: FeatureCache, batch_size: int, shuffle: bool = True, seed: int = 0):
        self.dataset = cache
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.seed = int(seed)
        self.passes = 0          # number of passes handed out so far

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def orders(self, first_pass: int, count: int, device="cpu") -> torch.Tensor:
        """int64 [count, N] row orders of passes first_pass .. first_pass+count-1, built on ``device``."""
        n = len(self.dataset)
        if not self.shuffle:
            return torch.arange(n, dtype=torch.int64, device=device).repeat(count, 1)
        return hashed_orders(self.seed, first_pass, count, n, device)

    def orders_of(self, pass_ids, device="cpu") -> torch.Tensor:
        """int64 [len(pass_ids), N] row orders of the given passes (any subset, any order), built on ``device``."""
        n = len(self.dataset)
        if not self.shuffle:
            return torch.arange(n, dtype=torch.int64, device=device).repeat(len(pass_ids), 1)
        return hashed_orders_of(self.seed, pass_ids, n, device)

    def order_for_pass(self, k: int) -> torch.Tensor:
        return self.orders(int(k), 1)[0]

    def take_passes(self, count: int) -> int:
        """Reserve ``count`` consecutive passes; returns the first pass number."""
        k = self.passes
        self.passes += int(count)
        return k

    def __iter__(self):
        order = self.order_for_pass(self.take_passes(1))
        c = self.dataset
        dev = c.device
        for s in range(0, len(order), self.batch_size):
            rows = order[s:s + self.batch_size].to(dev)
            rgb, ske = c.rgb_cat.index_select(0, rows), c.ske_cat.index_select(0, rows)
            if c.logit_rgb is not None:      # multitask: the cached backbone logits ride behind the taps of their modality
                rgb = torch.cat((rgb, c.logit_rgb.index_select(0, rows)), 1)
                ske = torch.cat((ske, c.logit_ske.index_select(0, rows)), 1)
            yield {'rgb': rgb, 'ske': ske, 'label': c.labels.index_select(0, rows)}


def synthetic_ntu_cache(n_rows: int, seed: int, num_outputs: int = 60, vid_len_ske: int = 32,
                        signal: float = 3.0, with_backbone_logits: bool = False) -> FeatureCache:
    """NTU-shaped synthetic split.

    Features are |N(0,1)| (taps are post-ReLU, globally pooled => non-negative); ``signal`` *
    onehot(label) is added to the first ``num_outputs`` columns of rgb tap 0 and ske tap 3 so
    that accuracy is learnable and val-acc parity is meaningful.
    """
    g = torch.Generator()
    g.manual_seed(int(seed))
    ds, dr = ske_widths(vid_len_ske), D_RGB
    ske = torch.randn(n_rows, sum(ds), generator=g).abs_()
    rgb = torch.randn(n_rows, sum(dr), generator=g).abs_()
    labels = torch.randint(0, num_outputs, (n_rows,), generator=g, dtype=torch.int64)
    onehot = torch.zeros(n_rows, num_outputs).scatter_(1, labels[:, None], float(signal))
    rgb[:, :num_outputs] += onehot                              # rgb tap 0
    o3 = sum(ds[:3])
    ske[:, o3:o3 + num_outputs] += onehot                       # ske tap 3
    lr = ls = None
    if with_backbone_logits:
        lr = torch.randn(n_rows, num_outputs, generator=g) + onehot * 0.5
        ls = torch.randn(n_rows, num_outputs, generator=g) + onehot * 0.5
    return FeatureCache(ske, rgb, labels, vid_len_ske, lr, ls)
