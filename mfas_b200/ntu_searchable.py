"""Drop-in for /root/reference/models/search/ntu_searchable.py on cached backbone taps.

Same module attributes the search driver uses (models/searchable.py:23,256-260):
``train_sampled_models``, ``get_possible_layer_configurations``, ``get_central_states``,
``set_central_states``, ``Searchable_Skeleton_Image_Net`` -- same signatures, same state_dict key
names, same return types -- but candidates are trained by the CUDA library (mfas_b200/csrc), many
at a time, with nothing returning to the host between the first batch and the last epoch.

What is deliberately different from the reference:
  * the frozen backbones are replaced by a FeatureCache (D5): ``rgbnet`` / ``skenet`` are
    parameter-free tap splitters kept only so ``.load_state_dict`` calls keep working;
  * accuracies come back as 0-dim float64 *CPU* tensors (D10) so ``np.array(accs)`` in
    models/search/tools.py:47-49 works;
  * there is no CPU execution path: a non-CUDA ``device`` raises.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .cache import D_RGB, FeatureCache, ske_widths
from .engine import CandidateGroup, GroupLayout, flags_from_args
from .scheduler import LRCosineAnnealingScheduler


# ----------------------------------------------------------------------------------------------
# small modules of the reference's model tree
# ----------------------------------------------------------------------------------------------
class GlobalPooling2D(nn.Module):
    """Mean over every dim but (batch, channel) (/root/reference/models/auxiliary/aux_models.py:54-64).
    Cached taps are already pooled, so on the hot path this is the identity."""

    def forward(self, x):
        return x.reshape(x.size(0), x.size(1), -1).mean(2)


class AlphaScalarMultiplication(nn.Module):
    """Scalar modality gate x*sigmoid(a), y*(1-sigmoid(a)) (aux_models.py:94-111)."""

    def __init__(self, size_alpha_x, size_alpha_y):
        super().__init__()
        self.size_alpha_x = size_alpha_x
        self.size_alpha_y = size_alpha_y
        self.alpha_x = nn.Parameter(torch.zeros(1, dtype=torch.float32))

    def forward(self, x, y):
        s = torch.sigmoid(self.alpha_x)
        return x * s, y * (1.0 - s)


class CachedTaps(nn.Module):
    """Stands where a frozen backbone stood (models/central/ntu.py:17-183): parameter-free, splits a
    row of concatenated cached taps back into the backbone's output structure.  Accepts (and
    ignores) any backbone checkpoint so ``rmode.skenet.load_state_dict(torch.load(...))``
    (ntu_searchable.py:46-49) keeps working."""

    def __init__(self, widths, kind):
        super().__init__()
        self.widths, self.kind = tuple(widths), kind

    def forward(self, x):
        n = sum(self.widths)
        logits = x[:, n:] if x.shape[1] > n else None         # cached backbone logits ride behind the taps (multitask)
        taps = torch.split(x[:, :n], self.widths, 1)
        if self.kind == "rgb":                       # Visual.forward 6-tuple, central/ntu.py:50
            return (None, *taps, logits)
        if self.kind == "lenet":                     # GP_LeNet / GP_LeNet_Deeper.forward, central/avmnist.py:57,112: (logits, gp1, gp2, ...)
            return (logits, *taps)
        return [None] * 4 + list(taps), logits       # Skeleton.forward, central/ntu.py:183

    def load_state_dict(self, state_dict, strict=True, assign=False):
        return nn.modules.module._IncompatibleKeys([], [])


def _activation(kind):
    kind = int(kind)
    if kind == 0:
        return nn.ReLU()
    if kind == 1:
        return nn.Sigmoid()
    if kind == 2:
        return nn.LeakyReLU()
    raise ValueError(f"activation id {kind} not in {{0,1,2}}")


# ----------------------------------------------------------------------------------------------
# the searchable fusion network
# ----------------------------------------------------------------------------------------------
class Searchable_Skeleton_Image_Net(nn.Module):
    """Searchable fusion head (/root/reference/models/search/ntu_searchable.py:178-301).

    conf: int array [L, 3], one row per fusion step: [ske tap, rgb tap, nonlinearity].
    forward(tensor_tuple): tensor_tuple = (rgb, ske) with rgb [B, 5632] / ske [B, 1920] the
    concatenated cached taps (what FeatureCacheLoader yields); returns [B, num_outputs] logits
    computed by the CUDA library.  The sub-modules are ordinary nn.Linear / BatchNorm1d objects so
    ``state_dict()`` has the reference's keys; once on a CUDA device their storage is the flat
    arenas the kernels update in place.
    """

    # hooks of the sibling networks over other tap sets (mfas_b200.mmimdb_searchable): None / 0 = the NTU taps, softmax-CE head
    _widths_kw = None
    _extra_flags = 0

    @staticmethod
    def _tap_widths(args):
        """(first-modality tap widths, second-modality tap widths), ntu_searchable.py:291-292"""
        return ske_widths(args.vid_len[1]), D_RGB

    def __init__(self, args, conf):
        super().__init__()
        self.conf = conf
        self.args = args
        ds, dr = self._tap_widths(args)
        self.rgbnet = CachedTaps(dr, "rgb")
        self.skenet = CachedTaps(ds, "ske")
        cf = np.asarray(conf).reshape(-1, 3)
        # construction order == reference order, so a given torch seed gives the same init
        self.alphas = nn.ModuleList([AlphaScalarMultiplication(ds[int(c[0])], dr[int(c[1])]) for c in cf])
        self.gp_v = nn.ModuleList([GlobalPooling2D() for _ in cf])
        self.gp_s = nn.ModuleList([GlobalPooling2D() for _ in cf])
        self.fusion_layers = self._create_fc_layers(cf)
        self.central_classifier = nn.Linear(args.inner_representation_size, args.num_outputs)
        for m in self.modules():
            if isinstance(m, AlphaScalarMultiplication):
                nn.init.normal_(m.alpha_x, 0.0, 0.1)
        self._group = None          # CandidateGroup whose arenas back this module
        self._slot = 0
        self._fwd_steps = 0

    def _create_fc_layers(self, cf):
        H, drpt, bn = self.args.inner_representation_size, self.args.drpt, self.args.batchnorm
        layers = []
        for i, c in enumerate(cf):
            in_size = self.alphas[i].size_alpha_x + self.alphas[i].size_alpha_y + (H if i > 0 else 0)
            nl = _activation(c[2])
            if drpt > 1e-10 and bn:
                op = nn.Sequential(nn.Linear(in_size, H), nl, nn.BatchNorm1d(H), nn.Dropout(drpt))
            elif drpt > 1e-10:
                op = nn.Sequential(nn.Linear(in_size, H), nl, nn.Dropout(drpt))
            elif bn:
                op = nn.Sequential(nn.Linear(in_size, H), nl, nn.BatchNorm1d(H))
            else:   # the reference falls through every branch here (ntu_searchable.py:274-284)
                raise UnboundLocalError("local variable 'op' referenced before assignment "
                                        "(no layer recipe for drpt<1e-10 and batchnorm=False)")
            layers.append(op)
        return nn.ModuleList(layers)

    def central_params(self):
        return [{'params': self.alphas.parameters()},
                {'params': self.fusion_layers.parameters()},
                {'params': self.central_classifier.parameters()}]

    # ---- arena plumbing ---------------------------------------------------------------------
    def _named_state(self):
        d = dict(self.named_parameters())
        d.update(dict(self.named_buffers()))
        return d

    def attach(self, group: CandidateGroup, slot: int, copy_in: bool = True):
        """Make candidate ``slot`` of ``group`` the storage of this module's tensors."""
        with torch.no_grad():
            for name, t in self._named_state().items():
                if name not in group.slots[slot]:
                    continue
                v = group.view(slot, name)
                if copy_in:
                    v.copy_(t.detach().to(v.device).reshape(v.shape))
                t.data = v
        self._group, self._slot = group, slot
        return self

    def native(self, device=None, batch_max=None) -> CandidateGroup:
        """The 1-candidate group backing this module (created on first use / after .to()).  ``batch_max`` (default: keep
        the current group, or MFAS_MAX_BATCH for a new one) re-creates the group for another maximum batch -- a training
        loop asks for its own batch size, which decides the kernels that serve the group (mfas_group_create)."""
        w = self.fusion_layers[0][0].weight
        if device is None:
            device = w.device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("mfas_b200 computes on CUDA devices only; move the model with .to('cuda') "
                               "(there is no CPU fallback)")
        g = self._group
        if g is not None and w.device == device and w.data_ptr() == g.view(self._slot, "fusion_layers.0.0.weight").data_ptr() \
                and (batch_max is None or g.batch_max == int(batch_max)):
            return g
        a = self.args
        cf = np.asarray(self.conf).reshape(-1, 3)
        fl = flags_from_args(a) | self._extra_flags
        if fl & _lib.FLAG_PLAIN:
            fl &= ~_lib.FLAG_BN                        # the AV-MNIST recipe has no BatchNorm whatever args.batchnorm says
        g = CandidateGroup([cf], a.inner_representation_size, a.num_outputs, fl, device,
                           batch_max=int(batch_max or _lib.MAX_BATCH), drop_p=float(a.drpt) if a.drpt > 1e-10 else 0.0,
                           drop_seed=int(getattr(a, "dropout_seed", 0)), vid_len_ske=getattr(a, "vid_len", (8, 32))[1],
                           widths=self._widths_kw)
        self.attach(g, 0, copy_in=True)
        return g

    # ---- forward ----------------------------------------------------------------------------
    def forward(self, tensor_tuple):
        rgb, ske = tensor_tuple[0], tensor_tuple[1]          # caller passes (rgb, ske), ntu_searchable.py:208
        nr, ns = sum(self.rgbnet.widths), sum(self.skenet.widths)
        if rgb.dim() != 2 or ske.dim() != 2:
            raise ValueError("expected cached taps: [B, %d] for the second modality, [B, %d] for the first" % (nr, ns))
        multitask = bool(getattr(self.args, "multitask", False))
        vis_logits = rgb[:, nr:] if rgb.shape[1] > nr else None      # visual_classifier / skel_classifier, :213,:217
        ske_logits = ske[:, ns:] if ske.shape[1] > ns else None
        if multitask and (vis_logits is None or ske_logits is None):
            raise ValueError("args.multitask needs the cached backbone logits behind the taps (FeatureCache(logit_rgb=, "
                             "logit_ske=); FeatureCacheLoader appends them)")
        g = self.native(rgb.device)
        B = rgb.shape[0]
        if B > _lib.MAX_BATCH:
            raise ValueError(f"batch of {B} rows exceeds MFAS_MAX_BATCH={_lib.MAX_BATCH}")
        if B > g.batch_max:                                # the group was sized by a training loop with a smaller batch
            g = self.native(rgb.device, batch_max=_lib.MAX_BATCH)
        cache = FeatureCache(ske[:, :ns].contiguous(), rgb[:, :nr].contiguous(),
                             torch.zeros(B, dtype=torch.int64, device=rgb.device), getattr(self.args, "vid_len", (8, 32))[1],
                             vis_logits.contiguous() if multitask else None, ske_logits.contiguous() if multitask else None,
                             widths=self._widths_kw)
        rows = torch.arange(B, dtype=torch.int32, device=rgb.device)
        logits, _, _ = g.forward(cache, rows, train=self.training, step=self._fwd_steps)
        if self.training:
            self._fwd_steps += 1
        out = logits[self._slot].clone()
        if not multitask:
            return out
        return out, vis_logits, ske_logits                       # ntu_searchable.py:244-247


# ----------------------------------------------------------------------------------------------
# search space / weight sharing
# ----------------------------------------------------------------------------------------------
def get_possible_layer_configurations(progression_index):
    """All [ske tap, rgb tap, activation] rows of one fusion step: 4 x 4 x 2 = 32
    (/root/reference/models/search/ntu_searchable.py:105-119)."""
    return [[t, v, n] for t in range(4) for v in range(4) for n in range(2)]


_ACT_TAG = {0: '.A_relu', 1: '.A_sigmoid', 2: '.A_lrelu'}


def _shared_key(idx_layer, layer, conf_row):
    return str(idx_layer) + '.L_' + str(layer[0].in_features) + '_' + str(layer[0].out_features) + \
        _ACT_TAG.get(int(conf_row[2]), '')


def get_central_states(model, state_dict, using_dataparallel=False):
    """Store every fusion layer under '<l>.L_<in>_<out>.A_<act>' (ntu_searchable.py:123-149)."""
    net = model.module if using_dataparallel else model
    for idx_layer, layer in enumerate(net.fusion_layers):
        name = _shared_key(idx_layer, layer, net.conf[idx_layer])
        print(('Updating' if name in state_dict else 'Creating') + ' shared weight with ID: {}'.format(name))
        state_dict[name] = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    return state_dict


def set_central_states(model, state_dict, using_dataparallel=False):
    """Load shared fusion layers whose key matches (ntu_searchable.py:152-174)."""
    net = model.module if using_dataparallel else model
    for idx_layer, layer in enumerate(net.fusion_layers):
        name = _shared_key(idx_layer, layer, net.conf[idx_layer])
        if name in state_dict:
            with torch.no_grad():
                own = layer.state_dict()
                for k, v in state_dict[name].items():
                    own[k].copy_(v)          # in place: the tensors are views of the CUDA arenas
            print('Loaded shared weight with ID: {}'.format(name))


# ----------------------------------------------------------------------------------------------
# parameter initialisation without building nn.Modules (the search driver never looks at the models)
# ----------------------------------------------------------------------------------------------
def _init_tensors(group, slot, fill):
    """Visit the tensors of candidate ``slot`` in the order the reference constructor creates them
    (ntu_searchable.py:179-204): L x Linear(weight, bias) [+ BatchNorm1d], classifier, then the alphas."""
    lay = group.layouts[slot]
    bn = bool(lay.flags & _lib.FLAG_BN)
    for l in range(lay.L):
        K = lay.K[l]
        fill(f"fusion_layers.{l}.0.weight", "kaiming", K)
        fill(f"fusion_layers.{l}.0.bias", "uniform", K)
        if bn:
            fill(f"fusion_layers.{l}.2.weight", "ones", 0)
            fill(f"fusion_layers.{l}.2.bias", "zeros", 0)
            fill(f"fusion_layers.{l}.2.running_mean", "zeros", 0)
            fill(f"fusion_layers.{l}.2.running_var", "ones", 0)
    fill("central_classifier.weight", "kaiming", lay.H)
    fill("central_classifier.bias", "uniform", lay.H)
    for l in range(lay.L):
        fill(f"alphas.{l}.alpha_x", "normal", 0)


def init_host_arenas(group, host_p, host_b, slots=None):
    """torch's default initialisers applied straight to pinned host arenas, consuming the global CPU
    generator exactly like ``searchable_type(args, conf)`` would for every candidate in order -- same
    seed, same weights as the reference constructor -- without creating a single nn.Module."""
    from .host_init import init_host_arenas_fast
    if os.environ.get("MFAS_HOST_INIT") != "torch" and init_host_arenas_fast(group, host_p, host_b, _init_tensors, slots):
        return
    for c in (range(group.n) if slots is None else slots):
        def fill(name, kind, fan_in, c=c):
            arena, off, shape = group.slots[c][name]
            base = host_b if arena == "b" else host_p
            o = int(group.b_off[c] if arena == "b" else group.p_off[c]) + int(off)
            n = int(np.prod(shape)) if shape else 1
            t = base[o:o + n].view(shape)
            if kind == "kaiming":
                nn.init.kaiming_uniform_(t, a=math.sqrt(5))
            elif kind == "uniform":
                bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
                nn.init.uniform_(t, -bound, bound)
            elif kind == "ones":
                t.fill_(1.0)
            elif kind == "zeros":
                t.zero_()
            elif kind == "normal":
                nn.init.normal_(t, 0.0, 0.1)
        _init_tensors(group, c, fill)


class _ShareLayout:
    """A full call's GroupLayout with its arena offsets remapped for ONE rank's share: candidate ``mine[k]`` of the call sits
    at candidate k of ``group``; every other candidate shares one scratch slot behind the share (its draws are consumed --
    the constructor's RNG order is that of the whole call -- and overwritten by the next)."""

    def __init__(self, full, mine, group):
        self.n, self.layouts, self.slots = full.n, full.layouts, full.slots
        n_p = int(group.p_off[-1]) if group is not None else 0
        n_b = int(group.b_off[-1]) if group is not None else 0
        where = {j: k for k, j in enumerate(mine)}
        sizes_p = [int(full.p_off[c + 1] - full.p_off[c]) for c in range(full.n)]
        sizes_b = [int(full.b_off[c + 1] - full.b_off[c]) for c in range(full.n)]
        self.p_off = np.array([int(group.p_off[where[c]]) if c in where else n_p for c in range(full.n)] + [0], dtype=np.int64)
        self.b_off = np.array([int(group.b_off[where[c]]) if c in where else n_b for c in range(full.n)] + [0], dtype=np.int64)
        self.n_p = n_p + max([sizes_p[c] for c in range(full.n) if c not in where], default=0)
        self.n_b = n_b + max([sizes_b[c] for c in range(full.n) if c not in where], default=0)


_STAGING = {}


def _staging(n_p, n_b):
    """Pinned host staging arenas, kept across calls (pinning 0.5 GB costs more than filling it).  Zeroed once: the
    initialisers overwrite every parameter on every call, the alignment gaps between tensors stay zero."""
    hp, hb = _STAGING.get("p"), _STAGING.get("b")
    if hp is None or hp.numel() < n_p:
        hp = _STAGING["p"] = torch.zeros(n_p, dtype=torch.float32).pin_memory()
    if hb is None or hb.numel() < n_b:
        hb = _STAGING["b"] = torch.zeros(n_b, dtype=torch.float32).pin_memory()
    return hp[:n_p], hb[:n_b]


def init_on_device(group, base_seed, cand_ids=None):
    """Same distributions as the reference constructor, drawn on the GPU by ONE launch of the library's counter-based
    generator keyed by (base_seed, candidate id, tensor, element) (mfas_group_init_params; the ids are the ``cand_ids`` the
    group was created with): the result does not depend on which rank / group / device a candidate lands in.  Not
    bit-compatible with the reference constructor's CPU stream (use the default host initialisation for seed parity)."""
    group.init_params(base_seed)


# ----------------------------------------------------------------------------------------------
# candidate trainer
# ----------------------------------------------------------------------------------------------
def _feature_cache_of(loader, what):
    ds = getattr(loader, "dataset", None)
    if not isinstance(ds, FeatureCache):
        raise TypeError(f"dataloaders['{what}'].dataset must be a mfas_b200.FeatureCache (pre-extracted backbone "
                        f"taps); got {type(ds).__name__}. Build one with mfas_b200.cache (DESIGN.md, 'Feature cache').")
    return ds


def pass_orders(loader, first_pass, count, n_rows, device="cpu"):
    """Row orders of ``count`` consecutive passes over ``loader`` as an int32 [count, n_rows] tensor (built on
    ``device`` when the loader can: one batched sort instead of count host-side randperms)."""
    if hasattr(loader, "orders"):
        return loader.orders(first_pass, count, device).to(torch.int32)
    if hasattr(loader, "order_for_pass"):
        return torch.stack([loader.order_for_pass(first_pass + k) for k in range(count)]).to(torch.int32)
    shuffle = not isinstance(getattr(loader, "sampler", None), torch.utils.data.SequentialSampler)
    if shuffle:
        return torch.stack([torch.randperm(n_rows) for _ in range(count)]).to(torch.int32)
    return torch.arange(n_rows, dtype=torch.int32).repeat(count, 1)


def _reserve_passes(loader, count):
    return loader.take_passes(count) if hasattr(loader, "take_passes") else 0


def cosine_lrs(args, n_train, n_steps):
    """LR of every optimiser step of one candidate (LRCosineAnnealingScheduler, per batch)."""
    sched = LRCosineAnnealingScheduler(args.eta_max, args.eta_min, args.Ti, args.Tm, n_train / args.batchsize)
    return [sched.step() for _ in range(n_steps)]


def fanout_devices(args, device):
    """CUDA devices a single-process ``train_sampled_models`` call spreads its candidates over.  Default: just ``device``.
    ``args.fanout_gpus`` (or MFAS_FANOUT_GPUS in the environment) = N, "all" or a list of device indices turns the fan-out
    on: the reference's driver is ONE unseeded process (models/searchable.py:48-137, models/search/tools.py:53), so this is
    the mode in which an unmodified ``main_searchable_ntu.py`` uses the 8 GPUs of a box -- no torchrun, no collective: one
    candidate group per device, enqueued from one host thread per device, results gathered by D2H copies."""
    spec = getattr(args, "fanout_gpus", None)
    if spec is None:
        spec = os.environ.get("MFAS_FANOUT_GPUS")
    if spec is None or spec in (0, 1, "", "0", "1"):
        return [device]
    n = torch.cuda.device_count()
    if isinstance(spec, (list, tuple)):
        ids = [int(i) for i in spec]
    elif str(spec) == "all":
        ids = list(range(n))
    else:
        ids = list(range(min(int(spec), n)))
    first = device.index if device.index is not None else torch.cuda.current_device()
    ids = [first] + [i for i in ids if i != first]
    if any(i < 0 or i >= n for i in ids):
        raise ValueError(f"fanout_gpus={spec!r}: this process sees {n} CUDA device(s)")
    return [torch.device("cuda", i) for i in ids]


def _print_epoch_logs(stats, n_train, n_dev):
    for e in range(stats.shape[0]):           # same lines as train_searchable/ntu.py:78-79
        print('{} Loss: {:.4f} Acc: {:.4f}'.format('train', stats[e, 0] / n_train, stats[e, 1] / n_train))
        print('{} Loss: {:.4f} Acc: {:.4f}'.format('dev', stats[e, 2] / n_dev, stats[e, 3] / n_dev))


class TrainerSpec:
    """What distinguishes the candidate trainers of the tap sets (NTU here, MM-IMDB / AV-MNIST in their modules): everything
    else -- grouping, initialisation, sharding over ranks or devices, the device-side epoch loop -- is shared."""
    own_class = None                 # searchable_type for which no nn.Module needs to be built
    widths = None                    # (first-modality tap widths, second-modality tap widths); None = the NTU taps
    metric_name = "Acc"

    def cache_of(self, loader, what):
        return _feature_cache_of(loader, what)

    def flags(self, args):
        return flags_from_args(args)

    def check(self, args, flags, preaccuracies):
        if preaccuracies:   # the reference passes init_f1=..., which train_ntu_track_acc does not accept (:85-89)
            raise TypeError("train_ntu_track_acc() got an unexpected keyword argument 'init_f1'")
        if flags & _lib.FLAG_MULTITASK:
            # the reference calls train_ntu_track_acc WITHOUT multitask= (ntu_searchable.py:81-83) while the model returns a
            # 3-tuple (:244-247), so torch.max(output, 1) raises there; multitask training goes through train_ntu_track_acc
            raise TypeError("max() received an invalid combination of arguments - got (tuple, int): train_sampled_models "
                            "does not pass multitask to the training loop (use train_ntu_track_acc(..., multitask=True))")

    def best_init(self, preaccuracies, idx):
        return 0.0

    def result(self, best):
        return best

    def load_backbones(self, rmode, args):
        for net, cp in ((rmode.skenet, args.ske_cp), (rmode.rgbnet, args.rgb_cp)):
            fn = os.path.join(args.checkpointdir, cp)
            if os.path.isfile(fn):          # backbone weights are irrelevant once taps are cached
                net.load_state_dict(torch.load(fn))

    def log_epochs(self, stats, n_train, n_dev):
        _print_epoch_logs(stats, n_train, n_dev)

    def vid_len(self, args):
        return getattr(args, "vid_len", (8, 32))[1]


class _NTUSpec(TrainerSpec):
    pass


def train_sampled_models(sampled_configurations, searchable_type, dataloaders,
                         args, device,
                         return_model=[], premodels=[], preaccuracies=[],
                         train_only_central_params=True,
                         state_dict=dict()):
    """Train every sampled configuration and return its best dev accuracy, in input order.

    Signature and semantics of /root/reference/models/search/ntu_searchable.py:23-102: Adam(lr=eta_max,
    weight_decay=1e-4) with the per-batch cosine LR, ``args.epochs`` x (train pass, dev pass), strict-'>'
    best-dev tracking and rollback to the best weights.  All candidates of the call are trained
    concurrently by the CUDA library (sequentially only when ``args.weightsharing`` chains them
    through ``state_dict``); with torch.distributed initialised they are sharded over ranks, with
    ``args.fanout_gpus`` over the devices of this process.
    """
    spec = _NTUSpec()
    spec.own_class = Searchable_Skeleton_Image_Net
    return train_sampled(spec, train_sampled_models, sampled_configurations, searchable_type, dataloaders, args, device,
                         return_model, premodels, preaccuracies, state_dict)


def train_sampled(spec, entry, sampled_configurations, searchable_type, dataloaders, args, device,
                  return_model, premodels, preaccuracies, state_dict):
    """The body shared by the ``train_sampled_models`` of every tap set (``entry`` = that function: it receives ``last_stats``)."""
    from . import dist as mdist
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("mfas_b200.train_sampled_models needs a CUDA device (no CPU fallback)")
    import time
    timing = os.environ.get("MFAS_TIMING") == "1"       # phase timings of the host path on stderr (synchronising!)
    t_last = [time.perf_counter()]

    def lap(what):
        if timing:
            torch.cuda.synchronize(device)
            now = time.perf_counter()
            import sys
            sys.stderr.write(f"[mfas timing] {what}: {(now - t_last[0]) * 1e3:.1f} ms\n")
            t_last[0] = now
    flags = spec.flags(args)
    spec.check(args, flags, preaccuracies)
    train_host = spec.cache_of(dataloaders['train'], 'train')
    dev_host = spec.cache_of(dataloaders['dev'], 'dev')
    n_train, n_dev = len(train_host), len(dev_host)
    E, B = int(args.epochs), int(args.batchsize)
    steps = math.ceil(n_train / B)

    todo = [i for i in range(len(sampled_configurations)) if not return_model or i in return_model]
    rank, world = mdist.world()
    # Fast path: the driver never looks at the model objects (models/searchable.py:90,120), so when our own class
    # is the searchable_type no nn.Module is built at all -- parameters are initialised straight into the arenas.
    weightsharing = bool(getattr(args, "weightsharing", False))
    use_dp = bool(getattr(args, "use_dataparallel", False))
    direct = (searchable_type is spec.own_class and not premodels and not return_model and not weightsharing)
    devices = fanout_devices(args, device) if (direct and world == 1) else [device]
    dev_init = bool(getattr(args, "init_on_device", world > 1 or len(devices) > 1)) and direct
    base_seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if dev_init else 0      # torch.manual_seed governs it
    if world > 1:
        # one process per GPU (torchrun): every rank trains rank 0's list with rank 0's seed, whatever its own driver sampled
        sampled_configurations, base_seed = mdist.sync_call_inputs(sampled_configurations, base_seed)
    models = {}
    if not direct:
        # every candidate is constructed here, in order, exactly like the reference does (RNG parity)
        for idx in todo:
            rmode = searchable_type(args, sampled_configurations[idx])
            if not premodels:
                spec.load_backbones(rmode, args)
            else:
                src = premodels[idx].module if use_dp else premodels[idx]
                rmode.load_state_dict(src.state_dict())
            models[idx] = rmode

    first_tr = _reserve_passes(dataloaders['train'], len(todo) * E)
    first_dv = _reserve_passes(dataloaders['dev'], len(todo) * E)
    lrs = cosine_lrs(args, n_train, E * steps)
    drop_p = float(args.drpt) if args.drpt > 1e-10 else 0.0
    init_best = [spec.best_init(preaccuracies, i) for i in todo]        # MM-IMDB: init_f1 = preaccuracies[idx]

    # with return_model every rank needs every trained model: no sharding then
    mine = mdist.my_share(len(todo)) if not (weightsharing or return_model) else list(range(len(todo)))
    lap("setup")
    if world > 1 and getattr(args, "broadcast_cache", True):
        # rank 0's split crosses PCIe once and NVLink once (NCCL broadcast); the copy is kept with the host cache, so later
        # calls over the same loaders reuse it.  Every rank walks the same sequence of calls, so the collective stays matched.
        def resident(host):
            key = f"{device}|broadcast"
            if key not in host._device_copies:
                host._device_copies[key] = mdist.broadcast_cache(host if rank == 0 else None, device)
            return host._device_copies[key]
        train_dev, dev_dev = resident(train_host), resident(dev_host)
    else:
        train_dev = train_host.to(device)
        dev_dev = dev_host.to(device)
    lap("feature cache H2D")
    accs = torch.zeros(len(todo), dtype=torch.float64)
    all_stats = torch.zeros(len(todo), max(E, 1), 4, dtype=torch.float64)

    def make_group(js, device=device):
        confs = [np.asarray(sampled_configurations[todo[j]]).reshape(-1, 3) for j in js]
        g = CandidateGroup(confs, args.inner_representation_size, args.num_outputs, flags, device, batch_max=B,
                           drop_p=drop_p, drop_seed=int(getattr(args, "dropout_seed", 0)),
                           cand_ids=[todo[j] for j in js], vid_len_ske=spec.vid_len(args), widths=spec.widths)
        g.set_adam(0.9, 0.999, 1e-8, 1e-4)                                   # op.Adam(..., weight_decay=1e-4), :65
        return g

    def orders_for(js, loader, first, n_rows, device):
        if E == 0:
            return torch.zeros(len(js), 0, n_rows, dtype=torch.int32, device=device)
        if js == list(range(js[0], js[0] + len(js))):                   # contiguous share: one batched sort
            return pass_orders(loader, first + js[0] * E, len(js) * E, n_rows, device).view(len(js), E, n_rows)
        if hasattr(loader, "orders_of"):                                # round-robin share: one batched sort of exactly
            ids = [first + j * E + e for j in js for e in range(E)]     # this rank's passes
            return loader.orders_of(ids, device).to(torch.int32).view(len(js), E, n_rows)
        if hasattr(loader, "orders"):                                   # sort the covering range once, then pick
            lo, hi = min(js), max(js)
            allp = pass_orders(loader, first + lo * E, (hi - lo + 1) * E, n_rows, device).view(hi - lo + 1, E, n_rows)
            return allp[torch.tensor([j - lo for j in js], device=allp.device)]
        return torch.stack([pass_orders(loader, first + j * E, E, n_rows, device) for j in js])

    def run(js, g=None):
        """train the candidates todo[j], j in js, as one group"""
        g = g or make_group(js)
        for k, j in enumerate(js):
            if not direct:
                models[todo[j]].attach(g, k, copy_in=True)                   # rmode.to(device), :72
                if weightsharing:
                    set_central_states(models[todo[j]], state_dict, use_dp)
        ptr = orders_for(js, dataloaders['train'], first_tr, n_train, device)
        pdv = orders_for(js, dataloaders['dev'], first_dv, n_dev, device)
        lap("batch orders")
        stats, best, _ = g.train_run(train_dev, dev_dev, ptr, pdv, lrs, E, B, best_init=[init_best[j] for j in js])
        entry.last_engine = g.engine
        lap("train_run (enqueue + GPU)")
        stats, best = stats.cpu(), best.cpu()                               # the one D2H of the call
        g.check()
        lap("D2H of the results")
        for k, j in enumerate(js):
            accs[j] = spec.result(float(best[k]))
            all_stats[j] = stats[k]
            if args.verbose:
                print('Now training: ')
                print(sampled_configurations[todo[j]])
                spec.log_epochs(stats[k].numpy(), n_train, n_dev)
            if not direct:
                m = models[todo[j]]
                m.train(False)                                               # train_searchable/ntu.py:87
                if weightsharing:
                    get_central_states(m, state_dict, use_dp)
        return g

    def full_layout():
        return GroupLayout([np.asarray(sampled_configurations[i]).reshape(-1, 3) for i in todo], args.inner_representation_size,
                           args.num_outputs, flags, spec.vid_len(args), spec.widths)

    if direct and len(devices) > 1:
        # ---- single-process fan-out: candidate j -> devices[j % n] (the same round-robin as the multi-process mode, so both
        # modes and the 1-GPU run train identical candidates: initial weights keyed by (seed, candidate) or drawn from the
        # constructor's CPU stream, batch orders keyed by (loader seed, pass)); one host thread per device -- ctypes and the
        # CUDA runtime release the GIL while they enqueue -- and one D2H copy of the results per device.
        import threading
        nd = len(devices)
        shares = [mdist.shard(len(todo), r, nd) for r in range(nd)]
        full = hp = hb = None
        if not dev_init:                # this process trains every candidate of the call: the whole call is staged once
            full = full_layout()
            hp, hb = _staging(int(full.p_off[-1]), int(full.b_off[-1]))
            init_host_arenas(full, hp, hb)
        results, errors = [None] * nd, []

        def worker(r):
            try:
                js, dev = shares[r], devices[r]
                if not js:
                    return
                with torch.cuda.device(dev):
                    g = make_group(js, dev)
                    if dev_init:
                        init_on_device(g, base_seed, [todo[j] for j in js])
                    else:
                        for k, j in enumerate(js):
                            g.params[int(g.p_off[k]):int(g.p_off[k + 1])].copy_(hp[int(full.p_off[j]):int(full.p_off[j + 1])], non_blocking=True)
                            g.bufs[int(g.b_off[k]):int(g.b_off[k + 1])].copy_(hb[int(full.b_off[j]):int(full.b_off[j + 1])], non_blocking=True)
                    tr_d, dv_d = train_host.to(dev), dev_host.to(dev)
                    ptr = orders_for(js, dataloaders['train'], first_tr, n_train, dev)
                    pdv = orders_for(js, dataloaders['dev'], first_dv, n_dev, dev)
                    stats, best, _ = g.train_run(tr_d, dv_d, ptr, pdv, lrs, E, B, best_init=[init_best[j] for j in js])
                    entry.last_engine = g.engine
                    results[r] = (stats.cpu(), best.cpu())
                    g.check()
                    g.close()
            except BaseException as ex:          # re-raised on the calling thread
                errors.append(ex)

        threads = [threading.Thread(target=worker, args=(r,), name=f"mfas-fanout-{r}") for r in range(1, nd)]
        for t in threads:
            t.start()
        worker(0)
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        for r in range(nd):
            if results[r] is None:
                continue
            stats, best = results[r]
            for k, j in enumerate(shares[r]):
                accs[j] = spec.result(float(best[k]))
                all_stats[j] = stats[k]
        if args.verbose:
            for j in range(len(todo)):
                print('Now training: ')
                print(sampled_configurations[todo[j]])
                spec.log_epochs(all_stats[j].numpy(), n_train, n_dev)
        lap(f"fan-out over {nd} devices")
    elif direct:
        if mine:
            g = make_group(mine)
            lap("group creation")
            if dev_init:
                init_on_device(g, base_seed, [todo[j] for j in mine])
            else:
                # Reference-compatible initialisation: one pinned host arena per group, filled in the constructor's
                # RNG order for EVERY candidate of the call (other ranks' draws are consumed and dropped, so a
                # candidate gets the same weights wherever it runs), then a single H2D copy.
                full = g if len(mine) == len(todo) else full_layout()
                if full is g:
                    hp, hb = _staging(int(full.p_off[-1]), int(full.b_off[-1]))
                    # a quarter of the candidates at a time: the H2D copy of one slice overlaps the fill of the next
                    step_c = max(1, (g.n + 3) // 4)
                    for c0 in range(0, g.n, step_c):
                        c1 = min(g.n, c0 + step_c)
                        init_host_arenas(full, hp, hb, slots=range(c0, c1))
                        pa, pb = int(g.p_off[c0]), int(g.p_off[c1])
                        ba, bb = int(g.b_off[c0]), int(g.b_off[c1])
                        g.params[pa:pb].copy_(hp[pa:pb], non_blocking=True)
                        g.bufs[ba:bb].copy_(hb[ba:bb], non_blocking=True)
                else:
                    # a share of the call: this rank's candidates are written where its group's arenas expect them, every
                    # other candidate's draws land in one scratch slot behind them -- the staging arena stays the size of the
                    # share (+ one candidate) however many ranks there are, and the upload is one copy per arena
                    remap = _ShareLayout(full, mine, g)
                    hp, hb = _staging(remap.n_p, remap.n_b)
                    init_host_arenas(remap, hp, hb)
                    g.params.copy_(hp[:int(g.p_off[-1])], non_blocking=True)
                    g.bufs.copy_(hb[:int(g.b_off[-1])], non_blocking=True)
                    hp[int(g.p_off[-1]):].zero_()       # the scratch slot: the staging arenas keep zeros wherever no parameter lives
                    hb[int(g.b_off[-1]):].zero_()
            lap("parameter initialisation + H2D")
            run(mine, g)
            g.close()                       # the driver never sees the models of the direct path: free the workspace now
            del g
            lap("group teardown")
        elif not dev_init:
            # nothing to train on this rank, but the constructor's draws are consumed all the same: the CPU generators of the
            # ranks stay in lockstep, so the NEXT call still gives a candidate the same weights wherever it runs
            remap = _ShareLayout(full_layout(), [], None)
            hp, hb = _staging(remap.n_p, remap.n_b)
            init_host_arenas(remap, hp, hb)
            hp.zero_()
            hb.zero_()
    elif weightsharing:                    # candidates are chained through state_dict: one at a time
        for j in mine:
            run([j])
    elif mine:
        run(mine)

    accs = mdist.gather_results(accs, len(todo))
    lap("gather")
    real_accuracies = [accs[j].clone() for j in range(len(todo))]
    entry.last_stats = all_stats
    if return_model:
        return real_accuracies, [models[i] for i in todo]
    return real_accuracies
