"""CPU oracle of the AV-MNIST searchable fusion network (TEST INFRASTRUCTURE ONLY -- see oracle/mfas_oracle.py).

Restates /root/reference/models/search/avmnist_searchable.py:184-297 (``Searchable_Audio_Image_Net``: 5 audio taps
``channels * {1,2,4,8,16}``, 3 image taps ``channels * {1,2,4}``, fusion steps Linear -> activation [-> Dropout] WITHOUT
BatchNorm, :276-285) and the loop of train_searchable/avmnist.py:14-85 (softmax-CE, best-dev rollback) on top of the NTU
oracle's arithmetic: only the tap widths and the layer recipe differ.

Parity status: PINNED to tests/golden/avmnist.npz, produced by executing the reference's own class and loop
(tests/golden/gen_golden_avmnist.py).
"""
from __future__ import annotations

import numpy as np

from .mfas_oracle import FusionHead, train_track_acc


def tap_widths(channels):
    c = int(channels)
    return (c, 2 * c, 4 * c, 8 * c, 16 * c), (c, 2 * c, 4 * c)          # avmnist_searchable.py:288-292


class AudioImageFusionHead(FusionHead):
    """One AV-MNIST candidate: the audio modality takes FusionHead's ``ske`` role, the image modality its ``rgb`` role."""

    def __init__(self, conf, H, C, state, channels, drpt=0.0, alphas=False, **kw):
        super().__init__(conf, H, C, state, batchnorm=False, drpt=drpt, alphas=alphas, plain=True, widths=tap_widths(channels), **kw)
        assert self.conf[:, 0].max() < 5 and self.conf[:, 1].max() < 3


def split_of(cache):
    """FeatureCache (audio in the first-modality slot, image in the second) -> the dict the NTU loop restatement walks."""
    return dict(ske=[t.numpy() for t in cache.ske_taps()], rgb=[t.numpy() for t in cache.rgb_taps()], labels=cache.labels.numpy())


train_track_acc = train_track_acc          # train_avmnist_track_acc == train_ntu_track_acc up to what it prints (avmnist.py:14-85)
