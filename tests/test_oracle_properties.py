"""Size-independent properties of the oracles (CPU): the hand-derived backward against central finite differences in
float64 (both heads), pooling linearity, metric / loss bounds, scheduler range, batch orders are permutations.  These do
not need the reference: they hold for any correct restatement, and the CUDA path is held to the same oracles."""
import numpy as np
import torch

from helpers import D_IMAGE, D_TEXT, init_states
from oracle import mfas_oracle as O
from oracle import mmimdb_head as MH
from oracle import mmimdb_oracle as MO
from oracle import pooling as P


def _fd_check(make_head, taps, labels, n_probe=6, eps=1e-5):
    """central differences of the mean loss in float64 against head.backward, on a few entries of every tensor"""
    rng = np.random.default_rng(0)
    with O.precision(np.float64):
        head = make_head()
        head.state = {k: np.asarray(v, np.float64) if np.asarray(v).dtype.kind == "f" else v for k, v in head.state.items()}
        logits, tape = head.forward(*taps, train=True)
        grads = head.backward(logits, labels, tape)
        s0 = {k: np.array(v, copy=True) for k, v in head.state.items()}
        worst = 0.0
        for name, g in grads.items():
            flat = g.reshape(-1)
            for idx in rng.choice(flat.size, size=min(n_probe, flat.size), replace=False):
                vals = []
                for sgn in (+1.0, -1.0):
                    head.state = {k: np.array(v, copy=True) for k, v in s0.items()}
                    head.state[name].reshape(-1)[idx] += sgn * eps
                    lg, _ = head.forward(*taps, train=True)
                    vals.append(float(head.loss_and_dlogits(lg, labels)[0]))
                fd = (vals[0] - vals[1]) / (2 * eps)
                scale = max(abs(fd), abs(float(flat[idx])), 1e-6)
                worst = max(worst, abs(fd - float(flat[idx])) / scale)
                assert abs(fd - float(flat[idx])) <= 2e-4 * scale + 1e-9, (name, int(idx), fd, float(flat[idx]))
    return worst


def test_hand_derived_backward_matches_finite_differences_ce_head():
    conf = [[0, 0, 1], [1, 0, 2], [0, 1, 1]]                 # sigmoid / leaky-ReLU (smooth where probed), BatchNorm in train mode
    H, B, C = 16, 7, 60
    rng = np.random.default_rng(1)
    ske = [np.abs(rng.standard_normal((B, w))) for w in (128, 256, 1024, 512)]
    rgb = [np.abs(rng.standard_normal((B, w))) for w in (512, 1024, 2048, 2048)]
    y = rng.integers(0, C, B)
    init = init_states([conf], H, C, True, 0.0, 4)[0]
    _fd_check(lambda: O.FusionHead(conf, H, C, init), (ske, rgb), y)
    # with the modality gates: d(alpha) too
    _fd_check(lambda: O.FusionHead(conf, H, C, init, alphas=True), (ske, rgb), y)


def test_hand_derived_backward_matches_finite_differences_multilabel_head():
    conf = [[1, 2, 1], [0, 0, 2]]
    H, B, C = 16, 6, 23
    rng = np.random.default_rng(2)
    text = [np.abs(rng.standard_normal((B, w))) for w in D_TEXT]
    image = [np.abs(rng.standard_normal((B, w))) for w in D_IMAGE]
    z = (rng.random((B, C)) < 0.2).astype(np.float64)
    q = rng.random(C) * 6 + 0.5
    init = init_states([conf], H, C, True, 0.0, 5, widths=(D_TEXT, D_IMAGE))[0]
    _fd_check(lambda: MO.TextImageFusionHead(conf, H, C, init, q), (text, image), z)


def test_pooling_is_a_mean():
    rng = np.random.default_rng(3)
    x, y = rng.standard_normal((3, 5, 4, 6)).astype(np.float32), rng.standard_normal((3, 5, 4, 6)).astype(np.float32)
    assert np.allclose(P.global_pool(2 * x + y), 2 * P.global_pool(x) + P.global_pool(y), atol=1e-6)
    assert np.array_equal(P.global_pool(np.full((2, 3, 7), 1.5, np.float32)), np.full((2, 3), 1.5, np.float32))
    v = rng.standard_normal((4, 9)).astype(np.float32)
    assert np.array_equal(P.global_pool(v), v)                  # a vector tap: view(B, C, 1).mean(2)


def test_multilabel_loss_and_metric_bounds():
    rng = np.random.default_rng(4)
    z = (rng.random((12, 23)) < 0.2).astype(np.float32)
    z[0] = 0
    x = rng.standard_normal((12, 23)).astype(np.float32) * 3
    q = (rng.random(23) * 5 + 0.5).astype(np.float32)
    loss, dl = MH.weighted_bce_with_logits(x, z, q)
    assert loss > 0 and dl.shape == x.shape
    assert ((dl <= 0) | (z == 0)).all() and ((dl >= 0) | (z == 1)).all()      # positives pull logits up, negatives down
    f1 = MH.f1_samples(x, z)
    assert 0.0 <= f1 <= 1.0
    perfect = np.where(z > 0.5, 5.0, -5.0).astype(np.float32)
    rows = MO._f1_rows(perfect, z)
    assert rows[0] == 0.0 and (rows[1:][z[1:].sum(1) > 0] == 1.0).all()        # empty truth + empty prediction scores 0 (sklearn's 0/0)
    assert MH.f1_samples(np.full_like(x, -9.0), z) == 0.0                     # nothing predicted


def test_scheduler_range_and_orders_are_permutations():
    from mfas_b200.cache import hashed_orders
    sch = O.CosineRestartLR(1e-3, 1e-6, 1, 2, 10)
    lrs = [sch.step() for _ in range(200)]
    assert max(lrs) <= 1e-3 + 1e-12 and min(lrs) >= 1e-6 - 1e-12
    assert lrs[0] == 1e-3 and any(lrs[i + 1] > lrs[i] for i in range(199)), "warm restarts raise the LR again"
    o = hashed_orders(5, 0, 4, 97)
    assert all(torch.equal(torch.sort(r).values, torch.arange(97)) for r in o)
    assert not torch.equal(o[0], o[1])
