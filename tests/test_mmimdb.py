"""MM-IMDB searchable fusion (BASELINE.json configs[3]).

CPU part: the oracle (oracle/mmimdb_oracle.py) against the fixture produced by executing the reference's own loop, loss,
scheduler and sklearn metric (tests/golden/gen_golden_mmimdb_path.py), and the host logic of the drop-in module.
GPU part (``-m gpu``): the CUDA path through the C ABI against the oracle and the fixture.  Tolerances as in
tests/test_gpu_parity.py: every step at 1e-4 (logits, loss, gradients), trajectories at TRAJ_LOSS / samples.
"""
import os

import numpy as np
import pytest
import torch

from helpers import (D_IMAGE, D_TEXT, GOLDEN_DIR, MMIMDB_CASE, init_states, make_mmimdb_args, report_traj, sample_tensor,
                     split_np_mmimdb)
from oracle import mfas_oracle as O
from oracle import mmimdb_oracle as MO

import json

TOL = 1e-4
# Trajectories (see tests/test_gpu_parity.py): epoch losses at 1e-4, trained weights at relative L2 1e-3, F1 at one borderline
# sigmoid per epoch -- or 2 x the fp32 band of the reference algorithm itself on that trajectory where that is larger
# (tests/golden/noise_floor.py, case "mmimdb": candidate 0 of the fixture, three fusion steps trained at eta_max = 1e-2, is
# chaotic in float32 -- six correct summation orders end 1.3e-2 apart in the dev loss and 0.14 in the weights; candidate 1: 2e-6).
TRAJ_LOSS, TRAJ_W = 1e-4, 1e-3
_FLOOR = json.load(open(os.path.join(GOLDEN_DIR, "noise_floor.json")))["cases"]["mmimdb"]
DEV = "cuda:0"
WIDTHS = (D_TEXT, D_IMAGE)


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _rel_max(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _case():
    import mfas_b200.mmimdb_searchable as mm
    cs = MMIMDB_CASE
    gold = np.load(os.path.join(GOLDEN_DIR, "mmimdb_path.npz"))
    train = mm.synthetic_mmimdb_cache(cs["n_train"], cs["data_seed"])
    dev = mm.synthetic_mmimdb_cache(cs["n_dev"], cs["data_seed"] + 1)
    inits = init_states(cs["confs"], cs["H"], 23, True, 0.0, cs["model_seed"], widths=WIDTHS)
    return mm, cs, gold, train, dev, inits


def _loaders(mm, cs, train, dev, ci):
    return {"train": mm.TextImageCacheLoader(train, cs["B"], True, cs["loader_seed"] + ci),
            "dev": mm.TextImageCacheLoader(dev, cs["B"], True, cs["loader_seed"] + 50000 + ci)}


# ------------------------------------------------------------------------------------------------------------------
# CPU: oracle pinned to the executed reference loop; host logic
# ------------------------------------------------------------------------------------------------------------------
def test_oracle_matches_reference_loop():
    mm, cs, gold, train, dev, inits = _case()
    trs, dvs = split_np_mmimdb(train), split_np_mmimdb(dev)
    assert np.array_equal(gold["pos_weight"], trs["pos_weight"])
    for ci, conf in enumerate(cs["confs"]):
        loaders = _loaders(mm, cs, train, dev, ci)
        head = MO.TextImageFusionHead(conf, cs["H"], 23, inits[ci], trs["pos_weight"])
        rows = loaders["train"].order_for_pass(0)[:cs["B"]].numpy()
        tx, im, z = MO._taps_of(trs, rows)
        logits, loss, grads = head.train_step(tx, im, z, 1e-3)
        ref = gold[f"c{ci}/step/logits"]
        assert np.abs(logits - ref).max() < TOL * np.abs(ref).max()
        assert abs(float(loss) - float(gold[f"c{ci}/step/loss"])) < TOL * float(gold[f"c{ci}/step/loss"])
        for k, g in grads.items():
            r = gold[f"c{ci}/step/grad/{k}"]
            assert np.abs(g - r).max() < TOL * max(np.abs(r).max(), 1e-12), (ci, k)
        # the loop: best F1, per-epoch dev F1 (the reference prints 4 decimals), rolled-back weights
        head = MO.TextImageFusionHead(conf, cs["H"], 23, inits[ci], trs["pos_weight"])
        sched = O.CosineRestartLR(cs["eta_max"], 1e-6, cs["Ti"], 2, cs["n_train"] / cs["B"])
        orders = lambda ph, e: loaders["train" if ph == "train" else "dev"].order_for_pass(e).numpy()
        best, stats = MO.train_track_f1(head, sched, trs, dvs, cs["B"], orders, cs["epochs"])
        f1s = np.array([s["dev_f1"] for s in stats])
        fl = _FLOOR[ci]
        assert np.abs(f1s - gold[f"c{ci}/epoch_dev_f1"]).max() <= max(1.0 / cs["n_dev"], 2 * fl["dev_f1_diff"]) + 1e-4, (f1s, gold[f"c{ci}/epoch_dev_f1"])
        assert abs(float(best) - float(gold[f"c{ci}/best_f1"])) <= max(1.0 / cs["n_dev"], 2 * fl["best_f1_diff"]), (best, gold[f"c{ci}/best_f1"])
        for k, v in head.state.items():
            if k.startswith("alphas") or k.endswith("num_batches_tracked") or k.endswith(".bias") or "running" in k:
                continue
            assert _rel_l2(sample_tensor(v)["sample"], gold[f"c{ci}/final/{k}/sample"]) < max(TRAJ_W, 2 * fl["weights_rel_l2"]), (ci, k)


def test_torch_port_matches_reference_fixture():
    """oracle/torch_port_mmimdb.py (the timed CPU baseline of configs[3]) against the executed reference pieces: one step."""
    from oracle import torch_port_mmimdb as TP
    mm, cs, gold, train, dev, inits = _case()
    for ci, conf in enumerate(cs["confs"]):
        m = TP.TextImageHeadTorch(conf, cs["H"], 23)
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in inits[ci].items() if not k.startswith("alphas")})
        m.train(True)
        rows = _loaders(mm, cs, train, dev, ci)["train"].order_for_pass(0)[:cs["B"]]
        logits = m(train.ske_cat[rows], train.rgb_cat[rows])
        loss = TP.weighted_bce_with_logits(logits, train.labels[rows], train.pos_weight)
        loss.backward()
        assert np.abs(logits.detach().numpy() - gold[f"c{ci}/step/logits"]).max() < 1e-5 * np.abs(gold[f"c{ci}/step/logits"]).max()
        assert abs(loss.item() - float(gold[f"c{ci}/step/loss"])) < 1e-5 * float(gold[f"c{ci}/step/loss"])
        for k, p_ in m.named_parameters():
            r = gold[f"c{ci}/step/grad/{k}"]
            assert np.abs(p_.grad.numpy() - r).max() < 1e-4 * max(np.abs(r).max(), 1e-12), (ci, k)
    tr, ev = TP.timed_sample(cs["confs"][1], 64, 23, train.ske_cat, train.rgb_cat, train.labels, train.pos_weight, 32, 2, 2)
    assert tr > 0 and ev > 0


def test_oracle_nan_loss_escape():
    """A NaN train loss ends the run with the best F1 seen before (train_searchable/mmimdb.py:105-109)."""
    mm, cs, gold, train, dev, inits = _case()
    trs, dvs = split_np_mmimdb(train), split_np_mmimdb(dev)
    head = MO.TextImageFusionHead(cs["confs"][1], cs["H"], 23, inits[1], trs["pos_weight"])
    head.state["central_classifier.bias"][0] = np.nan
    sched = O.CosineRestartLR(1e-3, 1e-6, 1, 2, cs["n_train"] / cs["B"])
    orders = lambda ph, e: np.arange(cs["n_train"] if ph == "train" else cs["n_dev"])
    with np.errstate(all="ignore"):
        best, stats = MO.train_track_f1(head, sched, trs, dvs, cs["B"], orders, 3)
    assert float(best) == 0.0 and len(stats) == 1 and "dev_f1" not in stats[0]


def test_host_layout_init_and_errors():
    """Host-only logic of the drop-in module: layouts over the text / image tap set, arena initialisation equal to
    constructing the modules in order (and to the fixture's initial state), search space, loud errors."""
    import mfas_b200.mmimdb_searchable as mm
    from mfas_b200 import _lib
    from mfas_b200.engine import GroupLayout, plan_layout
    from mfas_b200.ntu_searchable import init_host_arenas
    cs = MMIMDB_CASE
    flags = _lib.FLAG_BN | _lib.FLAG_MULTILABEL
    lay = plan_layout(cs["confs"][0], cs["H"], 23, flags, widths=WIDTHS)
    assert [lay.K[l] for l in range(3)] == [128 + 512, 64 + 512 + 64, 128 + 512 + 64]
    assert lay.flags & _lib.FLAG_MULTILABEL
    with pytest.raises(ValueError):
        plan_layout([[2, 0, 0]], 64, 23, flags, widths=WIDTHS)            # only 2 text taps
    # algorithmic bytes with the multi-label inputs: features once, p / m / v read + written, B x C targets + C weights
    from mfas_b200.engine import algorithmic_counts
    P_ = sum(lay.K[l] * 64 + 64 + 2 * 64 for l in range(3)) + 23 * 64 + 23
    F_sel = sum(lay.d_ske[l] + lay.d_rgb[l] for l in range(3))
    cnt = algorithmic_counts(lay, 32)
    assert cnt["train_bytes"] == 4 * (32 * F_sel + 6 * P_) + 4 * (32 * 23 + 23)
    assert cnt["eval_bytes"] == 4 * (32 * F_sel + P_) + 4 * (32 * 23 + 23)
    assert len(mm.get_possible_layer_configurations(0)) == 16
    args = make_mmimdb_args(cs["H"], cs["B"], cs["epochs"])
    with pytest.raises(ValueError):
        mm.Searchable_Text_Image_Net(args, [[0, 4, 0]])
    # initialisation: arenas filled in constructor order == modules built back to back == tests/helpers.init_states
    inits = init_states(cs["confs"], cs["H"], 23, True, 0.0, cs["model_seed"], widths=WIDTHS)
    gl = GroupLayout([np.array(c) for c in cs["confs"]], cs["H"], 23, flags, widths=WIDTHS)
    hp, hb = torch.zeros(int(gl.p_off[-1])), torch.zeros(int(gl.b_off[-1]))
    torch.manual_seed(cs["model_seed"])
    init_host_arenas(gl, hp, hb)
    torch.manual_seed(cs["model_seed"])
    mods = [mm.Searchable_Text_Image_Net(args, c) for c in cs["confs"]]
    for ci, m in enumerate(mods):
        sd = m.state_dict()
        assert set(sd) == set(inits[ci])
        for k, v in sd.items():
            assert np.array_equal(v.numpy(), inits[ci][k]), k
            kind, off, shape = gl.slots[ci][k]
            if kind == "n":
                continue
            base, o = (hb, int(gl.b_off[ci])) if kind == "b" else (hp, int(gl.p_off[ci]))
            n = int(np.prod(shape)) if shape else 1
            assert np.array_equal(base[o + int(off):o + int(off) + n].numpy().reshape(shape), inits[ci][k]), k
    # no CPU path
    train = mm.synthetic_mmimdb_cache(64, 1)
    assert train.multilabel and train.labels.shape == (64, 23) and train.widths == WIDTHS
    loaders = {"train": mm.TextImageCacheLoader(train, 32), "dev": mm.TextImageCacheLoader(train, 32)}
    b = next(iter(loaders["train"]))
    assert b["text"].shape == (32, 192) and b["image"].shape == (32, 2048) and b["label"].shape == (32, 23)
    with pytest.raises(RuntimeError):
        mm.train_sampled_models([np.array(cs["confs"][1])], mm.Searchable_Text_Image_Net, loaders, args, torch.device("cpu"))
    with pytest.raises(RuntimeError):
        mods[0](b["text"], b["image"])


# ------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# ------------------------------------------------------------------------------------------------------------------
def _group(confs, H, B, keep_grads=False, C=23):
    from mfas_b200 import _lib
    from mfas_b200.engine import CandidateGroup
    g = CandidateGroup(confs, H, C, _lib.FLAG_BN | _lib.FLAG_MULTILABEL, DEV, batch_max=B, keep_grads=keep_grads, widths=WIDTHS)
    g.set_adam(0.9, 0.999, 1e-8, 1e-4)
    return g


def _close(a, b, tol, what, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = scale if scale is not None else max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / s
    assert err < tol, f"{what}: rel err {err:.3e} >= {tol:.1e}"


@pytest.mark.gpu
@pytest.mark.parametrize("H,B,nrows,confs,engine", [
    (64, 32, 32, MMIMDB_CASE["confs"], "tc"),
    (256, 64, 64, [[[1, 3, 1], [0, 0, 0]], [[1, 1, 0]]], "tc"),      # BASELINE configs[3]: inner_repr=256
    (128, 64, 50, [[[0, 1, 0], [0, 2, 1], [1, 3, 2]], [[0, 0, 1]]], "tc"),   # the 64-wide text tap at every depth, a short batch
    (256, 64, 64, [[[1, 3, 1], [0, 0, 0]], [[1, 1, 0]]], "ffma"),
    (32, 128, 100, [[[0, 2, 2], [1, 1, 1], [0, 0, 0], [1, 3, 0]]], "ffma"),
])
def test_gpu_single_step_vs_oracle(H, B, nrows, confs, engine, monkeypatch):
    """One optimiser step per candidate: logits, weighted-BCE loss, every gradient, exact-match count; then the eval-mode
    forward with its F1 statistic."""
    import mfas_b200.mmimdb_searchable as mm
    train = mm.synthetic_mmimdb_cache(160, 71)
    trs = split_np_mmimdb(train)
    inits = init_states(confs, H, 23, True, 0.0, 3, widths=WIDTHS)
    if engine == "ffma":
        monkeypatch.setenv("MFAS_ENGINE", "ffma")
    g = _group(confs, H, B, keep_grads=True)
    assert g.engine == engine            # inner_repr % 64 == 0 and batch <= 64: tensor cores, ragged text taps included
    for ci in range(g.n):
        g.load_state(ci, inits[ci])
    rows = torch.randperm(160, generator=torch.Generator().manual_seed(5))[:nrows]
    tc = train.to(DEV)
    logits, loss, exact = g.train_step(tc, rows, lr=1e-3)
    g.check()
    logits, loss, exact = logits.cpu().numpy(), loss.cpu().numpy(), exact.cpu().numpy()
    tx, im, z = MO._taps_of(trs, rows.numpy())
    heads = []
    for ci, conf in enumerate(confs):
        head = MO.TextImageFusionHead(conf, H, 23, inits[ci], trs["pos_weight"])
        with O.precision(np.float64):
            h64 = MO.TextImageFusionHead(conf, H, 23, inits[ci], trs["pos_weight"])
            l64, tape = h64.forward(tx, im, train=True)
            g64 = h64.backward(l64, z, tape)
        ol, oloss, ograds = head.train_step(tx, im, z, 1e-3)
        heads.append(head)
        _close(logits[ci], ol, TOL, f"c{ci} logits")
        assert abs(loss[ci] - float(oloss)) < TOL * float(oloss), (loss[ci], oloss)
        pred = MO._sigmoid(ol) > np.float32(0.3)
        assert int(exact[ci]) == int((pred == (z > 0.5)).all(1).sum())
        got = g.state(ci, "g")
        for k, ref in ograds.items():
            gmax = max(np.abs(ref).max(), 1e-12)
            noise = float(np.abs(ref - g64[k]).max() / gmax)          # what fp32 rounding alone does to this tensor
            tol = max(TOL, 6 * noise)      # (NOISE_X of tests/test_gpu_parity.py)
            assert tol < 20 * TOL, (k, noise)
            assert _rel_l2(got[k], g64[k]) < tol, f"c{ci} grad {k}: rel L2 {_rel_l2(got[k], g64[k]):.2e} vs float64 ground truth"
            _close(got[k], g64[k], 3 * tol, f"c{ci} grad {k} vs float64 ground truth", scale=gmax)
            _close(got[k], ref, 3 * tol + noise, f"c{ci} grad {k}", scale=gmax)
        for k, ref in head.state.items():
            if "running" in k:
                _close(g.state(ci)[k], ref, TOL, f"c{ci} {k}")
    # eval-mode forward from the oracle's post-step state: logits and the statistics of an eval pass over all rows
    for ci in range(g.n):
        g.load_state(ci, heads[ci].state)
    lg, ls, _ = g.forward(tc, rows, train=False)
    out = g.eval_pass(tc, B).cpu().numpy()
    g.check()
    for ci, head in enumerate(heads):
        ol, _ = head.forward(tx, im, train=False)
        _close(lg[ci].cpu().numpy(), ol, TOL, f"c{ci} eval logits")
        oloss, _ = head.loss_and_dlogits(ol, z)
        assert abs(float(ls[ci]) - float(oloss)) < TOL * float(oloss)
        f1_sum, loss_sum = 0.0, 0.0
        for s0 in range(0, 160, B):
            r = np.arange(s0, min(160, s0 + B))
            a, b_, zz = MO._taps_of(trs, r)
            o, _ = head.forward(a, b_, train=False)
            f1_sum += float(MO._f1_rows(o, zz).sum())
            loss_sum += float(head.loss_and_dlogits(o, zz)[0]) * len(r)
        assert abs(out[ci, 1] - f1_sum) <= 1.0 + 1e-9, (out[ci, 1], f1_sum)      # at most a borderline sigmoid or two
        assert abs(out[ci, 0] - loss_sum) < TOL * loss_sum


@pytest.mark.gpu
def test_gpu_train_sampled_models_vs_reference_fixture():
    """The drop-in entry point against what the reference's own loop produced on the same inputs and initial weights."""
    mm, cs, gold, train, dev, inits = _case()
    trs, dvs = split_np_mmimdb(train), split_np_mmimdb(dev)
    args = make_mmimdb_args(cs["H"], cs["B"], cs["epochs"], Ti=cs["Ti"], eta_max=cs["eta_max"])
    for ci, conf in enumerate(cs["confs"]):
        loaders = _loaders(mm, cs, train, dev, ci)
        torch.manual_seed(cs["model_seed"])
        # candidate ci alone, initial weights = the fixture's: construct the preceding candidates to advance the RNG
        for prev in cs["confs"][:ci]:
            mm.Searchable_Text_Image_Net(args, prev)
        f1, models = mm.train_sampled_models([np.array(conf)], mm.Searchable_Text_Image_Net, loaders, args,
                                             torch.device(DEV), return_model=[0])
        st = mm.train_sampled_models.last_stats.numpy()[0]
        assert f1[0].dtype == torch.float64 and f1[0].dim() == 0 and f1[0].device.type == "cpu"
        fl = _FLOOR[ci]
        slack = max(1.0 / cs["n_dev"], 2 * fl["dev_f1_diff"])
        report_traj(f"mmimdb c{ci} best dev F1 vs reference fixture", abs(float(f1[0]) - float(gold[f"c{ci}/best_f1"])), max(1.0 / cs["n_dev"], 2 * fl["best_f1_diff"]))
        # against the oracle trajectory: epoch losses
        head = MO.TextImageFusionHead(conf, cs["H"], 23, inits[ci], trs["pos_weight"])
        sched = O.CosineRestartLR(cs["eta_max"], 1e-6, cs["Ti"], 2, cs["n_train"] / cs["B"])
        orders = lambda ph, e: loaders["train" if ph == "train" else "dev"].order_for_pass(e).numpy()
        best, ostats = MO.train_track_f1(head, sched, trs, dvs, cs["B"], orders, cs["epochs"])
        report_traj(f"mmimdb c{ci} epoch train loss vs oracle", _rel_max(st[:, 0] / cs["n_train"], [s["train_loss"] for s in ostats]), max(TRAJ_LOSS, 2 * fl["loss_rel"]))
        report_traj(f"mmimdb c{ci} epoch dev loss vs oracle", _rel_max(st[:, 2] / cs["n_dev"], [s["dev_loss"] for s in ostats]), max(TRAJ_LOSS, 2 * fl["loss_rel"]))
        report_traj(f"mmimdb c{ci} epoch dev F1 vs reference fixture", float(np.abs(st[:, 3] / cs["n_dev"] - gold[f"c{ci}/epoch_dev_f1"]).max()), slack + 1e-4)
        m = models[0]
        assert not m.training
        sd = m.state_dict()
        # rolled-back weights against the fixture's -- when both runs picked the same best epoch (two epochs of a
        # candidate can sit within one borderline sigmoid of each other)
        if int(np.argmax(st[:, 3])) == int(np.argmax(gold[f"c{ci}/epoch_dev_f1"])):
            for k, v in sd.items():
                if k.startswith("alphas") or k.endswith("num_batches_tracked") or k.endswith(".bias") or "running" in k:
                    continue
                report_traj(f"mmimdb c{ci} final {k}", _rel_l2(sample_tensor(v.cpu().numpy())["sample"], gold[f"c{ci}/final/{k}/sample"]), max(TRAJ_W, 2 * fl["weights_rel_l2"]))
        # the returned model is rolled back to its best epoch: an eval pass over dev reproduces the best F1
        out = m.native().eval_pass(dev.to(DEV), cs["B"]).cpu().numpy()
        assert abs(out[0, 1] / cs["n_dev"] - float(f1[0])) < 1e-12
        # and model(text, image) -- the reference loop's call -- gives the oracle's eval logits for those weights
        b = next(iter(mm.TextImageCacheLoader(dev, 16, False)))
        lg = m(b["text"].to(DEV), b["image"].to(DEV)).cpu().numpy()
        oh = MO.TextImageFusionHead(conf, cs["H"], 23, {k: v.cpu().numpy() for k, v in sd.items()}, trs["pos_weight"])
        ol, _ = oh.forward([t[:16] for t in dvs["text"]], [t[:16] for t in dvs["image"]], train=False)
        _close(lg, ol, TOL, f"c{ci} model(text, image)")


@pytest.mark.gpu
def test_gpu_reference_loop_signature_and_batched_equals_solo():
    """train_mmimdb_track_f1 with the reference's argument list on one module; the same candidate trained inside a
    batched train_sampled_models call gives bit-identical statistics (candidates are independent)."""
    mm, cs, gold, train, dev, inits = _case()
    from mfas_b200.scheduler import LRCosineAnnealingScheduler
    args = make_mmimdb_args(cs["H"], cs["B"], 2, Ti=1, eta_max=cs["eta_max"])
    confs = [np.array(c) for c in cs["confs"]]
    mk = lambda: {"train": mm.TextImageCacheLoader(train, cs["B"], True, 9), "dev": mm.TextImageCacheLoader(dev, cs["B"], True, 19)}
    torch.manual_seed(1)
    f1s = mm.train_sampled_models(confs, mm.Searchable_Text_Image_Net, mk(), args, torch.device(DEV))
    batched = mm.train_sampled_models.last_stats.clone()
    assert all(0.0 < float(f) <= 1.0 for f in f1s)
    torch.manual_seed(1)
    model = mm.Searchable_Text_Image_Net(args, confs[0]).to(DEV)
    crit = mm.WeightedCrossEntropyWithLogits(train.pos_weight.numpy())
    opt = torch.optim.Adam(model.parameters(), lr=args.eta_max, weight_decay=1e-4)
    sched = LRCosineAnnealingScheduler(args.eta_max, args.eta_min, args.Ti, args.Tm, cs["n_train"] / cs["B"])
    best = mm.train_mmimdb_track_f1(model, crit, opt, sched, mk(), {"train": cs["n_train"], "dev": cs["n_dev"]},
                                    device=torch.device(DEV), num_epochs=2)
    assert isinstance(best, float) and abs(best - float(f1s[0])) < 1e-9
    assert np.allclose(mm.train_mmimdb_track_f1.last_stats.numpy(), batched[0].numpy(), rtol=1e-6, atol=0)
    assert not model.training and len(opt.state) > 0


@pytest.mark.gpu
def test_gpu_full_size_properties():
    """BASELINE configs[3] shapes (inner_repr=256, bs=64) at the dataset's size (15552 / 2608 rows,
    /root/reference/datasets/mm_imdb.py:100-105): the loss falls, F1 rises above the predict-everything level, statistics
    are self-consistent."""
    import mfas_b200.mmimdb_searchable as mm
    train, dev = mm.synthetic_mmimdb_cache(15552, 1), mm.synthetic_mmimdb_cache(2608, 2)
    args = make_mmimdb_args(256, 64, 2, Ti=1)
    loaders = {"train": mm.TextImageCacheLoader(train, 64, True, 100), "dev": mm.TextImageCacheLoader(dev, 64, True, 200)}
    torch.manual_seed(0)
    confs = [np.array([[1, 3, 0], [0, 1, 1]]), np.array([[0, 0, 0]]), np.array([[1, 2, 1], [1, 0, 0], [0, 3, 0]])]
    f1s = mm.train_sampled_models(confs, mm.Searchable_Text_Image_Net, loaders, args, torch.device(DEV))
    st = mm.train_sampled_models.last_stats.numpy()
    assert st.shape == (3, 2, 4)
    assert (st[:, 1, 0] < st[:, 0, 0]).all(), "train loss must fall from epoch 0 to 1"
    assert (st[:, :, 1] <= 15552).all() and (st[:, :, 3] <= 2608).all() and (st >= 0).all()
    assert all(float(f) > 0.5 for f in f1s), [float(f) for f in f1s]
    assert np.allclose([float(f) for f in f1s], st[:, :, 3].max(1) / 2608)
    # the oracle finishes this size in seconds: same batches, same initial weights (the direct path draws them in the
    # constructor's order, i.e. tests/helpers.init_states)
    trs, dvs = split_np_mmimdb(train), split_np_mmimdb(dev)
    inits = init_states([c.tolist() for c in confs], 256, 23, True, 0.0, 0, widths=WIDTHS)
    for ci, conf in enumerate(confs):
        head = MO.TextImageFusionHead(conf, 256, 23, inits[ci], trs["pos_weight"])
        sched = O.CosineRestartLR(1e-3, 1e-6, 1, 2, 15552 / 64)
        orders = lambda ph, e: loaders["train" if ph == "train" else "dev"].order_for_pass(ci * 2 + e).numpy()
        best, ostats = MO.train_track_f1(head, sched, trs, dvs, 64, orders, 2)
        assert abs(float(f1s[ci]) - float(best)) < 0.03, (ci, float(f1s[ci]), float(best))
        report_traj(f"mmimdb full size c{ci} epoch train loss vs oracle", _rel_max(st[ci, :, 0] / 15552, [s["train_loss"] for s in ostats]), 1e-2)
        report_traj(f"mmimdb full size c{ci} epoch dev F1 vs oracle", _rel_max(st[ci, :, 3] / 2608, [s["dev_f1"] for s in ostats]), 0.05)
        report_traj(f"mmimdb full size c{ci} best F1 vs oracle", abs(float(f1s[ci]) - float(best)), 0.03)
