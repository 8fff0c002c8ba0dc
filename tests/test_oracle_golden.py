"""Pins oracle/mfas_oracle.py (numpy restatement + hand-derived backward) to fixtures produced
by executing the unmodified reference (tests/golden/gen_golden.py)."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN_CASES, GOLDEN_DIR, WS_CASES, init_states, rel_err, sample_tensor, split_np
from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache
from oracle import mfas_oracle as O

TOL = 1e-4      # north_star: "within 1e-4 relative fp tolerance"
TRAJ_W, TRAJ_V, TRAJ_LOSS = 1e-3, 1e-2, 1e-4      # trajectories: weights rel-L2, near-zero vectors rel-L2, epoch losses (see test_gpu_parity.py)
_FLOOR = json.load(open(os.path.join(GOLDEN_DIR, "noise_floor.json")))["cases"]


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _check_final(head_state, g, prefix):
    for k, v in head_state.items():
        if k.startswith("alphas"):
            continue
        ref = g[f"{prefix}/final/{k}/sample"]
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(ref[0]), k
            continue
        vec = k.endswith(".bias") or "running" in k
        err = _rel_l2(sample_tensor(v)["sample"], ref)
        assert err < (TRAJ_V if vec else TRAJ_W), (k, err)


def _setup(cs):
    train = synthetic_ntu_cache(cs["n_train"], cs["data_seed"])
    dev = synthetic_ntu_cache(cs["n_dev"], cs["data_seed"] + 1)
    return train, dev


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_oracle_matches_reference_fixture(name):
    cs = GOLDEN_CASES[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    train, dev = _setup(cs)
    seed = int(g["meta/loader_seed"])
    ltr = FeatureCacheLoader(train, cs["B"], True, seed)
    ldv = FeatureCacheLoader(dev, cs["B"], True, seed + 50000)
    trs, dvs = split_np(train), split_np(dev)
    inits = init_states(cs["confs"], cs["H"], 60, cs["bn"], cs["drpt"], cs["model_seed"])
    E, B = cs["epochs"], cs["B"]
    for ci, conf in enumerate(cs["confs"]):
        # --- one step: logits, loss, gradients (autograd vs hand-derived)
        head = O.FusionHead(conf, cs["H"], 60, inits[ci], batchnorm=cs["bn"], drpt=cs["drpt"], alphas=cs.get("alphas", False))
        rows = ltr.order_for_pass(ci * E)[:B].numpy()
        sk, rg, y = O._taps_of(trs, rows)
        logits, tape = head.forward(sk, rg, train=True)
        loss, _ = head.ce_loss(logits, y)
        assert rel_err(logits, g[f"c{ci}/step0_logits"]) < TOL
        assert abs(float(loss) - float(g[f"c{ci}/step0_loss"])) < TOL * abs(float(g[f"c{ci}/step0_loss"]))
        grads = head.backward(logits, y, tape)
        with O.precision(np.float64):       # fp32 noise of the reference arithmetic itself (cancelling sums such as d(beta))
            h64 = O.FusionHead(conf, cs["H"], 60, inits[ci], batchnorm=cs["bn"], drpt=cs["drpt"], alphas=cs.get("alphas", False))
            l64, t64 = h64.forward(sk, rg, train=True)
            g64 = h64.backward(l64, y, t64)
        for k, v in grads.items():
            if k.startswith("alphas") and not cs.get("alphas", False):
                continue
            s = sample_tensor(v)
            ref = g[f"c{ci}/grad/{k}/sample"]
            scale = max(float(g[f"c{ci}/grad/{k}/amax"]), 1e-12)
            noise = float(np.abs(np.asarray(v, np.float64) - g64[k]).max() / scale)
            tol = max(TOL, 4 * noise)
            assert tol < 10 * TOL, (k, noise)
            assert np.abs(s["sample"] - ref).max() / scale < tol, k
            assert abs(s["s2"] - float(g[f"c{ci}/grad/{k}/s2"])) <= 2 * TOL * float(g[f"c{ci}/grad/{k}/s2"]) + 1e-20

        # --- full run: per-batch losses, dev accuracy, best-epoch rollback, final weights
        head = O.FusionHead(conf, cs["H"], 60, inits[ci], batchnorm=cs["bn"], drpt=cs["drpt"], alphas=cs.get("alphas", False))
        sched = O.CosineRestartLR(1e-3, 1e-6, cs["Ti"], 2, cs["n_train"] / B)
        per_batch = {"train": [], "dev": []}

        def orders(phase, epoch, ci=ci):
            return (ltr if phase == "train" else ldv).order_for_pass(ci * E + epoch).numpy()

        best, stats = O.train_track_acc(head, sched, trs, dvs, B, orders, E)
        n_tb = -(-cs["n_train"] // B)
        exp_train_loss = (g[f"c{ci}/train_loss"] * np.minimum(B, cs["n_train"] - B * np.arange(n_tb))).sum(1) / cs["n_train"]
        got = np.array([s["train_loss"] for s in stats])
        tol = max(TRAJ_LOSS, 2 * _FLOOR[name][ci]["loss_rel"])      # the fp32 band of the algorithm itself (tests/golden/noise_floor.py)
        assert np.abs(got - exp_train_loss).max() / np.abs(exp_train_loss).max() < tol
        exp_dev_acc = g[f"c{ci}/dev_correct"].sum(1) / cs["n_dev"]
        got_acc = np.array([s["dev_acc"] for s in stats])
        assert np.abs(got_acc - exp_dev_acc).max() <= 1.0 / cs["n_dev"] + 1e-12
        assert abs(float(best) - float(g[f"c{ci}/best_acc"])) <= 1.0 / cs["n_dev"] + 1e-12
        _check_final(head.state, g, f"c{ci}")


def test_oracle_weightsharing_matches_reference_fixture():
    """args.weightsharing: the oracle's restatement of get / set_central_states (ntu_searchable.py:123-174) and of the
    chained candidate loop (:36-97) against the unmodified reference's run (tests/golden/wsh.npz)."""
    name, cs = "wsh", WS_CASES["wsh"]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    train, dev = _setup(cs)
    seed = int(g["meta/loader_seed"])
    ltr, ldv = FeatureCacheLoader(train, cs["B"], True, seed), FeatureCacheLoader(dev, cs["B"], True, seed + 50000)
    inits = init_states(cs["confs"], cs["H"], 60, cs["bn"], cs["drpt"], cs["model_seed"])
    E, B = cs["epochs"], cs["B"]
    heads = [O.FusionHead(c, cs["H"], 60, inits[ci]) for ci, c in enumerate(cs["confs"])]
    scheds = [O.CosineRestartLR(1e-3, 1e-6, cs["Ti"], 2, cs["n_train"] / B) for _ in heads]
    shared = {}
    accs, stats = O.train_sampled_heads(heads, scheds, split_np(train), split_np(dev), B,
                                        lambda ph, ci, e: (ltr if ph == "train" else ldv).order_for_pass(ci * E + e).numpy(), E,
                                        weightsharing=True, shared=shared)
    assert sorted(shared) == list(g["meta/shared_keys"])
    n_tb = -(-cs["n_train"] // B)
    for ci in range(len(heads)):
        exp = (g[f"c{ci}/train_loss"] * np.minimum(B, cs["n_train"] - B * np.arange(n_tb))).sum(1) / cs["n_train"]
        got = np.array([s["train_loss"] for s in stats[ci]])
        assert np.abs(got - exp).max() / np.abs(exp).max() < max(TRAJ_LOSS, 2 * _FLOOR[name][ci]["loss_rel"])
        assert np.array_equal(np.array([s["dev_acc"] for s in stats[ci]]) * cs["n_dev"], g[f"c{ci}/dev_correct"].sum(1).astype(np.float64))
        assert abs(float(accs[ci]) - float(g[f"c{ci}/best_acc"])) < 1e-12
        _check_final(heads[ci].state, g, f"c{ci}")
    for key, sd in shared.items():
        for k, v in sd.items():
            if k.endswith("weight"):
                assert _rel_l2(sample_tensor(v)["sample"], g[f"shared/{key}/{k}/sample"]) < TRAJ_W, (key, k)
            elif k.endswith("num_batches_tracked"):
                assert int(v) == int(g[f"shared/{key}/{k}/sample"][0])


def test_scheduler_restart_quirk():
    """Restart fires only when Tcur/Ti hits an odd integer exactly (scheduler.py:35-38)."""
    s = O.CosineRestartLR(1e-3, 1e-6, 1, 2, 4.0)      # nbpe integral -> restart at it=4
    etas = [s.step() for _ in range(6)]
    assert etas[0] == pytest.approx(1e-3)
    assert etas[4] == pytest.approx(1e-6)
    assert s.Ti == 2 and etas[5] == pytest.approx(1e-3)
    s = O.CosineRestartLR(1e-3, 1e-6, 1, 2, 4.5)      # non-integral nbpe: keeps going past pi
    etas = [s.step() for _ in range(12)]
    assert s.Ti == 1 and etas[9] == pytest.approx(1e-3)


def test_no_recipe_for_nodrop_nobn():
    with pytest.raises(UnboundLocalError):
        O.FusionHead([[0, 0, 0]], 16, 60, {}, batchnorm=False, drpt=0.0)


def test_algorithmic_counts_match_survey():
    c = O.algorithmic_counts([[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]], 128, 60, 64)
    assert c["F_sel"] == 7680 and c["P"] == 1041468 and c["K"] == [1536, 2432, 1408, 2688]
    assert abs(c["train_bytes"] - 26.96e6) < 0.01e6 and abs(c["eval_bytes"] - 6.13e6) < 0.01e6


@pytest.mark.parametrize("name", ["cfg1", "mixL"])
def test_torch_port_matches_reference_fixture(name):
    """oracle/torch_port.py (the CPU baseline timed by bench.py) reproduces the reference run."""
    import torch
    from oracle.torch_port import FusionHeadTorch, train_candidate
    torch.set_num_threads(1)
    cs = GOLDEN_CASES[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    train, dev = _setup(cs)
    seed = int(g["meta/loader_seed"])
    ltr = FeatureCacheLoader(train, cs["B"], True, seed)
    ldv = FeatureCacheLoader(dev, cs["B"], True, seed + 50000)
    inits = init_states(cs["confs"], cs["H"], 60, cs["bn"], cs["drpt"], cs["model_seed"])
    E, B = cs["epochs"], cs["B"]
    for ci, conf in enumerate(cs["confs"]):
        m = FusionHeadTorch(conf, cs["H"], 60, cs["bn"], cs["drpt"])
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in inits[ci].items() if not k.startswith("alphas")})
        best, stats = train_candidate(
            m, (train.ske_cat, train.rgb_cat, train.labels), (dev.ske_cat, dev.rgb_cat, dev.labels),
            lambda ph, e, ci=ci: (ltr if ph == "train" else ldv).order_for_pass(ci * E + e), B, E, Ti=cs["Ti"])
        assert abs(best - float(g[f"c{ci}/best_acc"])) < 1e-12
        for k, v in m.state_dict().items():
            ref = g[f"c{ci}/final/{k}/sample"]
            got = sample_tensor(v.numpy())["sample"]
            assert np.abs(got - ref).max() <= 1e-6 * max(float(g[f"c{ci}/final/{k}/amax"]), 1e-12), k


def test_oracle_matches_reference_found_flow_multitask():
    """main_found_ntu.py flow executed by the unmodified reference (tests/golden/gen_golden_found.py): multitask
    3-head loss + alpha gates, stage 1 (1 epoch) and stage 2 (fresh Adam), rollback, test pass."""
    from helpers import FOUND_MT_CASE as cs
    g = np.load(os.path.join(GOLDEN_DIR, "found_mt.npz"))
    splits = {k: synthetic_ntu_cache(n, cs["data_seed"] + i, with_backbone_logits=True)
              for i, (k, n) in enumerate((("train", cs["n_train"]), ("dev", cs["n_dev"]), ("test", cs["n_test"])))}
    loaders = {k: FeatureCacheLoader(v, cs["B"], True, cs["loader_seed"] + 1000 * i) for i, (k, v) in enumerate(splits.items())}
    sp = {k: split_np(v) for k, v in splits.items()}
    init = init_states([cs["conf"]], cs["H"], 60, True, 0.0, cs["model_seed"])[0]
    head = O.FusionHead(cs["conf"], cs["H"], 60, init, alphas=cs["alphas"])
    rows = []
    first = 0
    for stage, epochs in ((1, 1), (2, cs["epochs"])):
        head.adam, head.t = {}, 0                                   # a fresh torch.optim.Adam per stage (main_found_ntu.py:109,133)
        sched = O.CosineRestartLR(1e-3, 1e-6, cs["Ti"], 2, cs["n_train"] / cs["B"])
        orders = lambda ph, e, first=first: loaders["train" if ph == "train" else "dev"].order_for_pass(first + e).numpy()
        best, stats = O.train_track_acc(head, sched, sp["train"], sp["dev"], cs["B"], orders, epochs, multitask=True)
        first += epochs
        for s in stats:
            rows += [("train", s["train_loss"], s["train_acc"]), ("dev", s["dev_loss"], s["dev_acc"])]
        assert abs(float(best) - float(g["interm_acc" if stage == 1 else "final_acc"])) < 1e-4 + 1.0 / cs["n_dev"]
    assert [r[0] for r in rows] == list(g["epoch_phase"])
    # the reference prints 4 decimals
    assert np.abs(np.array([r[1] for r in rows]) - g["epoch_loss"]).max() < 0.51e-4 + TRAJ_LOSS * g["epoch_loss"].max()
    assert np.abs(np.array([r[2] for r in rows]) - g["epoch_acc"]).max() <= 1.0 / cs["n_dev"] + 1e-4
    acc = O.test_track_acc(head, sp["test"], cs["B"], loaders["test"].order_for_pass(0).numpy(), multitask=True)
    assert abs(float(acc) - float(g["test_acc"])) <= 1.0 / cs["n_test"] + 1e-12
    for k, v in head.state.items():
        if k.endswith("num_batches_tracked") or k.endswith(".bias") or "running" in k:
            continue
        ref = g[f"final/{k}/sample"]
        got = sample_tensor(v)["sample"]
        assert np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30) < TRAJ_W, k


def test_mmimdb_head_matches_reference():
    """SURVEY 8(f)-1, first piece: the weighted BCE-with-logits loss, its hand-derived gradient and the F1-samples metric
    of the MM-IMDB head against the reference class (autograd gradient) and sklearn, executed by gen_golden_mmimdb.py."""
    from oracle import mmimdb_head as MH
    fx = np.load(os.path.join(GOLDEN_DIR, "mmimdb_head.npz"))
    for name in ("a", "b", "c"):
        loss, dl = MH.weighted_bce_with_logits(fx[f"{name}_logits"], fx[f"{name}_targets"], fx[f"{name}_pos_weight"])
        assert abs(float(loss) - float(fx[f"{name}_loss"])) < 1e-5 * abs(float(fx[f"{name}_loss"])), name
        ref = fx[f"{name}_dlogits"]
        assert np.abs(dl - ref).max() < 1e-4 * np.abs(ref).max(), (name, np.abs(dl - ref).max(), np.abs(ref).max())
        assert abs(MH.f1_samples(fx[f"{name}_logits"], fx[f"{name}_targets"]) - float(fx[f"{name}_f1"])) < 1e-12, name
