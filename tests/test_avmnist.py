"""AV-MNIST searchable fusion (/root/reference/models/search/avmnist_searchable.py,
train_searchable/avmnist.py): 5 audio + 3 image taps, Linear -> activation fusion steps without BatchNorm.

CPU part: the oracle (oracle/avmnist_oracle.py) against the fixture produced by executing the reference's own class and loop
(tests/golden/gen_golden_avmnist.py), and the host logic of the drop-in module.  GPU part (``-m gpu``): the CUDA path through
the C ABI against the oracle and the fixture, tolerances as in tests/test_gpu_parity.py.
"""
import argparse
import os

import numpy as np
import pytest
import torch

from helpers import AVMNIST_CASE as CS, GOLDEN_DIR, report_traj, sample_tensor
from oracle import avmnist_oracle as AO
from oracle import mfas_oracle as O

TOL, TRAJ_W, TRAJ_V, TRAJ_LOSS = 1e-4, 1e-3, 1e-2, 1e-4
DEV = "cuda:0"


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _args(**kw):
    d = dict(inner_representation_size=CS["H"], num_outputs=10, channels=CS["channels"], drpt=0.0, batchnorm=False, alphas=CS["alphas"],
             multitask=False, weightsharing=False, batchsize=CS["B"], checkpointdir="/nonexistent", audio_cp="aud", rgb_cp="rgb", eta_max=1e-3,
             eta_min=1e-6, Ti=CS["Ti"], Tm=2, use_dataparallel=False, verbose=False, epochs=CS["epochs"])
    d.update(kw)
    return argparse.Namespace(**d)


def _case():
    import mfas_b200.avmnist_searchable as av
    gold = np.load(os.path.join(GOLDEN_DIR, "avmnist.npz"))
    train = av.synthetic_avmnist_cache(CS["n_train"], CS["data_seed"], CS["channels"])
    dev = av.synthetic_avmnist_cache(CS["n_dev"], CS["data_seed"] + 1, CS["channels"])
    loaders = {"train": av.AudioImageCacheLoader(train, CS["B"], True, CS["loader_seed"]),
               "dev": av.AudioImageCacheLoader(dev, CS["B"], True, CS["loader_seed"] + 50000)}
    inits = [{k[len(f"c{ci}/init/"):]: gold[k] for k in gold.files if k.startswith(f"c{ci}/init/")} for ci in range(len(CS["confs"]))]
    return av, gold, train, dev, loaders, inits


def _check_final(state, gold, ci, what):
    for k, v in state.items():
        ref = gold[f"c{ci}/final/{k}/sample"]
        got = sample_tensor(np.asarray(v.cpu() if torch.is_tensor(v) else v))["sample"]
        if k.startswith("alphas"):
            report_traj(f"{what} c{ci} final {k} (absolute)", float(np.abs(got - ref).max()), 1e-4)
        else:
            report_traj(f"{what} c{ci} final {k}", _rel_l2(got, ref), TRAJ_V if k.endswith(".bias") else TRAJ_W)


# ------------------------------------------------------------------------------------------------------------------
# CPU
# ------------------------------------------------------------------------------------------------------------------
def test_oracle_matches_the_executed_reference():
    av, gold, train, dev, loaders, inits = _case()
    trs, dvs = AO.split_of(train), AO.split_of(dev)
    E, B = CS["epochs"], CS["B"]
    for ci, conf in enumerate(CS["confs"]):
        head = AO.AudioImageFusionHead(conf, CS["H"], 10, inits[ci], CS["channels"], alphas=CS["alphas"])
        rows = loaders["train"].order_for_pass(ci * E)[:B].numpy()
        sk, rg, y = O._taps_of(trs, rows)
        logits, tape = head.forward(sk, rg, train=True)
        loss, _ = head.ce_loss(logits, y)
        assert np.abs(logits - gold[f"c{ci}/step0_logits"]).max() < TOL * np.abs(gold[f"c{ci}/step0_logits"]).max()
        assert abs(float(loss) - float(gold[f"c{ci}/step0_loss"])) < TOL * float(gold[f"c{ci}/step0_loss"])
        grads = head.backward(logits, y, tape)
        for k, v in grads.items():
            scale = max(float(gold[f"c{ci}/grad/{k}/amax"]), 1e-12)
            assert np.abs(sample_tensor(v)["sample"] - gold[f"c{ci}/grad/{k}/sample"]).max() / scale < TOL, (ci, k)
        head = AO.AudioImageFusionHead(conf, CS["H"], 10, inits[ci], CS["channels"], alphas=CS["alphas"])
        sched = O.CosineRestartLR(1e-3, 1e-6, CS["Ti"], 2, CS["n_train"] / B)
        best, stats = AO.train_track_acc(head, sched, trs, dvs, B, lambda ph, e, ci=ci: loaders[ph].order_for_pass(ci * E + e).numpy(), E)
        assert abs(float(best) - float(gold[f"c{ci}/best_acc"])) < 1e-12
        _check_final(head.state, gold, ci, "avmnist oracle")


def test_module_mirrors_the_reference_class():
    av, gold, train, dev, loaders, inits = _case()
    torch.manual_seed(CS["model_seed"])
    for ci, conf in enumerate(CS["confs"]):
        m = av.Searchable_Audio_Image_Net(_args(), np.array(conf))
        sd = m.state_dict()
        assert sorted(sd.keys()) == sorted(inits[ci].keys())
        for k, v in sd.items():                      # same keys, shapes and -- for a given torch seed -- the same initial values
            assert np.array_equal(v.numpy(), inits[ci][k]), (ci, k)
        assert [type(x).__name__ for x in m.fusion_layers[0]] == ["Linear", "ReLU" if conf[0][2] == 0 else "Sigmoid" if conf[0][2] == 1 else "LeakyReLU"]
        assert m.audnet is m.skenet and len(m.central_params()) == 3
    assert len(av.get_possible_layer_configurations(0)) == 30
    with pytest.raises(ValueError):
        av.Searchable_Audio_Image_Net(_args(channels=24), np.array([[0, 0, 0]]))          # tap widths must be multiples of 32
    with pytest.raises(ValueError):
        av.Searchable_Audio_Image_Net(_args(), np.array([[5, 0, 0]]))
    b = next(iter(loaders["train"]))
    assert set(b) == {"image", "audio", "label"} and b["image"].shape == (CS["B"], 7 * CS["channels"]) and b["audio"].shape == (CS["B"], 31 * CS["channels"])


# ------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["tc", "ffma"])
def test_gpu_single_step_vs_oracle_and_fixture(engine, monkeypatch):
    from mfas_b200 import _lib
    from mfas_b200.engine import CandidateGroup
    av, gold, train, dev, loaders, inits = _case()
    if engine == "ffma":
        monkeypatch.setenv("MFAS_ENGINE", "ffma")
    flags = _lib.FLAG_PLAIN | (_lib.FLAG_ALPHAS if CS["alphas"] else 0)
    g = CandidateGroup([np.array(c) for c in CS["confs"]], CS["H"], 10, flags, DEV, batch_max=CS["B"], keep_grads=True,
                       widths=av.tap_widths(CS["channels"]))
    assert g.engine == engine
    g.set_adam(0.9, 0.999, 1e-8, 1e-4)
    for ci in range(g.n):
        g.load_state(ci, inits[ci])
    E, B = CS["epochs"], CS["B"]
    rows = torch.stack([loaders["train"].order_for_pass(ci * E)[:B] for ci in range(g.n)])
    logits, loss, correct = g.train_step(train.to(DEV), rows, lr=1e-3)
    g.check()
    trs = AO.split_of(train)
    for ci, conf in enumerate(CS["confs"]):
        head = AO.AudioImageFusionHead(conf, CS["H"], 10, inits[ci], CS["channels"], alphas=CS["alphas"])
        sk, rg, y = O._taps_of(trs, rows[ci].numpy())
        ol, oloss, ograds = head.train_step(sk, rg, y, 1e-3)
        lg = logits[ci].cpu().numpy()
        assert np.abs(lg - gold[f"c{ci}/step0_logits"]).max() < TOL * np.abs(gold[f"c{ci}/step0_logits"]).max()
        assert abs(float(loss[ci]) - float(gold[f"c{ci}/step0_loss"])) < TOL * float(gold[f"c{ci}/step0_loss"])
        assert int(correct[ci]) == int((ol.argmax(1) == y).sum())
        got = g.state(ci, "g")
        for k, ref in ograds.items():
            scale = max(float(gold[f"c{ci}/grad/{k}/amax"]), 1e-12)
            assert np.abs(sample_tensor(got[k])["sample"] - gold[f"c{ci}/grad/{k}/sample"]).max() / scale < TOL, (ci, k)
            assert _rel_l2(got[k], ref) < TOL, (ci, k, _rel_l2(got[k], ref))
        after = g.state(ci)
        for k in ograds:                              # one Adam step from zero moments: p - lr * sign-like update, held to the oracle's
            well = np.abs(ograds[k]) > 0.25 * np.abs(ograds[k]).max()
            if well.any():
                assert np.abs(after[k][well] - head.state[k][well]).max() < 5 * TOL * max(np.abs(head.state[k]).max(), 1e-12), (ci, k)


@pytest.mark.gpu
def test_gpu_train_sampled_models_vs_reference_fixture():
    av, gold, train, dev, loaders, inits = _case()
    confs = [np.array(c) for c in CS["confs"]]
    torch.manual_seed(CS["model_seed"])
    accs, models = av.train_sampled_models(confs, av.Searchable_Audio_Image_Net, loaders, _args(), torch.device(DEV),
                                           return_model=list(range(len(confs))))
    assert av.train_sampled_models.last_engine == "tc"
    for ci in range(len(confs)):
        assert accs[ci].dtype == torch.float64 and accs[ci].dim() == 0 and accs[ci].device.type == "cpu"
        report_traj(f"avmnist c{ci} best dev accuracy (samples)", abs(float(accs[ci]) - float(gold[f"c{ci}/best_acc"])) * CS["n_dev"], 1 + 1e-9)
        assert not models[ci].training
        _check_final(models[ci].state_dict(), gold, ci, "avmnist")
    # the direct path (no nn.Module built) draws the same initial weights from the constructor's stream: same accuracies
    loaders2 = {"train": av.AudioImageCacheLoader(train, CS["B"], True, CS["loader_seed"]),
                "dev": av.AudioImageCacheLoader(dev, CS["B"], True, CS["loader_seed"] + 50000)}
    torch.manual_seed(CS["model_seed"])
    accs2 = av.train_sampled_models(confs, av.Searchable_Audio_Image_Net, loaders2, _args(), torch.device(DEV))
    assert [float(a) for a in accs2] == [float(a) for a in accs]


@pytest.mark.gpu
def test_gpu_reference_loop_signature(capsys):
    """train_avmnist_track_acc / test_avmnist_track_acc with the reference's argument lists on one module."""
    from mfas_b200.scheduler import LRCosineAnnealingScheduler
    av, gold, train, dev, loaders, inits = _case()
    loaders["test"] = av.AudioImageCacheLoader(dev, CS["B"], False, 0)
    args = _args()
    torch.manual_seed(CS["model_seed"])
    model = av.Searchable_Audio_Image_Net(args, np.array(CS["confs"][0])).to(DEV)
    opt = torch.optim.Adam(model.central_params(), lr=args.eta_max, weight_decay=1e-4)
    sch = LRCosineAnnealingScheduler(args.eta_max, args.eta_min, args.Ti, args.Tm, CS["n_train"] / CS["B"])
    sizes = {"train": CS["n_train"], "dev": CS["n_dev"], "test": CS["n_dev"]}
    best = av.train_avmnist_track_acc(model, [torch.nn.CrossEntropyLoss()], opt, sch, loaders, sizes, device=torch.device(DEV),
                                      num_epochs=CS["epochs"], multitask=False)
    out = capsys.readouterr().out
    assert out.count("train Acc: ") == CS["epochs"] and out.count("dev Acc: ") == CS["epochs"] and "Loss" not in out
    assert abs(float(best) - float(gold["c0/best_acc"])) <= 1.0 / CS["n_dev"] + 1e-12
    acc = av.test_avmnist_track_acc(model, loaders, sizes, device=torch.device(DEV))
    assert acc.dtype == torch.float64 and abs(float(acc) - float(best)) < 1e-12          # rolled back to the best dev epoch; test split = dev split
    lg = model((train.rgb_cat[:8].to(DEV), train.ske_cat[:8].to(DEV)))                    # (image, sound), avmnist_searchable.py:207
    assert lg.shape == (8, 10)
