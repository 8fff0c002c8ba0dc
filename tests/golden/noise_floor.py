#!/usr/bin/env python3
"""fp32 noise floor of the REFERENCE ALGORITHM ITSELF on every trajectory the parity tests compare.

    python tests/golden/noise_floor.py        # writes tests/golden/noise_floor.json  (CPU, ~1 min)

The pinned oracle (oracle/mfas_oracle.py) is run over each trajectory in float64 (ground truth) and as an ENSEMBLE of float32
realizations -- the plain one plus the same arithmetic with the fusion-step products summed in 2, 3, 5, 8 and 16 column ranges
(``oracle.summation_order``: what any split-K GEMM does) -- from the same initial weights, batch orders and learning rates.
What differs between the runs is rounding only, so the largest distance of a realization from the float64 run is what ANY
correct fp32 implementation (the reference on another BLAS, the CUDA path) may be away from the committed fixture.  The band is
heavy-tailed: the rounding of a pre-activation decides on which side of the ReLU kink it falls, and one flipped derivative moves
an epoch loss by 1e-5 .. 5e-4 (traj_cfg2: dev loss 2e-6 in six realizations, 1.4e-4 in two, 4.7e-4 in one).  The tests bound the
CUDA path by

    trained weights : relative L2 <= TRAJ_W    = 1e-3   (measured band: <= 1.4e-4)
    epoch losses    : relative    <= TRAJ_LOSS = 1e-4, or 2 x the band measured here for that trajectory where that is larger
                      (train and dev losses share one band: both are mean cross-entropies under the same weights)
    accuracy        : <= 1 sample (measured band: 0)

and read the per-case band from the JSON this script writes; nothing is loosened by prose.
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import FOUND_CONFS, GOLDEN_CASES, WS_CASES, init_states, split_np  # noqa: E402
from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache  # noqa: E402
from oracle import mfas_oracle as O  # noqa: E402

LOADER_SEED = 100
WEIGHTS = lambda k: k.endswith("0.weight") or k.endswith("2.weight") or k == "central_classifier.weight"


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


SPLITS = (1, 2, 3, 5, 8, 16)         # the float32 ensemble


def run(cs, dtype, loader_seeds=(LOADER_SEED, LOADER_SEED + 50000), data_seeds=None, init_seed=None, splits=1):
    train = synthetic_ntu_cache(cs["n_train"], data_seeds[0] if data_seeds else cs["data_seed"])
    dev = synthetic_ntu_cache(cs["n_dev"], data_seeds[1] if data_seeds else cs["data_seed"] + 1)
    ltr = FeatureCacheLoader(train, cs["B"], True, loader_seeds[0])
    ldv = FeatureCacheLoader(dev, cs["B"], True, loader_seeds[1])
    inits = init_states(cs["confs"], cs["H"], 60, cs["bn"], cs["drpt"], cs["model_seed"] if init_seed is None else init_seed)
    E, B = cs["epochs"], cs["B"]
    with O.precision(dtype), O.summation_order(splits):
        heads = [O.FusionHead(c, cs["H"], 60, inits[ci], batchnorm=cs["bn"], alphas=cs.get("alphas", False)) for ci, c in enumerate(cs["confs"])]
        scheds = [O.CosineRestartLR(1e-3, 1e-6, cs["Ti"], 2, cs["n_train"] / B) for _ in heads]
        orders = lambda ph, ci, e: (ltr if ph == "train" else ldv).order_for_pass(ci * E + e).numpy()
        accs, stats = O.train_sampled_heads(heads, scheds, split_np(train), split_np(dev), B, orders, E,
                                            weightsharing=cs.get("weightsharing", False))
    return accs, stats, [{k: np.array(v, np.float64) for k, v in h.state.items()} for h in heads]


def _relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def floor_of(cs, **kw):
    a64, s64, w64 = run(cs, np.float64, **kw)
    out = [dict(weights_rel_l2=0.0, vectors_rel_l2=0.0, loss_rel=0.0, train_loss_rel=0.0, dev_loss_rel=0.0, dev_correct_diff=0.0,
                train_correct_diff=0.0, best_acc_diff=0.0) for _ in cs["confs"]]
    for sp in SPLITS:
        a32, s32, w32 = run(cs, np.float32, splits=sp, **kw)
        for ci, o in enumerate(out):
            col = lambda st, k: [e[k] for e in st[ci]]
            real = dict(
                weights_rel_l2=max(rel_l2(w32[ci][k], w64[ci][k]) for k in w32[ci] if WEIGHTS(k)),
                vectors_rel_l2=max(rel_l2(w32[ci][k], w64[ci][k]) for k in w32[ci] if k.endswith(".bias") or "running" in k),
                train_loss_rel=_relmax(col(s32, "train_loss"), col(s64, "train_loss")),          # the tests' metric: max |diff| / max |ref|
                dev_loss_rel=_relmax(col(s32, "dev_loss"), col(s64, "dev_loss")),
                dev_correct_diff=float(np.abs(np.array(col(s32, "dev_acc")) - np.array(col(s64, "dev_acc"))).max() * cs["n_dev"]),
                train_correct_diff=float(np.abs(np.array(col(s32, "train_acc")) - np.array(col(s64, "train_acc"))).max() * cs["n_train"]),
                best_acc_diff=abs(float(a32[ci]) - float(a64[ci])))
            real["loss_rel"] = max(real["train_loss_rel"], real["dev_loss_rel"])
            for k, v in real.items():
                o[k] = max(o[k], v)
    return out


def floor_mmimdb():
    """The MM-IMDB fixture trajectory (tests/golden/gen_golden_mmimdb_path.py): 40 Adam steps at eta_max = 1e-2.  Candidate 0
    (three fusion steps) is chaotic at that learning rate -- the band is ~1e-2 in the losses and ~0.1 in the weights."""
    from helpers import D_IMAGE, D_TEXT, MMIMDB_CASE as cs, split_np_mmimdb
    import mfas_b200.mmimdb_searchable as mm
    from oracle import mmimdb_oracle as MO
    widths = (D_TEXT, D_IMAGE)
    train, dev = mm.synthetic_mmimdb_cache(cs["n_train"], cs["data_seed"]), mm.synthetic_mmimdb_cache(cs["n_dev"], cs["data_seed"] + 1)
    trs, dvs = split_np_mmimdb(train), split_np_mmimdb(dev)
    inits = init_states(cs["confs"], cs["H"], 23, True, 0.0, cs["model_seed"], widths=widths)
    out = []
    for ci, conf in enumerate(cs["confs"]):
        ltr = mm.TextImageCacheLoader(train, cs["B"], True, cs["loader_seed"] + ci)
        ldv = mm.TextImageCacheLoader(dev, cs["B"], True, cs["loader_seed"] + 50000 + ci)

        def one(dt, sp):
            with O.precision(dt), O.summation_order(sp):
                head = MO.TextImageFusionHead(conf, cs["H"], 23, inits[ci], trs["pos_weight"])
                sch = O.CosineRestartLR(cs["eta_max"], 1e-6, cs["Ti"], 2, cs["n_train"] / cs["B"])
                best, st = MO.train_track_f1(head, sch, trs, dvs, cs["B"], lambda ph, e: (ltr if ph == "train" else ldv).order_for_pass(e).numpy(), cs["epochs"])
            return float(best), st, {k: np.array(v, np.float64) for k, v in head.state.items()}

        b64, s64, w64 = one(np.float64, 1)
        o = dict(weights_rel_l2=0.0, loss_rel=0.0, dev_f1_diff=0.0, best_f1_diff=0.0)
        for sp in SPLITS:
            b, s_, w = one(np.float32, sp)
            col = lambda st, k: [e[k] for e in st]
            o["weights_rel_l2"] = max(o["weights_rel_l2"], max(rel_l2(w[k], w64[k]) for k in w if WEIGHTS(k)))
            o["loss_rel"] = max(o["loss_rel"], _relmax(col(s_, "train_loss"), col(s64, "train_loss")), _relmax(col(s_, "dev_loss"), col(s64, "dev_loss")))
            o["dev_f1_diff"] = max(o["dev_f1_diff"], float(np.abs(np.array(col(s_, "dev_f1")) - np.array(col(s64, "dev_f1"))).max()))
            o["best_f1_diff"] = max(o["best_f1_diff"], abs(b - b64))
        out.append(o)
    return out


def main():
    res = {}
    for name, cs in {**GOLDEN_CASES, **WS_CASES}.items():
        res[name] = floor_of(cs)
    # tests/test_gpu_parity.py::test_run_vs_oracle_trajectory_cfg2_shapes (conf 4, H=128, B=64, 3 epochs, two LR restarts)
    traj = dict(confs=[FOUND_CONFS[4]], H=128, B=64, n_train=448, n_dev=192, epochs=3, bn=True, drpt=0.0, Ti=1, model_seed=1, data_seed=5)
    res["traj_cfg2"] = floor_of(traj, loader_seeds=(7, 8), data_seeds=(5, 6))
    worst = {k: max(max(c[k] for c in v) for v in res.values()) for k in res["cfg2"][0]}
    res["mmimdb"] = floor_mmimdb()
    out = dict(how="oracle/mfas_oracle.py: the largest distance of six float32 realizations (summation orders %s) from the float64 run of the "
                   "same trajectory (tests/golden/noise_floor.py)" % (SPLITS,), worst=worst, cases=res)
    path = os.path.join(os.environ.get("MFAS_GOLDEN_OUT", HERE), "noise_floor.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(worst, indent=1))


if __name__ == "__main__":
    main()
