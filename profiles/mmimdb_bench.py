#!/usr/bin/env python3
"""BASELINE.json configs[3] as a reported extra (not the driver's bench line): MM-IMDB text+image searchable fusion,
64 candidates x 3 epochs, inner_repr=256, bs=64, synthetic taps at the dataset's size (15552 train / 2608 dev rows), one
GPU.  Times `mfas_b200.mmimdb_searchable.train_sampled_models` end to end (host buffers in, F1 list out) after one warm-up
call (median of 3 calls) and prints one JSON line with candidate-epochs/s and the HBM-roofline fraction of the train steps
(algorithmic bytes, SURVEY.md 8(d), / wall time).  Adds `cpu_baseline`: the PyTorch-CPU port of the same step on the host cores (bounded sample).
Usage: python profiles/mmimdb_bench.py [n_candidates] [epochs] [--cpu-only]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfas_b200.mmimdb_searchable as mm  # noqa: E402
from helpers import make_mmimdb_args  # noqa: E402
from mfas_b200 import _lib  # noqa: E402
from mfas_b200.engine import algorithmic_counts, plan_layout  # noqa: E402


def cpu_baseline(confs, train, n_sample=4, train_steps=40, eval_steps=10):
    """The PyTorch-CPU port of the same step (oracle/torch_port_mmimdb.py) on the host cores, a bounded sample: the first
    ``n_sample`` candidates, ``train_steps`` + ``eval_steps`` timed steps each, scaled to a 243 + 41-step candidate-epoch."""
    from oracle import torch_port_mmimdb as TP
    torch.set_num_threads(os.cpu_count() or 1)
    steps_tr, steps_dv = -(-15552 // 64), -(-2608 // 64)
    secs = []
    for c in confs[:n_sample]:
        tr, ev = TP.timed_sample(c, 256, 23, train.ske_cat, train.rgb_cat, train.labels, train.pos_weight, 64, train_steps, eval_steps)
        secs.append(steps_tr * tr + steps_dv * ev)
    return {"value": 1.0 / (sum(secs) / len(secs)), "unit": "candidate-epochs/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle/torch_port_mmimdb.py, {n_sample} candidates x ({train_steps} train + {eval_steps} eval steps), "
                      f"scaled to {steps_tr} + {steps_dv} steps per candidate-epoch"}


def main():
    args_ = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_cand = int(args_[0]) if len(args_) > 0 else 64
    epochs = int(args_[1]) if len(args_) > 1 else 3
    if "--cpu-only" in sys.argv:                                # the baseline leg alone (no GPU needed)
        rows = mm.get_possible_layer_configurations(0)
        rng = np.random.default_rng(0)
        confs = [np.array([rows[i] for i in rng.integers(0, len(rows), size=2)]) for _ in range(n_cand)]
        print(json.dumps({"cpu_baseline": cpu_baseline(confs, mm.synthetic_mmimdb_cache(15552, 1))}))
        return
    dev = torch.device("cuda:0")
    train, devs = mm.synthetic_mmimdb_cache(15552, 1).pin(), mm.synthetic_mmimdb_cache(2608, 2).pin()
    rows = mm.get_possible_layer_configurations(0)
    rng = np.random.default_rng(0)
    confs = [np.array([rows[i] for i in rng.integers(0, len(rows), size=2)]) for _ in range(n_cand)]      # L=2 candidates
    args = make_mmimdb_args(256, 64, epochs, Ti=1)
    mk = lambda: {"train": mm.TextImageCacheLoader(train, 64, True, 100), "dev": mm.TextImageCacheLoader(devs, 64, True, 200)}
    warm = make_mmimdb_args(256, 64, 1, Ti=1)
    torch.manual_seed(0)
    # warm-up with the whole candidate list: pins the staging arenas, parks the device blocks, uploads the cache
    mm.train_sampled_models(confs, mm.Searchable_Text_Image_Net, mk(), warm, dev)
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        f1s = mm.train_sampled_models(confs, mm.Searchable_Text_Image_Net, mk(), args, dev)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt = sorted(times)[1]                                  # median of 3 calls
    flags = _lib.FLAG_BN | _lib.FLAG_MULTILABEL
    steps_tr, steps_dv = -(-15552 // 64), -(-2608 // 64)
    bytes_ce = 0.0
    for c in confs:
        cnt = algorithmic_counts(plan_layout(c, 256, 23, flags, widths=mm.WIDTHS), 64)
        bytes_ce += steps_tr * cnt["train_bytes"] + steps_dv * cnt["eval_bytes"]
    peak = 6546.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = bytes_ce * epochs / dt / 1e9
    print(json.dumps({"metric": "candidate-epochs/sec (MM-IMDB fusion, bs=64)", "value": n_cand * epochs / dt, "unit": "candidate-epochs/s",
                      "n_gpus": 1, "e2e_seconds": dt, "e2e_seconds_all": times, "engine": getattr(mm.train_sampled_models, "last_engine", "?"), "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"BASELINE configs[3]: MM-IMDB text+image searchable fusion, {n_cand} candidates x "
                                             f"{epochs} epochs, inner_repr=256, L=2, bs=64, 15552/2608 rows"},
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None},
                      "dev_f1_samples": {"min": float(min(f1s)), "max": float(max(f1s))},
                      "cpu_baseline": cpu_baseline(confs, train)}))


if __name__ == "__main__":
    main()
