#!/usr/bin/env python3
"""Benchmark of the MFAS candidate-training hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

metric : candidate-epochs/sec -- one candidate-epoch = ceil(N_train/B) Adam steps + ceil(N_dev/B) eval
         steps of one fusion head (reference: models/search/train_searchable/ntu.py:20-22).
workload (BASELINE.json configs[1]): found conf 4, inner_repr=128, batchnorm, bs=64, synthetic NTU-shaped
         cache N_train=10240 / N_dev=5120, `--candidates` heads per GPU trained concurrently for `--epochs`.
step   : one pass of the hot path = one train_run over all candidates of the rank (E epochs).
value  : whole-job throughput with inputs resident in HBM (CUDA events, max over ranks).
e2e    : the same metric through the public API `train_sampled_models` with HOST caches: model
         construction, H2D of cache / weights / batch orders and the D2H of the accuracies are timed.
extras : (default run, outside the `value` timer) the other BASELINE.json configurations, sharded over the N ranks of the run
         (strong scaling: the job is fixed, candidate j trains on rank j % N): `search256` (north_star's 256-candidate search
         iteration), `search32` (configs[2]), `mmimdb64` (configs[3]), `depth` (configs[4]: 18 roofline points, 148 candidates
         per point, points sharded over the ranks) -- each with `value` (device-resident), `e2e` (public API, host buffers) and the
         per-GPU fraction of the HBM roofline.  `--workload X` makes X the main line instead.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONF4 = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]]       # main_found_ntu.py:181-182
H, B, C = 128, 64, 60
N_TRAIN, N_DEV = 10240, 5120
METRIC, UNIT = "candidate-epochs/sec (NTU fusion, bs=64)", "candidate-epochs/s"


def workload_of(a, n_gpus):
    """BASELINE.json configs: 'cfg2' = configs[1] (the metric's configuration, the default and the only one the driver
    runs); 'search32' = configs[2] (the 32 one-step configurations [i,j,a], H=16, 5 epochs, sharded over the GPUs);
    'search256' = north_star's 256-candidate search iteration (8 parents x 32 unfolded rows, L=2, H=16, 1 epoch)."""
    import numpy as np
    rows32 = [[i, j, k] for i in range(4) for j in range(4) for k in range(2)]
    if a.workload == "cfg2":
        return {"name": "cfg2", "H": H, "E": a.epochs, "per_gpu": a.candidates, "scaling": "weak",
                "confs": lambda n: [np.array(CONF4) for _ in range(n)],
                "desc": "NTU found conf=4 (4-step fusion), inner_repr=128, batchnorm, drpt=0, bs=64, "
                        f"N_train={N_TRAIN}, N_dev={N_DEV}, {a.candidates} candidates/GPU x {a.epochs} epoch(s) per step"}
    if a.workload == "search32":
        allc = [np.array([r]) for r in rows32]
        return {"name": "search32", "H": 16, "E": 5, "per_gpu": 32 // n_gpus, "scaling": "strong", "confs": lambda n: allc[:n],
                "desc": "main_searchable_ntu.py first SMBO level: the 32 one-step configurations [i,j,a], inner_repr=16, batchnorm, "
                        f"drpt=0, bs=64, N_train={N_TRAIN}, N_dev={N_DEV}, 5 epochs, {32 // n_gpus} candidates/GPU"}
    parents = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0], [2, 2, 0], [0, 0, 1], [3, 2, 0], [2, 3, 1]]
    allc = [np.array([p_, r]) for p_ in parents for r in rows32]
    return {"name": "search256", "H": 16, "E": 1, "per_gpu": 256 // n_gpus, "scaling": "strong", "confs": lambda n: allc[:n],
            "desc": "256-candidate search iteration: 8 parents x 32 unfolded rows, L=2, inner_repr=16, batchnorm, drpt=0, bs=64, "
                    f"N_train={N_TRAIN}, N_dev={N_DEV}, 1 epoch, {256 // n_gpus} candidates/GPU"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--candidates", type=int, default=148, help="candidates per GPU (weak scaling); 148 = one fused-chain CTA per SM")
    ap.add_argument("--epochs", type=int, default=3, help="epochs per candidate per step (search driver default: --epochs 3)")
    ap.add_argument("--cpu-sample-steps", type=int, default=3200, help="train steps of the CPU baseline sample (3200 = 20 candidate-epochs of one candidate: 10-15 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "search32", "search256", "mmimdb64", "depth"],
                    help="cfg2 = BASELINE.json configs[1] (default, the metric's configuration); the others are reported extras")
    ap.add_argument("--no-extras", action="store_true", help="skip the search256 / search32 / mmimdb64 / depth extras of the default run")
    return ap.parse_args()


def workload_config(a, n_gpus):
    w = workload_of(a, n_gpus)
    return {"workload": w["desc"],
            "candidates_per_gpu": w["per_gpu"], "epochs_per_step": w["E"], "parallelism": f"candidates sharded x{n_gpus}",
            "l2": "inputs exceed L2 (cache 0.46 GB + per-candidate state 12.5 MB x candidates >> 126 MB)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_sample(train_steps, threads=None):
    """Time the PyTorch-CPU port of the reference path on a bounded sample of the cfg2 workload:
    `train_steps` optimiser steps + train_steps/2 eval steps of one candidate (the 2:1 ratio of a
    candidate-epoch), extrapolated to candidate-epochs/s."""
    import torch
    from mfas_b200.cache import synthetic_ntu_cache
    from oracle.torch_port import FusionHeadTorch, train_candidate
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    per_epoch = math.ceil(N_TRAIN / B)                          # 160 train (+ 80 eval) steps = one candidate-epoch of cfg2
    epochs = max(1, train_steps // per_epoch)                   # whole epochs over the full-size split once the sample is that long
    ts = min(train_steps, per_epoch)
    dev_steps = max(1, ts // 2)
    train = synthetic_ntu_cache(max(ts * B, 2 * B), 1)
    dev = synthetic_ntu_cache(max(dev_steps * B, B), 2)
    torch.manual_seed(0)
    model = FusionHeadTorch(CONF4, H, C, batchnorm=True)
    tr, dv = (train.ske_cat, train.rgb_cat, train.labels), (dev.ske_cat, dev.rgb_cat, dev.labels)
    orders = lambda ph, e: torch.randperm(len(train) if ph == "train" else len(dev))
    train_candidate(model, tr, dv, orders, B, 1, max_train_steps=4, max_dev_steps=2)        # warm-up
    t0 = time.perf_counter()
    _, st = train_candidate(model, tr, dv, orders, B, epochs, max_train_steps=ts, max_dev_steps=dev_steps)
    dt = time.perf_counter() - t0
    n_tr, n_dv = sum(r["train_steps"] for r in st), sum(r["dev_steps"] for r in st)
    frac = n_tr / per_epoch                                     # candidate-epochs that were run
    return {"value": frac / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"oracle/torch_port.py (PyTorch CPU restatement of the reference path), 1 candidate, "
                      f"{n_tr} train + {n_dv} eval steps of cfg2 ({epochs} epoch(s)) in {dt:.2f} s, "
                      f"scaled to a {per_epoch}+{math.ceil(N_DEV / B)}-step candidate-epoch"}


def gpu_eager_reference_sample(device, train_steps=80):
    """The reference's arithmetic stack as eager PyTorch ON THE GPU (the favourable case: cached
    features already resident on the device, no DataLoader): the same port as the CPU baseline moved to `device`,
    one candidate, a bounded sample.  north_star's ">= 10x the reference 1-GPU PyTorch" is read against this."""
    import torch
    from mfas_b200.cache import synthetic_ntu_cache
    from oracle.torch_port import FusionHeadTorch, train_candidate
    dev_steps = max(1, train_steps // 2)
    train = synthetic_ntu_cache(train_steps * B, 1)
    dev = synthetic_ntu_cache(dev_steps * B, 2)
    torch.manual_seed(0)
    model = FusionHeadTorch(CONF4, H, C, batchnorm=True).to(device)
    tr = tuple(t.to(device) for t in (train.ske_cat, train.rgb_cat, train.labels))
    dv = tuple(t.to(device) for t in (dev.ske_cat, dev.rgb_cat, dev.labels))
    orders = lambda ph, e: torch.randperm(len(train) if ph == "train" else len(dev)).to(device)
    train_candidate(model, tr, dv, orders, B, 1, max_train_steps=10, max_dev_steps=5)        # warm-up (cuBLAS handles, allocator)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    _, st = train_candidate(model, tr, dv, orders, B, 1, max_train_steps=train_steps, max_dev_steps=dev_steps)
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    frac = st[0]["train_steps"] / math.ceil(N_TRAIN / B)
    return {"value": frac / dt, "unit": UNIT, "kind": "port",
            "sample": f"oracle/torch_port.py as eager PyTorch on cuda (fp32, TF32 off), 1 candidate at a time as the reference trains them, "
                      f"features resident on the device, {st[0]['train_steps']} train + {st[0]['dev_steps']} eval steps of cfg2 in {dt:.2f} s, "
                      f"scaled to a {math.ceil(N_TRAIN / B)}+{math.ceil(N_DEV / B)}-step candidate-epoch"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from mfas_b200.cache import synthetic_ntu_cache
    from oracle.torch_port import FusionHeadTorch, train_candidate
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    # each step: a bounded sample = one candidate-epoch (160 train + 80 eval steps) of one candidate, ~0.6 s on 16 cores
    ts, ds = math.ceil(N_TRAIN / B), math.ceil(N_DEV / B)
    train, dev = synthetic_ntu_cache(ts * B, 1), synthetic_ntu_cache(ds * B, 2)
    tr, dv = (train.ske_cat, train.rgb_cat, train.labels), (dev.ske_cat, dev.rgb_cat, dev.labels)
    orders = lambda ph, e: torch.randperm(len(train) if ph == "train" else len(dev))
    torch.manual_seed(0)
    model = FusionHeadTorch(CONF4, H, C, batchnorm=True)
    for _ in range(a.warmup):
        train_candidate(model, tr, dv, orders, B, 1)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        train_candidate(model, tr, dv, orders, B, 1)
    dt = (time.perf_counter() - t0) / max(a.steps, 1)
    value = (ts / math.ceil(N_TRAIN / B)) / dt
    sample = f"oracle/torch_port.py on {threads} host threads; each step = {ts} train + {ds} eval steps of one cfg2 candidate (one candidate-epoch)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a, a.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))



# ---------------------------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, sharded over the ranks of the run (strong scaling)
# ---------------------------------------------------------------------------------------------------------------------
class Dist:
    """rank / world plumbing shared by the measurements (one process per GPU; NCCL when world > 1)."""

    def __init__(self, rank, world, device):
        self.rank, self.world, self.device = rank, world, device

    def barrier(self):
        import torch
        import torch.distributed as td
        if self.world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        import torch
        import torch.distributed as td
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj):
        import torch.distributed as td
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        td.all_gather_object(out, obj)
        return out

    def mine(self, n):
        return [j for j in range(n) if j % self.world == self.rank]


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def _timed_runs(D, fn, steps, warmup):
    """`steps` calls of fn() after `warmup`, CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
    import torch
    for _ in range(warmup):
        fn()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    D.barrier()
    return D.max_over_ranks(e0.elapsed_time(e1) / 1e3) / max(steps, 1), out


def _timed_calls(D, fn, calls, warmup):
    """Host-timed public-API calls (wall clock, barrier + synchronize on both sides, max over ranks)."""
    for _ in range(warmup):
        fn()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(calls):
        out = fn()
    D.barrier()
    return D.max_over_ranks((time.perf_counter() - t0) / max(calls, 1)), out


def extra_search(name, a, D, host_train, host_dev, train_dev, dev_dev):
    """BASELINE configs[2] (`search32`) / north_star's 256-candidate iteration (`search256`): the job is fixed, its candidates
    are sharded over the ranks (candidate j -> rank j % N, as train_sampled_models shards them)."""
    import copy
    import numpy as np
    import torch
    from helpers import make_args
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200 import _lib
    from mfas_b200.cache import FeatureCacheLoader
    from mfas_b200.engine import CandidateGroup, algorithmic_counts
    a2 = copy.copy(a)
    a2.workload = name
    wl = workload_of(a2, 1)
    n_total, E, Hw = wl["per_gpu"], wl["E"], wl["H"]
    confs = wl["confs"](n_total)
    mine = D.mine(n_total)
    steps_tr, steps_dv = math.ceil(N_TRAIN / B), math.ceil(N_DEV / B)
    args = make_args(Hw, B, E, bn=True, drpt=0.0, Ti=1)
    args.init_on_device = True
    peak, _ = hbm_peak()
    cnts = [algorithmic_counts(l_, B) for l_ in CandidateGroup.plan(confs, Hw, C, _lib.FLAG_BN)]
    job_bytes = sum(E * (steps_tr * c_["train_bytes"] + steps_dv * c_["eval_bytes"]) for c_ in cnts)
    out = {"workload": wl["desc"].replace(f"{n_total} candidates/GPU", f"{n_total} candidates in all"), "scaling": "strong",
           "n_candidates": n_total, "n_candidates_per_gpu": len(mine) if D.world == 1 else [len([j for j in range(n_total) if j % D.world == r]) for r in range(D.world)],
           "epochs": E}
    # device-resident: this rank's share as one group
    g = None
    if mine:
        g = CandidateGroup([confs[j] for j in mine], Hw, C, _lib.FLAG_BN, D.device, batch_max=B, cand_ids=mine)
        g.set_adam(0.9, 0.999, 1e-8, 1e-4)
        ntu.init_on_device(g, 4242, mine)
        ltr, ldv = FeatureCacheLoader(host_train, B, True, 100), FeatureCacheLoader(host_dev, B, True, 200)
        ptr = ltr.orders_of([j * E + e for j in mine for e in range(E)], D.device).to(torch.int32).view(len(mine), E, N_TRAIN)
        pdv = ldv.orders_of([j * E + e for j in mine for e in range(E)], D.device).to(torch.int32).view(len(mine), E, N_DEV)
        lrs = ntu.cosine_lrs(args, N_TRAIN, E * steps_tr)
    fn = (lambda: g.train_run(train_dev, dev_dev, ptr, pdv, lrs, E, B)) if mine else (lambda: None)
    l0 = g.launches if g else 0
    dt, _ = _timed_runs(D, fn, 2, 2)
    out["gpu_launches_per_step"] = (g.launches - l0) // 4 if g else 0
    out["engine"] = g.engine if g else None
    if g:
        g.check()
        g.close()
    out["value"] = n_total * E / dt
    out["ms_per_step"] = dt * 1e3
    out["frac"] = job_bytes / dt / 1e9 / (D.world * peak)            # per-GPU fraction of the HBM roofline over the whole call
    # end to end through the public API: host caches in (H2D / broadcast inside the call), accuracies out
    loaders = {"train": FeatureCacheLoader(host_train, B, True, 100), "dev": FeatureCacheLoader(host_dev, B, True, 200)}

    def call():
        host_train.drop_device_copies(); host_dev.drop_device_copies()
        return ntu.train_sampled_models(confs, ntu.Searchable_Skeleton_Image_Net, loaders, args, D.device)

    dte, accs = _timed_calls(D, call, 2, 2)
    out["e2e"] = {"value": n_total * E / dte, "unit": UNIT, "ms_per_call": dte * 1e3, "frac": job_bytes / dte / 1e9 / (D.world * peak),
                  "h2d_bytes_per_step": int(host_train.nbytes() + host_dev.nbytes()), "d2h_bytes_per_step": int(len(mine) * (E * 4 + 1) * 8),
                  "init": "device generator keyed by (seed, candidate)", "accs_head": [float(x) for x in accs[:3]]}
    return out


def extra_mmimdb(a, D, n_total=64, E=3, Hm=256, Bm=64):
    """BASELINE configs[3]: MM-IMDB text+image searchable fusion, 64 two-step candidates x 3 epochs, inner_repr=256, bs=64,
    synthetic taps at the dataset's size (15552 train / 2608 dev rows), sharded over the ranks."""
    import numpy as np
    import torch
    from helpers import make_mmimdb_args
    import mfas_b200.mmimdb_searchable as mm
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200 import _lib
    from mfas_b200 import dist as mdist
    from mfas_b200.engine import CandidateGroup, algorithmic_counts
    n_tr, n_dv = 15552, 2608
    host_train, host_dev = mm.synthetic_mmimdb_cache(n_tr, 1).pin(), mm.synthetic_mmimdb_cache(n_dv, 2).pin()
    train_dev = mdist.broadcast_cache(host_train if D.rank == 0 else None, D.device)
    dev_dev = mdist.broadcast_cache(host_dev if D.rank == 0 else None, D.device)
    rows = mm.get_possible_layer_configurations(0)
    rng = np.random.default_rng(0)
    confs = [np.array([rows[i] for i in rng.integers(0, len(rows), size=2)]) for _ in range(n_total)]
    mine = D.mine(n_total)
    flags = _lib.FLAG_BN | _lib.FLAG_MULTILABEL
    steps_tr, steps_dv = math.ceil(n_tr / Bm), math.ceil(n_dv / Bm)
    peak, _ = hbm_peak()
    cnts = [algorithmic_counts(l_, Bm) for l_ in CandidateGroup.plan(confs, Hm, 23, flags, widths=mm.WIDTHS)]
    job_bytes = sum(E * (steps_tr * c_["train_bytes"] + steps_dv * c_["eval_bytes"]) for c_ in cnts)
    args = make_mmimdb_args(Hm, Bm, E, Ti=1)
    out = {"workload": f"MM-IMDB text+image searchable fusion (BASELINE configs[3]): {n_total} candidates x {E} epochs, inner_repr={Hm}, L=2, "
                       f"bs={Bm}, {n_tr}/{n_dv} rows", "scaling": "strong", "n_candidates": n_total, "epochs": E,
           "metric": "candidate-epochs/sec (MM-IMDB fusion, bs=64)"}
    g = None
    if mine:
        g = CandidateGroup([confs[j] for j in mine], Hm, 23, flags, D.device, batch_max=Bm, cand_ids=mine, widths=mm.WIDTHS)
        g.set_adam(0.9, 0.999, 1e-8, 1e-4)
        ntu.init_on_device(g, 4243, mine)
        ltr, ldv = mm.TextImageCacheLoader(host_train, Bm, True, 100), mm.TextImageCacheLoader(host_dev, Bm, True, 200)
        ptr = ltr.orders_of([j * E + e for j in mine for e in range(E)], D.device).to(torch.int32).view(len(mine), E, n_tr)
        pdv = ldv.orders_of([j * E + e for j in mine for e in range(E)], D.device).to(torch.int32).view(len(mine), E, n_dv)
        lrs = ntu.cosine_lrs(args, n_tr, E * steps_tr)
    fn = (lambda: g.train_run(train_dev, dev_dev, ptr, pdv, lrs, E, Bm)) if mine else (lambda: None)
    dt, _ = _timed_runs(D, fn, 2, 2)
    out["engine"] = g.engine if g else None
    if g:
        g.check()
        g.close()
    out["value"] = n_total * E / dt
    out["ms_per_step"] = dt * 1e3
    out["frac"] = job_bytes / dt / 1e9 / (D.world * peak)
    loaders = {"train": mm.TextImageCacheLoader(host_train, Bm, True, 100), "dev": mm.TextImageCacheLoader(host_dev, Bm, True, 200)}

    def call():
        host_train.drop_device_copies(); host_dev.drop_device_copies()
        return mm.train_sampled_models(confs, mm.Searchable_Text_Image_Net, loaders, args, D.device)

    dte, f1s = _timed_calls(D, call, 2, 2)
    n_params = sum(int(l_.n_params) for l_ in CandidateGroup.plan([confs[j] for j in mine], Hm, 23, flags, widths=mm.WIDTHS))
    out["e2e"] = {"value": n_total * E / dte, "unit": UNIT, "ms_per_call": dte * 1e3, "frac": job_bytes / dte / 1e9 / (D.world * peak),
                  "h2d_bytes_per_step": int(host_train.nbytes() + host_dev.nbytes() + 4 * n_params), "d2h_bytes_per_step": int(len(mine) * (E * 4 + 1) * 8),
                  "init": "host, the constructor's CPU RNG stream", "dev_f1_head": [float(x) for x in f1s[:3]]}
    return out


def extra_depth(a, D, train_dev, cands=148, Bd=128, steps=20):
    """BASELINE configs[4]: NTU fusion-depth sweep L in 1..6 x inner_repr in {64, 128, 256}, bs=128 -- the roofline fraction of
    the fused train step at every point (148 candidates per GPU, every candidate its own batch).  The 18 points are sharded
    over the ranks; every rank measures its points on its own GPU."""
    import numpy as np
    import torch
    from mfas_b200 import _lib
    from mfas_b200.engine import CandidateGroup, algorithmic_counts, plan_layout
    peak, _ = hbm_peak()
    points = [(H_, L_) for H_ in (64, 128, 256) for L_ in range(1, 7)]
    rows_out = []
    n_rows = len(train_dev)
    for idx in D.mine(len(points)):
        H_, L_ = points[idx]
        conf = np.array([CONF4[l % 4] for l in range(L_)])
        cnt = algorithmic_counts(plan_layout(conf, H_, C, _lib.FLAG_BN), Bd)
        g = CandidateGroup([conf] * cands, H_, C, _lib.FLAG_BN, D.device, batch_max=Bd)
        g.set_adam(0.9, 0.999, 1e-8, 1e-4)
        g.params.uniform_(-0.03, 0.03)
        g.bufs.fill_(1.0)
        gen = torch.Generator(device=D.device).manual_seed(idx)
        rws = [torch.randint(0, n_rows, (cands, Bd), device=D.device, generator=gen, dtype=torch.int32) for _ in range(steps + 5)]
        for i in range(5):
            g.train_step(train_dev, rws[i], lr=1e-3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            g.train_step(train_dev, rws[5 + i], lr=1e-3)
        e1.record()
        torch.cuda.synchronize()
        g.check()
        ms = e0.elapsed_time(e1) / steps
        gbs = cands * cnt["train_bytes"] / ms / 1e6
        rows_out.append({"L": L_, "inner_repr": H_, "batch": Bd, "candidates": cands, "engine": g.engine, "train_MB_per_candidate_step": cnt["train_bytes"] / 1e6,
                         "flop_per_byte": (cnt["fwd_flops"] + cnt["bwd_flops"]) / cnt["train_bytes"], "ms_per_step": ms,
                         "achieved_GBs": gbs, "frac": gbs / peak, "rank": D.rank})
        g.close()
    allrows = sorted([r for part in D.gather(rows_out) for r in part], key=lambda r: (r["inner_repr"], r["L"]))
    return {"workload": "NTU fusion-depth sweep (BASELINE configs[4]): L in 1..6 x inner_repr in {64,128,256}, rows conf4[l mod 4], bs=128, "
                        f"{cands} candidates per point, fused train step", "peak_GBs": peak, "points": allrows,
            "frac_min": min(r["frac"] for r in allrows), "frac_max": max(r["frac"] for r in allrows),
            "frac_mean": sum(r["frac"] for r in allrows) / len(allrows)}


def run_other_main(a, D, saved_stdout):
    """--workload mmimdb64 | depth as the main line (reported extras of BASELINE configs[3] / configs[4])."""
    import torch
    import torch.distributed as td
    from mfas_b200 import dist as mdist
    from mfas_b200.cache import synthetic_ntu_cache
    clocks = ClockSampler(D.device.index or 0)
    if D.rank == 0:
        clocks.start()
    if a.workload == "mmimdb64":
        r = extra_mmimdb(a, D)
        line = {"metric": r["metric"], "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "scaling": "strong",
                "e2e": r["e2e"], "roofline": {"bound": "hbm", "achieved": r["frac"] * hbm_peak()[0], "peak": hbm_peak()[0], "unit": "GB/s",
                                              "frac": r["frac"], "traffic": None, "kernel": "whole call (train + eval steps), per GPU"},
                "config": {"workload": r["workload"], "l2": "inputs exceed L2"}, "engine": r["engine"]}
    else:
        host = synthetic_ntu_cache(4096, 1).pin() if D.rank == 0 else None
        r = extra_depth(a, D, mdist.broadcast_cache(host, D.device))
        line = {"metric": "fused-step HBM GB/s (NTU fusion-depth sweep, bs=128)", "value": sum(p_["achieved_GBs"] for p_ in r["points"]) / len(r["points"]),
                "unit": "GB/s (mean over the 18 points)", "ms_per_step": sum(p_["ms_per_step"] for p_ in r["points"]), "scaling": "weak", "e2e": None,
                "roofline": {"bound": "hbm", "achieved": r["frac_mean"] * r["peak_GBs"], "peak": r["peak_GBs"], "unit": "GB/s", "frac": r["frac_mean"],
                             "traffic": None, "kernel": "fused train step, mean over the sweep"},
                "config": {"workload": r["workload"], "l2": "inputs exceed L2"}, "points": r["points"]}
    clk = clocks.stop() if D.rank == 0 else None
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if D.rank == 0:
        line.update({"n_gpus": D.world, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "f32",
                     "data": "synthetic", "clocks": clk, "gpu_launches": None})
        print(json.dumps(line))
    if D.world > 1:
        td.destroy_process_group()


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as td
    from helpers import make_args
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200 import dist as mdist
    from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache
    from mfas_b200.engine import CandidateGroup, algorithmic_counts
    from mfas_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective; the contract is ONE JSON line there
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        td.init_process_group("nccl", device_id=device)
    n_gpus = world
    D = Dist(rank, world, device)
    barrier, max_over_ranks = D.barrier, D.max_over_ranks

    if a.workload in ("mmimdb64", "depth"):
        run_other_main(a, D, saved_stdout)
        return

    wl = workload_of(a, n_gpus)
    M, E, Hw = wl["per_gpu"], wl["E"], wl["H"]
    # ---- inputs: rank 0 builds the cache, NCCL broadcasts it once -------------------------------
    host_train = synthetic_ntu_cache(N_TRAIN, 1).pin() if rank == 0 else None
    host_dev = synthetic_ntu_cache(N_DEV, 2).pin() if rank == 0 else None
    train_dev = mdist.broadcast_cache(host_train, device)
    dev_dev = mdist.broadcast_cache(host_dev, device)
    steps_tr, steps_dv = math.ceil(N_TRAIN / B), math.ceil(N_DEV / B)

    # ---- device-resident arm: `value` -------------------------------------------------------
    args = make_args(Hw, B, E, bn=True, drpt=0.0, Ti=1)
    all_confs = wl["confs"](M * n_gpus)
    confs = all_confs[rank * M:(rank + 1) * M]
    g = CandidateGroup(confs, Hw, C, _lib.FLAG_BN, device, batch_max=B, cand_ids=[rank * M + i for i in range(M)])
    g.set_adam(0.9, 0.999, 1e-8, 1e-4)
    torch.manual_seed(1234 + rank)
    for c in range(M):                                    # torch-default Linear init (kaiming-uniform bound 1/sqrt(K))
        for name in g.names(c):
            v = g.view(c, name)
            if name.endswith("0.weight") or name.endswith("0.bias") or name.startswith("central"):
                K = g.view(c, name.rsplit(".", 1)[0] + ".weight").shape[1]
                v.uniform_(-1 / math.sqrt(K), 1 / math.sqrt(K))
            elif name.endswith("2.weight") or name.endswith("running_var"):
                v.fill_(1.0)
    gen = torch.Generator().manual_seed(77 + rank)
    ptr = torch.stack([torch.stack([torch.randperm(N_TRAIN, generator=gen) for _ in range(E)]) for _ in range(M)]).to(device, torch.int32)
    pdv = torch.stack([torch.stack([torch.randperm(N_DEV, generator=gen) for _ in range(E)]) for _ in range(M)]).to(device, torch.int32)
    lrs = ntu.cosine_lrs(args, N_TRAIN, E * steps_tr)

    def one_step():
        return g.train_run(train_dev, dev_dev, ptr, pdv, lrs, E, B)

    for _ in range(a.warmup):
        one_step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = g.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        stats, best, _ = one_step()
    ev1.record()
    barrier()
    dt = max_over_ranks(ev0.elapsed_time(ev1) / 1e3) / max(a.steps, 1)
    launches = g.launches - l0
    clk = clocks.stop() if rank == 0 else None
    value = M * E * n_gpus / dt
    finite = bool(torch.isfinite(stats).all().item())

    # ---- roofline of the fused train step (all launches of one optimiser step, all candidates) ----
    cnts = [algorithmic_counts(l_, B) for l_ in g.layouts]
    cnt = {k: sum(c_[k] for c_ in cnts) / M for k in cnts[0]}          # per-candidate mean (the candidates of cfg2 are identical)
    engine = g.engine
    g_n_params = sum(int(l_.n_params) for l_ in g.layouts) / M
    rows = ptr[:, 0, :B].contiguous()
    for _ in range(5):
        g.train_step(train_dev, rows, 1e-4)
    n_t = 40
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    r0.record()
    for i in range(n_t):
        g.train_step(train_dev, ptr[:, 0, (i % steps_tr) * B:(i % steps_tr + 1) * B], 1e-4)
    r1.record()
    torch.cuda.synchronize()
    t_step = r0.elapsed_time(r1) / 1e3 / n_t
    # one dev pass (eval steps: forward stream + forward chain + head, running BatchNorm statistics)
    for _ in range(2):
        g.eval_pass(dev_dev, B, pdv[:, 0])
    torch.cuda.synchronize()
    r0.record()
    for _ in range(3):
        g.eval_pass(dev_dev, B, pdv[:, 0])
    r1.record()
    torch.cuda.synchronize()
    t_eval = r0.elapsed_time(r1) / 1e3 / 3 / steps_dv
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = M * cnt["train_bytes"] / t_step / 1e9
    # ncu --set full capture of the same step (dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r02_ncu_traffic.json: a committed capture, not measured in this run)
    traffic, per_kernel_traffic = None, {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        if tj.get("workload") == "cfg2" and wl["name"] == "cfg2":
            per_kernel_traffic = {k: v * M for k, v in tj["dram_bytes_per_launch_per_candidate"].items()}
            traffic = float(sum(per_kernel_traffic.values()))
    except Exception:
        pass
    # the three kernels of the step, each timed with CUDA events on the launching stream (C ABI: mfas_group_last_step_ms)
    kernels = []
    if engine == "tc" and wl["name"] == "cfg2":
        lay = g.layouts[0]
        L, Hh = lay.L, lay.H
        k_feat = sum(lay.d_ske[l] + lay.d_rgb[l] for l in range(L))
        k_all = sum(lay.K[l] for l in range(L))
        alg = {"k_tc_fwd_ws": 4 * (Hh * k_feat + B * k_feat),                                   # W feature columns + gathered x
               "k_chain_all": 4 * (2 * Hh * Hh * (L - 1) + 6 * (3 * Hh * L + C * Hh + C) + 4 * B * Hh * L),
               "k_tc_bwd_ws": 4 * (6 * Hh * k_all + B * k_all + B * Hh * L)}                    # p/m/v read+write, x, dz
        try:
            g.set_profiling(True)
            acc = [0.0, 0.0, 0.0]
            n_p = 20
            for i in range(n_p):
                g.train_step(train_dev, ptr[:, 0, (i % steps_tr) * B:(i % steps_tr + 1) * B], 1e-4)
                ms = g.last_step_ms()
                acc = [a_ + m_ for a_, m_ in zip(acc, ms)]
            g.set_profiling(False)
            for name, tms in zip(("k_tc_fwd_ws", "k_chain_all", "k_tc_bwd_ws"), acc):
                tms /= n_p
                gbs = M * alg[name] / (tms * 1e-3) / 1e9
                kernels.append({"kernel": name, "ms_per_launch": tms, "share_of_step": None, "algorithmic_bytes_per_launch": M * alg[name],
                                "achieved": gbs, "frac": gbs / peak,
                                "traffic": per_kernel_traffic.get(name)})
            tot = sum(k["ms_per_launch"] for k in kernels)
            for k in kernels:
                k["share_of_step"] = k["ms_per_launch"] / tot
        except Exception as ex:                      # profiling is an aid, never the measurement itself
            kernels = [{"error": str(ex)}]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/r02_ncu_traffic.json (ncu --set full capture r02fc of this command line's kernels, committed; not re-measured in this run)",
                "kernel": f"fused train step of {M} candidates (" + ("tc engine: k_tc_fwd_ws -> k_chain_all -> k_tc_bwd_ws, 3 launches; " if engine == "tc" else
                                                                       "ffma engine (inner_repr below the MMA tiles): per-layer CUDA-core kernels; ") +
                          "achieved = SURVEY 8(d) algorithmic bytes of the whole step / CUDA-event time of the whole step)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": M * cnt["train_bytes"], "ms_per_launch": t_step * 1e3,
                "eval_step": {"ms": t_eval * 1e3, "algorithmic_bytes": M * cnt["eval_bytes"], "achieved": M * cnt["eval_bytes"] / t_eval / 1e9,
                              "frac": M * cnt["eval_bytes"] / t_eval / 1e9 / peak},
                "tensor_tflops": M * (cnt["fwd_flops"] + cnt["bwd_flops"]) / t_step / 1e12,
                "kernels": kernels}
    g.close()

    # ---- e2e: the public API with HOST buffers ---------------------------------------------------
    e2e = None
    e2e_host = None
    if rank != 0 and (not a.no_e2e or not a.no_extras):       # every rank owns a host copy, as a real multi-process search would
        host_train = synthetic_ntu_cache(N_TRAIN, 1).pin()
        host_dev = synthetic_ntu_cache(N_DEV, 2).pin()
    if not a.no_e2e:
        loaders = {"train": FeatureCacheLoader(host_train, B, True, 100), "dev": FeatureCacheLoader(host_dev, B, True, 200)}
        n_params = int(g_n_params)

        def measure(init_on_device):
            import copy
            a2 = copy.copy(args)
            if init_on_device is not None:
                a2.init_on_device = init_on_device

            def e2e_step():
                host_train.drop_device_copies(); host_dev.drop_device_copies()   # H2D of the cache is inside the step
                return ntu.train_sampled_models(all_confs, ntu.Searchable_Skeleton_Image_Net, loaders, a2, device)

            for _ in range(3):              # warm-up: pins the staging arenas, fills the allocator caches, NCCL lazy init
                e2e_step()
            barrier()
            n_it = max(1, min(a.steps, 3))
            t0 = time.perf_counter()
            calls = []
            for _ in range(n_it):
                tc0 = time.perf_counter()
                accs = e2e_step()
                calls.append((time.perf_counter() - tc0) * 1e3)
            barrier()
            dte = max_over_ranks((time.perf_counter() - t0) / n_it)
            on_dev = bool(init_on_device)
            h2d = host_train.nbytes() + host_dev.nbytes() + (0 if on_dev else M * n_params * 4)
            return {"value": M * E * n_gpus / dte, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(M * (E * 4 + 1) * 8), "ms_per_step": dte * 1e3, "ms_per_call": calls,
                    "init": "device generator keyed by (seed, candidate)" if on_dev else
                            "host, bit-compatible with the reference constructor's CPU RNG stream",
                    "api": "mfas_b200.ntu_searchable.train_sampled_models(confs, Searchable_Skeleton_Image_Net, loaders, args, device): "
                           "host FeatureCache in pinned memory; parameter initialisation, cache / weight uploads, batch-order "
                           "generation and the accuracy read-back are all inside the timed region",
                    "accs_head": [float(x) for x in accs[:3]]}

        # the SAME two modes at every N: `e2e` = args.init_on_device=True (initial weights from a device generator keyed by
        # (seed, candidate): the work per rank does not depend on N); `e2e_host_init` = the reference-compatible constructor
        # stream (every rank draws the stream of the WHOLE call and keeps its share: that cost grows with N)
        e2e = measure(True)
        e2e_host = measure(False)

    extras = None
    if wl["name"] == "cfg2" and not a.no_extras:
        extras = {}
        for nm in ("search256", "search32"):
            try:
                extras[nm] = extra_search(nm, a, D, host_train, host_dev, train_dev, dev_dev)
            except Exception as ex:                      # an extra never takes the main line down
                extras[nm] = {"error": repr(ex)}
        for nm, fn in (("mmimdb64", lambda: extra_mmimdb(a, D)), ("depth", lambda: extra_depth(a, D, train_dev))):
            try:
                extras[nm] = fn()
            except Exception as ex:
                extras[nm] = {"error": repr(ex)}

    cpu = None
    gpu_eager = None
    if rank == 0 and n_gpus == 1 and not a.no_cpu_baseline and wl["name"] == "cfg2":
        cpu = cpu_reference_sample(a.cpu_sample_steps)
        try:
            gpu_eager = gpu_eager_reference_sample(device)
        except Exception as ex:                      # a reported baseline, never the measurement itself
            gpu_eager = {"error": str(ex)}

    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a, n_gpus), "clocks": clk, "e2e": e2e, "e2e_host_init": e2e_host, "gpu_launches": int(launches),
            "extras": extras,
            "roofline": roofline, "cpu_baseline": cpu, "reference_gpu_eager": gpu_eager, "finite": finite,
            "hbm_ceiling_cand_epochs_per_s_per_gpu": peak * 1e9 / (steps_tr * cnt["train_bytes"] + steps_dv * cnt["eval_bytes"])}))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
