#!/usr/bin/env python3
"""Host-side phase timings (MFAS_TIMING=1, synchronising) of one train_sampled_models call for the search workloads."""
import os, sys, time, copy
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_args
import mfas_b200.ntu_searchable as ntu
from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache
dev = torch.device("cuda:0")
rows32 = [[i, j, k] for i in range(4) for j in range(4) for k in range(2)]
parents = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0], [2, 2, 0], [0, 0, 1], [3, 2, 0], [2, 3, 1]]
host_train, host_dev = synthetic_ntu_cache(10240, 1).pin(), synthetic_ntu_cache(5120, 2).pin()
for name, confs, E in (("search256", [np.array([p, r]) for p in parents for r in rows32], 1), ("search32x1/8", [np.array([p, r]) for p in parents for r in rows32][:32], 1),
                       ("search32", [np.array([r]) for r in rows32], 5)):
    for on_dev in (True, False):
        args = make_args(16, 64, E, bn=True, drpt=0.0, Ti=1); args.init_on_device = on_dev
        loaders = {"train": FeatureCacheLoader(host_train, 64, True, 100), "dev": FeatureCacheLoader(host_dev, 64, True, 200)}
        def call():
            host_train.drop_device_copies(); host_dev.drop_device_copies()
            return ntu.train_sampled_models(confs, ntu.Searchable_Skeleton_Image_Net, loaders, args, dev)
        os.environ.pop("MFAS_TIMING", None)
        for _ in range(3): call()
        torch.cuda.synchronize(); t0 = time.perf_counter(); call(); torch.cuda.synchronize()
        print(f"=== {name} init_on_device={on_dev}: {1e3 * (time.perf_counter() - t0):.1f} ms per call; phases:", flush=True)
        os.environ["MFAS_TIMING"] = "1"
        call()
        sys.stderr.flush()
