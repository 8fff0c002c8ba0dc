#!/usr/bin/env python3
"""Target of the compute-sanitizer runs (profiles/run_gpu_r02a.sh): one small train_run + eval pass + step of every kernel family.

    compute-sanitizer --tool memcheck  python profiles/sanitize_target.py
    compute-sanitizer --tool racecheck python profiles/sanitize_target.py

Shapes are tiny (the tools serialise and instrument every launch) but cover every kernel: the tcgen05 engine at inner_repr 128
(persistent forward / fused chain + tensor-core head / persistent backward + Adam), at inner_repr 16 (masked tiles), at 128-row
batches and inner_repr 256, the CUDA-core engine, the alpha gates, the multi-label head and the pooling kernel.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import FOUND_CONFS, init_states  # noqa: E402
from mfas_b200 import _lib  # noqa: E402
from mfas_b200.cache import synthetic_ntu_cache  # noqa: E402
from mfas_b200.engine import CandidateGroup  # noqa: E402

DEV = "cuda:0"


def run(confs, H, B, flags=_lib.FLAG_BN, engine=None, E=1, ntr=None, ndv=None):
    ntr, ndv = ntr or 2 * B + 3, ndv or B + 5
    train, dev = synthetic_ntu_cache(ntr, 1).to(DEV), synthetic_ntu_cache(ndv, 2).to(DEV)
    if engine:
        os.environ["MFAS_ENGINE"] = engine
    g = CandidateGroup(confs, H, 60, flags, DEV, batch_max=B)
    os.environ.pop("MFAS_ENGINE", None)
    g.set_adam(0.9, 0.999, 1e-8, 1e-4)
    for k, st in enumerate(init_states(confs, H, 60, True, 0.0, 1)):
        g.load_state(k, st)
    gen = torch.Generator().manual_seed(0)
    ptr = torch.stack([torch.stack([torch.randperm(ntr, generator=gen) for _ in range(E)]) for _ in confs])
    pdv = torch.stack([torch.stack([torch.randperm(ndv, generator=gen) for _ in range(E)]) for _ in confs])
    steps = -(-ntr // B)
    stats, best, _ = g.train_run(train, dev, ptr, pdv, [1e-3] * (E * steps), E, B)
    g.eval_pass(dev, B, pdv[:, 0])
    g.train_step(train, ptr[:, 0, :B], 1e-3)
    g.check()
    print(f"ok engine={g.engine} H={H} B={B} n={len(confs)} flags={flags} best={best.cpu().tolist()}", flush=True)
    g.close()


def main():
    assert torch.cuda.is_available()
    run([FOUND_CONFS[4], FOUND_CONFS[1][:2]], 128, 64)
    run([[[3, 1, 1], [1, 3, 0]], [[0, 0, 1]]], 16, 64)
    run([[[1, 3, 0], [3, 0, 1]]], 256, 128)
    run([FOUND_CONFS[4][:2]], 128, 128)
    run([FOUND_CONFS[0][:2]], 48, 16, engine="ffma")
    run([[[3, 1, 1], [1, 3, 0]]], 64, 32, flags=_lib.FLAG_BN | _lib.FLAG_ALPHAS)
    run([[[3, 1, 1], [1, 3, 0]], [[2, 2, 0]]], 16, 32, flags=_lib.FLAG_BN | _lib.FLAG_ALPHAS)      # the small-inner_repr kernels with the gates
    run([[[0, 1, 1], [1, 0, 0]]], 32, 128)                                                             # ... at 128-row batches
    # multi-label head (MM-IMDB tap set) and the pooling kernel
    import mfas_b200.mmimdb_searchable as mm
    tr, dv = mm.synthetic_mmimdb_cache(70, 1).to(DEV), mm.synthetic_mmimdb_cache(40, 2).to(DEV)
    g = CandidateGroup([np.array([[1, 2, 0], [0, 3, 1]])], 64, 23, _lib.FLAG_BN | _lib.FLAG_MULTILABEL, DEV, batch_max=32, widths=mm.WIDTHS)
    g.set_adam(0.9, 0.999, 1e-8, 1e-4)
    g.params.uniform_(-0.05, 0.05)
    g.bufs.fill_(1.0)
    gen = torch.Generator().manual_seed(0)
    ptr, pdv = torch.randperm(70, generator=gen)[None, None], torch.randperm(40, generator=gen)[None, None]
    g.train_run(tr, dv, ptr, pdv, [1e-3] * 3, 1, 32)
    g.check()
    print(f"ok multilabel engine={g.engine}", flush=True)
    g.close()
    from mfas_b200.cache_builder import global_pool_into
    x = torch.rand(4, 64, 3, 7, 7, device=DEV)
    out = torch.zeros(4, 64, device=DEV)
    global_pool_into(x, out)
    torch.cuda.synchronize()
    assert torch.allclose(out, x.flatten(2).mean(2), atol=1e-6)
    print("ok pooling", flush=True)
    _lib.lib().mfas_release_cached_memory()


if __name__ == "__main__":
    main()
