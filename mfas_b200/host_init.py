"""Reference-compatible parameter initialisation at memory speed.

``searchable_type(args, conf)`` in the reference (models/search/ntu_searchable.py:44, :200, :274-282) draws every
Linear weight and bias from torch's global CPU generator, one mt19937 word per float, serially.  The C helper
``mfas_host_uniform_fill`` (csrc/host_init.cpp) produces the same stream in bulk straight into the pinned arenas;
this module feeds it the layout, keeps torch's generator state in sync, and cross-checks the helper against torch
itself once per process (falling back to torch's own initialisers if the two ever disagree).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib

_mode = None          # None: unchecked; -1: helper disabled (mismatch); 0 / 1: use_fma flag that matches torch


def _fill(state: torch.Tensor, dsts, counts, lo, hi, use_fma: int) -> bool:
    n = len(dsts)
    d = (C.c_void_p * n)(*dsts)
    c = np.asarray(counts, dtype=np.int64)
    a = np.asarray(lo, dtype=np.float32)
    b = np.asarray(hi, dtype=np.float32)
    rc = _lib.lib().mfas_host_uniform_fill(state.data_ptr(), state.numel(), n, C.cast(d, C.c_void_p),
                                           c.ctypes.data, a.ctypes.data, b.ctypes.data, use_fma)
    return rc == 0


def _self_check() -> int:
    """Which arithmetic variant of the helper reproduces torch's uniform_ bit for bit on this machine (-1: none)."""
    saved = torch.get_rng_state()
    try:
        ref_state = saved.clone()
        torch.set_rng_state(ref_state)
        bound = 1.0 / math.sqrt(2432.0)
        want = torch.empty(3001, dtype=torch.float32).uniform_(-bound, bound)
        want2 = torch.empty(777, dtype=torch.float32).uniform_(-0.3, 0.3)
        after = torch.get_rng_state()
        for fma in (0, 1):
            st = saved.clone()
            got, got2 = torch.empty(3001, dtype=torch.float32), torch.empty(777, dtype=torch.float32)
            ok = _fill(st, [got.data_ptr(), got2.data_ptr()], [3001, 777], [-bound, -0.3], [bound, 0.3], fma)
            if ok and torch.equal(got, want) and torch.equal(got2, want2) and torch.equal(st, after):
                return fma
        return -1
    finally:
        torch.set_rng_state(saved)


def kaiming_bound(fan_in: int) -> float:
    """nn.init.kaiming_uniform_(w, a=sqrt(5)) as nn.Linear.reset_parameters calls it."""
    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
    return math.sqrt(3.0) * (gain / math.sqrt(fan_in))


def init_host_arenas_fast(group, host_p, host_b, visit, slots=None) -> bool:
    """Same result as ntu_searchable.init_host_arenas (torch initialisers in constructor order), through the bulk
    helper.  ``visit(group, slot, fill)`` enumerates the tensors.  Returns False (nothing touched) if the helper is
    not usable here."""
    global _mode
    if _mode is None:
        _mode = _self_check()
    if _mode < 0:
        return False
    state = torch.get_rng_state()
    pend = ([], [], [], [])

    def flush():
        if pend[0]:
            if not _fill(state, *pend, _mode):
                raise RuntimeError("mfas_host_uniform_fill rejected the generator state: " + _lib.lib().mfas_last_error().decode())
            for p in pend:
                p.clear()

    base_p, base_b = host_p.data_ptr(), host_b.data_ptr()
    for c in (range(group.n) if slots is None else slots):
        def fill(name, kind, fan_in, c=c):
            nonlocal state
            arena, off, shape = group.slots[c][name]
            o = int(group.b_off[c] if arena == "b" else group.p_off[c]) + int(off)
            n = int(np.prod(shape)) if shape else 1
            if kind in ("kaiming", "uniform"):
                bound = kaiming_bound(fan_in) if kind == "kaiming" else (1 / math.sqrt(fan_in) if fan_in > 0 else 0)
                pend[0].append((base_b if arena == "b" else base_p) + 4 * o)
                pend[1].append(n); pend[2].append(-bound); pend[3].append(bound)
                return
            t = (host_b if arena == "b" else host_p)[o:o + n]
            if kind == "ones":
                t.fill_(1.0)
            elif kind == "zeros":
                t.zero_()
            elif kind == "normal":            # N(0, 0.1) of the alphas: Box-Muller with a cached sample -- left to torch
                flush()
                torch.set_rng_state(state)
                nn.init.normal_(t.view(shape), 0.0, 0.1)
                state = torch.get_rng_state()
        visit(group, c, fill)
    flush()
    torch.set_rng_state(state)
    return True
