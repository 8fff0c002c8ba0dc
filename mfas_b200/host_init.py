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
import struct

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
                return fma if _normal_path_matches(saved, fma) else -1
        return -1
    finally:
        torch.set_rng_state(saved)


_OFF_CACHED, _OFF_VALID, _STATE_BYTES = 5024, 5040, 5056      # at::mt19937 state as torch serialises it (cached normal sample, its flag)


def _box_muller(raw4):
    """at::normal_distribution<double> on four generator words: (cos branch = the sample, sin branch = cached for the next draw)."""
    r1, r2, r3, r4 = (int(x) for x in raw4)
    u1 = (((r1 << 32) | r2) & ((1 << 53) - 1)) * (1.0 / (1 << 53))
    u2 = (((r3 << 32) | r4) & ((1 << 53) - 1)) * (1.0 / (1 << 53))
    r = math.sqrt(-2.0 * math.log1p(-u2))
    theta = 2.0 * math.pi * u1
    return r * math.cos(theta), r * math.sin(theta)


def _normal_path_matches(saved: torch.Tensor, fma: int) -> bool:
    """The scalar-normal emulation (raw words + Box-Muller + the cached-sample fields patched at fixed byte offsets of the
    serialised generator state) against torch itself: 3 then 2 scalar ``normal_`` draws -- odd and even counts, so the
    cache flag is exercised both ways -- must give the same values AND the same final generator state.  A torch build
    with another state layout fails here and the whole helper is disabled (torch's own initialisers are used instead)."""
    if saved.numel() != _STATE_BYTES:
        return False
    for n_draws in (3, 2):
        torch.set_rng_state(saved.clone())
        want = [float(torch.zeros(1, dtype=torch.float32).normal_(0.0, 0.1)) for _ in range(n_draws)]
        after = torch.get_rng_state()
        st = saved.clone()
        sbuf = st.numpy()
        valid = bool(struct.unpack_from("i", sbuf, _OFF_VALID)[0])
        cached = struct.unpack_from("d", sbuf, _OFF_CACHED)[0]
        got = []
        for _ in range(n_draws):
            if valid:
                z, valid = cached, False
            else:
                raw = np.zeros(4, dtype=np.uint32)
                if not _fill(st, [raw.ctypes.data], [-4], [0.0], [0.0], fma):
                    return False
                z, cached = _box_muller(raw)
                valid = True
            got.append(float(torch.zeros(1, dtype=torch.float32).fill_(z * 0.1 + 0.0)))
        struct.pack_into("d", sbuf, _OFF_CACHED, cached if valid else 0.0)
        struct.pack_into("i", sbuf, _OFF_VALID, 1 if valid else 0)
        if got != want or not torch.equal(st, after):
            return False
    return True


def kaiming_bound(fan_in: int) -> float:
    """nn.init.kaiming_uniform_(w, a=sqrt(5)) as nn.Linear.reset_parameters calls it."""
    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
    return math.sqrt(3.0) * (gain / math.sqrt(fan_in))


_TEMPLATES = {}


def _template(group, c, visit):
    """The tensors of candidate ``c`` in constructor order as (kind, in buffer arena, offset, numel, bound, shape), cached
    per layout: a search iteration initialises hundreds of candidates that share a handful of layouts."""
    key = bytes(group.layouts[c])
    tpl = _TEMPLATES.get(key)
    if tpl is None:
        tpl = []

        def rec(name, kind, fan_in):
            arena, off, shape = group.slots[c][name]
            n = int(np.prod(shape)) if shape else 1
            if kind in ("kaiming", "uniform"):
                bound = kaiming_bound(fan_in) if kind == "kaiming" else (1 / math.sqrt(fan_in) if fan_in > 0 else 0)
                tpl.append((0, arena == "b", int(off), n, bound, shape))
            else:
                tpl.append(({"ones": 1, "zeros": 2, "normal": 3}[kind], arena == "b", int(off), n, 0.0, shape))
        visit(group, c, rec)
        if len(_TEMPLATES) > 4096:
            _TEMPLATES.clear()
        _TEMPLATES[key] = tpl
    return tpl


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
    slot_list = list(range(group.n) if slots is None else slots)
    # N(0, 0.1) of the scalar alphas (at::normal_distribution<double> through cpu_serial_kernel): Box-Muller on two
    # 53-bit uniforms = 4 generator words, the sine branch cached in the generator for the next draw.  The words are
    # fetched raw inside the same bulk fill and turned into samples afterwards, so one C call (and its thread pipeline)
    # covers every candidate of the slice.
    sbuf = state.numpy()
    cache_valid = bool(struct.unpack_from("i", sbuf, _OFF_VALID)[0])
    cached = struct.unpack_from("d", sbuf, _OFF_CACHED)[0]
    raw = np.zeros(4 * max(1, sum(int(group.layouts[c].L) for c in slot_list)), dtype=np.uint32)
    n_raw = 0
    normals = []                      # (tensor view, offset into raw or None when the cached sample serves it)

    def flush():
        if pend[0]:
            if not _fill(state, *pend, _mode):
                raise RuntimeError("mfas_host_uniform_fill rejected the generator state: " + _lib.lib().mfas_last_error().decode())
            for p in pend:
                p.clear()

    def finish_normals():
        nonlocal cached
        for t, off in normals:
            if off is None:
                z = cached
            else:
                z, cached = _box_muller(raw[off:off + 4])
            t.fill_(z * 0.1 + 0.0)
        normals.clear()

    base_p, base_b = host_p.data_ptr(), host_b.data_ptr()
    np_p, np_b = host_p.numpy(), host_b.numpy()        # views of the same (pinned) memory
    p_off, b_off = group.p_off.tolist(), group.b_off.tolist()
    for c in slot_list:
        pb, bb = int(p_off[c]), int(b_off[c])
        for kind, is_b, off, n, bound, shape in _template(group, c, visit):
            o = (bb if is_b else pb) + off
            if kind == 0:                                  # kaiming / uniform
                pend[0].append((base_b if is_b else base_p) + 4 * o)
                pend[1].append(n); pend[2].append(-bound); pend[3].append(bound)
            elif kind == 1:
                (np_b if is_b else np_p)[o:o + n] = 1.0
            elif kind == 2:
                (np_b if is_b else np_p)[o:o + n] = 0.0
            elif n == 1:                                   # scalar normal
                t = (host_b if is_b else host_p)[o:o + 1]
                if cache_valid:
                    normals.append((t, None))
                    cache_valid = False
                else:
                    pend[0].append(raw.ctypes.data + 4 * n_raw)
                    pend[1].append(-4); pend[2].append(0.0); pend[3].append(0.0)
                    normals.append((t, n_raw))
                    n_raw += 4
                    cache_valid = True
            else:                                          # larger tensors take torch's vectorised normal_fill: left to torch
                t = (host_b if is_b else host_p)[o:o + n]
                flush()
                finish_normals()
                struct.pack_into("d", sbuf, _OFF_CACHED, cached if cache_valid else 0.0)
                struct.pack_into("i", sbuf, _OFF_VALID, 1 if cache_valid else 0)
                torch.set_rng_state(state)
                nn.init.normal_(t.view(shape), 0.0, 0.1)
                state = torch.get_rng_state()
                sbuf = state.numpy()
                cache_valid = bool(struct.unpack_from("i", sbuf, _OFF_VALID)[0])
                cached = struct.unpack_from("d", sbuf, _OFF_CACHED)[0]
    flush()
    finish_normals()
    struct.pack_into("d", sbuf, _OFF_CACHED, cached if cache_valid else 0.0)
    struct.pack_into("i", sbuf, _OFF_VALID, 1 if cache_valid else 0)
    torch.set_rng_state(state)
    return True
