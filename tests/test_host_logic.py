"""Host-side logic that needs no GPU: ABI surface, layouts, init parity, scheduler, sharding."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import FOUND_CONFS, ROOT, init_states, make_args


def test_library_exports_every_declared_symbol():
    """The C-ABI shared library loads without a GPU and exports exactly what include/mfas_b200.h declares."""
    from mfas_b200 import _lib
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "mfas_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)          # prose in comments is not a declaration
    declared = set(re.findall(r"\b(mfas_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().mfas_abi_version() == _lib.ABI_VERSION
    assert re.search(r"#define MFAS_ABI_VERSION (\d+)", header).group(1) == str(_lib.ABI_VERSION)
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_tensor_core_sass_present():
    """The shipped .so carries tcgen05 code: UTC*MMA (tcgen05.mma) and LDTM (tcgen05.ld) in SASS."""
    from mfas_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass
    assert "LDTM" in sass


def test_layout_matches_reference_shapes():
    from mfas_b200 import _lib
    from mfas_b200.engine import algorithmic_counts, plan_layout, tensor_slots
    lay = plan_layout(FOUND_CONFS[4], 128, 60, _lib.FLAG_BN)
    assert list(lay.K)[:4] == [1536, 2432, 1408, 2688]
    slots = tensor_slots(lay)
    n = sum(int(np.prod(shape)) for (arena, off, shape) in slots.values() if arena == "p")
    assert n == 1041472                                                    # central parameter count incl. 4 alphas
    cnt = algorithmic_counts(lay, 64)
    assert abs(cnt["train_bytes"] - 26.96e6) < 0.01e6 and abs(cnt["eval_bytes"] - 6.13e6) < 0.01e6
    # offsets are 16-byte aligned and tensors do not overlap
    spans = sorted((off, off + int(np.prod(shape))) for (arena, off, shape) in slots.values() if arena == "p")
    assert all(a % 4 == 0 for a, _ in spans) and all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    with pytest.raises(_lib.MfasError):
        plan_layout([[0, 0, 0]], 16, 60, 0)          # drpt<1e-10 and no BN: the reference has no recipe
    with pytest.raises(_lib.MfasError):
        plan_layout([[_lib.NUM_TAPS, 0, 0]], 16, 60, _lib.FLAG_BN)      # beyond the ABI's tap slots: the library's range error
    with pytest.raises(ValueError):
        plan_layout([[4, 0, 0]], 16, 60, _lib.FLAG_BN)                  # a slot the 4-tap NTU set leaves unused
    with pytest.raises(_lib.MfasError):
        plan_layout([[0, 0, 0]], 16, 60, _lib.FLAG_BN | _lib.FLAG_PLAIN)      # the AV-MNIST recipe has no BatchNorm
    assert plan_layout([[0, 0, 0]], 16, 10, _lib.FLAG_PLAIN).off_gamma[0] == -1      # ... and allows "no BatchNorm, no Dropout"


def test_module_has_reference_state_dict_and_init():
    """Same keys, same shapes and -- for a given torch seed -- the same initial values as the reference ctor."""
    import mfas_b200.ntu_searchable as ntu
    args = make_args(128, 64, 1, bn=True)
    torch.manual_seed(0)
    m = ntu.Searchable_Skeleton_Image_Net(args, np.array(FOUND_CONFS[4]))
    ref = init_states([FOUND_CONFS[4]], 128, 60, True, 0.0, 0)[0]          # pinned to the reference by gen_golden.py
    sd = m.state_dict()
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert np.array_equal(sd[k].numpy(), v), k
    assert [l[0].in_features for l in m.fusion_layers] == [1536, 2432, 1408, 2688]
    assert len(m.central_params()) == 3
    assert ntu.get_possible_layer_configurations(0) == [[t, v, n] for t in range(4) for v in range(4) for n in range(2)]
    with pytest.raises(RuntimeError):
        m((torch.zeros(2, 5632), torch.zeros(2, 1920)))                    # CPU tensors: no fallback


def test_direct_arena_init_equals_module_construction():
    """The module-free initialisation used by train_sampled_models consumes the CPU generator exactly like the
    constructor: bit-identical weights for every candidate of a call."""
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200 import _lib
    from mfas_b200.engine import GroupLayout
    confs = [np.array(FOUND_CONFS[4]), np.array([[0, 0, 0]]), np.array(FOUND_CONFS[0][:3])]
    args = make_args(64, 64, 1, bn=True)
    g = GroupLayout(confs, 64, 60, _lib.FLAG_BN)
    hp, hb = torch.zeros(int(g.p_off[-1])), torch.zeros(int(g.b_off[-1]))
    torch.manual_seed(5)
    ntu.init_host_arenas(g, hp, hb)
    torch.manual_seed(5)
    for c, conf in enumerate(confs):
        m = ntu.Searchable_Skeleton_Image_Net(args, conf)
        for k, v in m.state_dict().items():
            if k.endswith("tracked"):
                continue
            arena, off, shape = g.slots[c][k]
            base = hb if arena == "b" else hp
            o = int(g.b_off[c] if arena == "b" else g.p_off[c]) + int(off)
            assert torch.equal(base[o:o + int(np.prod(shape))].view(shape), v), (c, k)


@pytest.mark.parametrize("mode", ["torch", "fast"])
def test_bulk_mt19937_initialiser_is_torch_bit_for_bit(mode, monkeypatch):
    """csrc/host_init.cpp restates torch's CPU mt19937 + uniform_ stream in bulk: same weights as module
    construction (which is what the reference does per candidate) AND the same generator state afterwards."""
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200 import _lib, host_init
    from mfas_b200.engine import GroupLayout
    monkeypatch.setenv("MFAS_HOST_INIT", mode)
    if mode == "fast":
        assert host_init._self_check() >= 0, "bulk helper does not reproduce torch's uniform_ on this CPU"
    confs = [np.array(FOUND_CONFS[4]), np.array([[2, 3, 1]]), np.array(FOUND_CONFS[1])]
    args = make_args(128, 64, 1, bn=True)
    g = GroupLayout(confs, 128, 60, _lib.FLAG_BN)
    hp, hb = torch.zeros(int(g.p_off[-1])), torch.zeros(int(g.b_off[-1]))
    torch.manual_seed(11)
    torch.rand(1000)                                   # start in the middle of a 624-word block
    ntu.init_host_arenas(g, hp, hb)
    after_fast = torch.cat([torch.empty(1).normal_(), torch.rand(5)])     # 9 alphas: a Box-Muller sample is cached in the generator
    torch.manual_seed(11)
    torch.rand(1000)
    for c, conf in enumerate(confs):
        m = ntu.Searchable_Skeleton_Image_Net(args, conf)
        for k, v in m.state_dict().items():
            if k.endswith("tracked"):
                continue
            arena, off, shape = g.slots[c][k]
            base = hb if arena == "b" else hp
            o = int(g.b_off[c] if arena == "b" else g.p_off[c]) + int(off)
            assert torch.equal(base[o:o + int(np.prod(shape))].view(shape), v), (c, k)
    assert torch.equal(after_fast, torch.cat([torch.empty(1).normal_(), torch.rand(5)])), "generator state diverged after the initialisation"


def test_scheduler_matches_reference_semantics():
    from mfas_b200.scheduler import LRCosineAnnealingScheduler
    from oracle.mfas_oracle import CosineRestartLR
    for nb in (4.0, 4.5, 160.0, 7.5):
        a, b = LRCosineAnnealingScheduler(1e-3, 1e-6, 1, 2, nb), CosineRestartLR(1e-3, 1e-6, 1, 2, nb)
        assert [a.step() for _ in range(700)] == [b.step() for _ in range(700)]
        assert a.Ti == b.Ti


def test_loader_orders_are_placement_independent():
    from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache
    c = synthetic_ntu_cache(50, 3)
    a, b = FeatureCacheLoader(c, 8, True, 7), FeatureCacheLoader(c, 8, True, 7)
    k = a.take_passes(6)
    assert k == 0 and a.passes == 6
    assert torch.equal(a.order_for_pass(4), b.order_for_pass(4))
    # a rank's share of a sharded call (any subset of passes, one batched sort) sees the orders of the full range
    full, share = a.orders(0, 6), a.orders_of([5, 1, 3])
    assert torch.equal(share[0], full[5]) and torch.equal(share[1], full[1]) and torch.equal(share[2], full[3])
    batches = list(iter(b))
    assert len(batches) == 7 and batches[-1]['rgb'].shape == (2, 5632) and set(batches[0]) == {'rgb', 'ske', 'label'}
    assert torch.equal(torch.cat([x['label'] for x in batches]), c.labels[b.order_for_pass(0)])


def _dist_worker(rank, world, port, q):
    import torch.distributed as td
    from mfas_b200 import dist as mdist
    from mfas_b200.cache import synthetic_ntu_cache
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 7
        mine = mdist.my_share(n)
        vals = torch.zeros(n, dtype=torch.float64)
        for j in mine:
            vals[j] = 10.0 + j                     # "accuracy" of the candidates this rank trained
        full = mdist.gather_results(vals, n)
        cache = synthetic_ntu_cache(16, 9, with_backbone_logits=True) if rank == 0 else None
        got = mdist.broadcast_cache(cache, "cpu")
        assert got.logit_rgb is not None and got.logit_rgb.shape == (16, 60)      # multitask flows work on a broadcast cache
        # the reference's driver samples with unseeded numpy: every rank draws its own list; rank 0's list and seed win
        import warnings
        rng = np.random.RandomState(100 + rank)
        confs = [rng.randint(0, 4, size=(1 + rng.randint(3), 3)) for _ in range(5)]
        with warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            synced, seed = mdist.sync_call_inputs(confs, 1000 + rank)
        assert (len(wlist) == 1) == (rank != 0)
        from mfas_b200.mmimdb_searchable import WIDTHS, synthetic_mmimdb_cache     # multi-hot targets + pos_weight ride along
        ml = mdist.broadcast_cache(synthetic_mmimdb_cache(12, 4) if rank == 0 else None, "cpu")
        assert ml.multilabel and ml.widths == WIDTHS and ml.labels.shape == (12, 23)
        q.put((rank, mine, full.tolist(), float(got.rgb_cat.sum()), int(got.labels.sum()),
               float(ml.labels.sum() + ml.pos_weight.sum() + ml.ske_cat.sum()), [c.tolist() for c in synced], seed,
               float(got.logit_rgb.sum() + got.logit_ske.sum())))
    finally:
        td.destroy_process_group()


def test_sharding_and_gather_world2_gloo():
    """N>1 host logic on CPU: round-robin ownership, order-preserving gather, cache broadcast."""
    import torch.multiprocessing as mp
    from mfas_b200.cache import synthetic_ntu_cache
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    ref = synthetic_ntu_cache(16, 9, with_backbone_logits=True)
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    rng = np.random.RandomState(100)
    want = [rng.randint(0, 4, size=(1 + rng.randint(3), 3)).tolist() for _ in range(5)]
    assert res[0][6] == res[1][6] == want and res[0][7] == res[1][7] == 1000          # rank 0's list and seed on both ranks
    assert res[0][8] == res[1][8] == pytest.approx(float(ref.logit_rgb.sum() + ref.logit_ske.sum()))
    for r in res:
        assert r[2] == [10.0 + j for j in range(7)]                      # every rank sees all results, in input order
        assert r[3] == pytest.approx(float(ref.rgb_cat.sum())) and r[4] == int(ref.labels.sum())
    from mfas_b200.mmimdb_searchable import synthetic_mmimdb_cache
    ml = synthetic_mmimdb_cache(12, 4)
    assert res[0][5] == res[1][5] == pytest.approx(float(ml.labels.sum() + ml.pos_weight.sum() + ml.ske_cat.sum()))


def test_sliced_threaded_initialisation_equals_one_serial_fill(monkeypatch):
    """train_sampled_models fills the pinned arena a quarter of the candidates at a time (so the H2D copies overlap) and
    the C helper pipelines fills of >= 4 M words over threads: same bytes and same generator state as one serial fill."""
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200 import _lib, host_init
    from mfas_b200.engine import GroupLayout
    if host_init._self_check() < 0:
        pytest.skip("bulk helper does not reproduce torch's uniform_ on this CPU")
    confs = [np.array(FOUND_CONFS[4])] * 21 + [np.array([[2, 3, 1]])] * 3      # 22 M words; an odd number of alphas
    g = GroupLayout(confs, 128, 60, _lib.FLAG_BN)
    outs = []
    for threads, sliced in (("0", False), ("2", True), ("3", False)):
        monkeypatch.setenv("MFAS_HOST_INIT_THREADS", threads)
        hp, hb = torch.zeros(int(g.p_off[-1])), torch.zeros(int(g.b_off[-1]))
        torch.manual_seed(123)
        torch.rand(77)
        if sliced:
            for c0 in range(0, g.n, 7):
                ntu.init_host_arenas(g, hp, hb, slots=range(c0, min(g.n, c0 + 7)))
        else:
            ntu.init_host_arenas(g, hp, hb)
        outs.append((hp, hb, torch.cat([torch.empty(1).normal_(), torch.rand(3)])))
    for hp, hb, after in outs[1:]:
        assert torch.equal(hp, outs[0][0]) and torch.equal(hb, outs[0][1])
        assert torch.equal(after, outs[0][2]), "generator state diverged"


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in mfas_b200/_lib.py against the C compiler's view of include/mfas_b200.h: sizes and the
    offsets of every field (an ABI drift between the header and the binding would otherwise only show on a GPU)."""
    import ctypes as C
    import subprocess
    from mfas_b200 import _lib
    pairs = {"mfas_layout": _lib.Layout, "mfas_cache_desc": _lib.CacheDesc, "mfas_arenas": _lib.Arenas,
             "mfas_adam_hparams": _lib.AdamHParams, "mfas_run_args": _lib.RunArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/mfas_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("abi %d flags %d %d\\n", MFAS_ABI_VERSION, MFAS_FLAG_MULTITASK, MFAS_FLAG_MULTILABEL);', '  return 0;', '}']
    src = tmp_path / "abi_probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_probe"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = dict(l.split(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines())
    for cname, cls in pairs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    assert out["abi"] == f"{_lib.ABI_VERSION} flags {_lib.FLAG_MULTITASK} {_lib.FLAG_MULTILABEL}"


def test_argument_errors_need_no_gpu():
    """Argument validation happens before any CUDA call: bad arguments are MFAS_ERR_INVALID with a message, GPU or not."""
    import ctypes as C
    from mfas_b200 import _lib
    lib = _lib.lib()
    buf = (C.c_float * 8)()
    assert lib.mfas_global_pool(0, None, 1, 1, 1, None, 1, None) == -1
    assert lib.mfas_global_pool(0, C.addressof(buf), 2, 4, 1, C.addressof(buf), 3, None) == -1      # out_ld < C
    assert b"out_ld" in lib.mfas_last_error()
    assert lib.mfas_group_create(0, 0, None, 64, 0.0, 0, None, None) == -1
    lay = _lib.Layout()
    d = (C.c_int32 * _lib.NUM_TAPS)(*([64, 128] * (_lib.NUM_TAPS // 2)))
    conf = (C.c_int32 * 3)(0, 0, 7)
    assert lib.mfas_plan_layout(1, conf, 64, 23, _lib.FLAG_BN | _lib.FLAG_MULTILABEL, d, d, C.byref(lay)) == -1      # activation 7
    assert b"activation" in lib.mfas_last_error()


def test_backward_tile_list_covers_every_weight_once_and_never_straddles_a_source():
    """Tiling rules of the persistent weight-gradient kernel (host only, mfas_plan_bwd_tiles): for any tap set the tiles of a
    layer partition its [H, K] weight matrix, no tile reads x columns from two concat sources, the NTU list is the plain
    128-column walk it always was, classifier tiles come last."""
    import ctypes as C
    from mfas_b200 import _lib
    from mfas_b200.engine import plan_layout
    lib = _lib.lib()

    def tiles_of(layouts, head):
        arr = (_lib.Layout * len(layouts))(*layouts)
        n, nl = C.c_int64(), C.c_int64()
        assert lib.mfas_plan_bwd_tiles(arr, len(layouts), int(head), None, 0, C.byref(n), C.byref(nl)) == 0
        out = (C.c_int32 * (4 * n.value))()
        assert lib.mfas_plan_bwd_tiles(arr, len(layouts), int(head), C.addressof(out), n.value, C.byref(n), C.byref(nl)) == 0
        return np.array(out, dtype=np.int64).reshape(-1, 4), nl.value

    ntu = plan_layout(FOUND_CONFS[4], 128, 60, _lib.FLAG_BN)
    t, nl = tiles_of([ntu], head=True)
    plain = [(0, l, k, h) for l in range(4) for k in range(0, ntu.K[l], 128) for h in range(0, 128, 64)]
    assert [tuple(r) for r in t[:nl]] == plain
    assert [tuple(r) for r in t[nl:]] == [(0, 4, 0, 0)]                       # W_c [60, 128]: one classifier tile

    rng = np.random.default_rng(0)
    for trial in range(20):
        d0 = [int(32 * rng.integers(1, 9)) for _ in range(4)]
        d1 = [int(32 * rng.integers(1, 20)) for _ in range(4)]
        L, H = int(rng.integers(1, 5)), int(rng.choice([64, 128, 192, 256]))
        conf = [[int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(0, 3))] for _ in range(L)]
        lays = [plan_layout(conf, H, 23, _lib.FLAG_BN | _lib.FLAG_MULTILABEL, widths=(d0, d1)),
                plan_layout(conf[:1], H, 23, _lib.FLAG_BN | _lib.FLAG_MULTILABEL, widths=(d0, d1))]
        t, nl = tiles_of(lays, head=False)
        assert nl == len(t)
        for c, lay in enumerate(lays):
            for l in range(lay.L):
                cover = np.zeros((H, lay.K[l]), dtype=np.int32)
                bounds = (lay.d_ske[l], lay.d_ske[l] + lay.d_rgb[l], lay.K[l])
                for _, _, kc0, h0 in t[(t[:, 0] == c) & (t[:, 1] == l)]:
                    seg_end = next(b for b in bounds if kc0 < b)
                    kw = min(128, seg_end - kc0)
                    assert kw > 0 and kw % 16 == 0
                    cover[h0:h0 + 64, kc0:kc0 + kw] += 1
                assert (cover == 1).all(), (trial, c, l)


def test_share_sized_initialisation_equals_the_full_call():
    """A rank that trains a share of a call draws the constructor's stream for EVERY candidate of the call (so a candidate gets
    the same weights wherever it runs) but stages only its own: same bytes as the matching slices of a full fill, same
    generator state afterwards, staging arena of the share's size (+ one scratch candidate)."""
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200 import _lib
    from mfas_b200.engine import GroupLayout
    confs = [np.array(FOUND_CONFS[4]), np.array([[0, 0, 0]]), np.array(FOUND_CONFS[1][:2]), np.array([[2, 3, 1], [1, 1, 0]]),
             np.array(FOUND_CONFS[2][:3])]
    flags = _lib.FLAG_BN | _lib.FLAG_ALPHAS
    full = GroupLayout(confs, 64, 60, flags)
    torch.manual_seed(3)
    hp, hb = torch.zeros(int(full.p_off[-1])), torch.zeros(int(full.b_off[-1]))
    ntu.init_host_arenas(full, hp, hb)
    state_full = torch.get_rng_state()
    for mine in ([0, 2, 4], [1, 3], []):
        g = GroupLayout([confs[i] for i in mine], 64, 60, flags) if mine else None
        rm = ntu._ShareLayout(full, mine, g)
        assert rm.n_p < int(full.p_off[-1])
        torch.manual_seed(3)
        p2, b2 = torch.zeros(rm.n_p), torch.zeros(rm.n_b)
        ntu.init_host_arenas(rm, p2, b2)
        assert torch.equal(torch.get_rng_state(), state_full)
        for k, j in enumerate(mine):
            assert torch.equal(p2[int(g.p_off[k]):int(g.p_off[k + 1])], hp[int(full.p_off[j]):int(full.p_off[j + 1])]), (mine, k)
            assert torch.equal(b2[int(g.b_off[k]):int(g.b_off[k + 1])], hb[int(full.b_off[j]):int(full.b_off[j + 1])]), (mine, k)
