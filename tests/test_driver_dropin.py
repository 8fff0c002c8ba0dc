"""Drop-in boundary at the level of the reference's search driver (CPU, build container only): the UNMODIFIED
`ModelSearcher._epnas` loop (/root/reference/models/searchable.py:48-137) is run over this package's modules -- search-space
functions, configuration format, return types -- with the candidate trainer stubbed (training itself needs a GPU and is
covered by the `-m gpu` tests).  Skipped where /root/reference does not exist (the GPU box)."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="the reference tree is not present")


def _search_args(**kw):
    d = dict(lr_surrogate=1e-3, epochs_surrogate=3, search_iterations=2, max_progression_levels=2, num_samples=3,
             initial_temperature=10.0, final_temperature=0.2, temperature_decay=4.0, verbose=False, batchsize=16, epochs=1,
             inner_representation_size=64, num_outputs=23, drpt=0.0, batchnorm=True, alphas=False, multitask=False,
             weightsharing=False, use_dataparallel=False, eta_max=1e-3, eta_min=1e-6, Ti=1, Tm=2, vid_len=(8, 32),
             checkpointdir="", ske_cp="ske", rgb_cp="rgb")
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.fixture()
def ref_searchable():
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    sys.path.insert(0, REF)
    import mfas_b200.install as b200
    b200.install()
    import models.searchable as S
    yield S, b200
    sys.path[:] = saved_path
    for k in list(sys.modules):
        if k not in saved_mods and (k == "models" or k.startswith("models.") or k.startswith("matplotlib")):
            del sys.modules[k]


def test_reference_epnas_runs_over_the_mmimdb_module(ref_searchable, monkeypatch):
    S, b200 = ref_searchable
    import mfas_b200.mmimdb_searchable as mm
    train, dev = mm.synthetic_mmimdb_cache(64, 1), mm.synthetic_mmimdb_cache(32, 2)
    calls = []

    def fake_train(confs, searchable_type, dataloaders, args, device, state_dict=dict(), **kw):
        assert searchable_type is mm.Searchable_Text_Image_Net and set(dataloaders) == {"train", "dev"}
        for c in confs:
            c = np.asarray(c)
            assert c.ndim == 2 and c.shape[1] == 3 and c[:, 0].max() < 2 and c[:, 1].max() < 4 and c[:, 2].max() < 2
            searchable_type(args, c)                                  # every sampled configuration is constructible
        calls.append(len(confs))
        # what the real trainer returns: 0-dim float64 CPU tensors (np.array(accs) must work)
        return [torch.tensor(0.3 + 0.01 * (int(np.asarray(c).sum()) % 7), dtype=torch.float64) for c in confs]

    monkeypatch.setattr(mm, "train_sampled_models", fake_train)
    torch.manual_seed(0); np.random.seed(0)
    s_data = b200.cached_mmimdb_searcher(S, _search_args(), torch.device("cpu"), train, dev).search()
    assert calls[0] == 16 and all(c == 3 for c in calls[1:]) and len(calls) == 4      # 16 one-step rows, then K = 3 per step
    confs, accs = s_data.get_data(to_torch=False)            # grouped by depth: arrays [L, n, 3]; re-sampled confs overwrite
    assert 16 <= sum(np.asarray(c).shape[1] for c in confs) <= 16 + 3 * 3 and {np.asarray(c).shape[0] for c in confs} == {1, 2}


def test_reference_epnas_runs_over_the_ntu_module(ref_searchable, monkeypatch):
    S, b200 = ref_searchable
    import mfas_b200.ntu_searchable as ntu
    from mfas_b200.cache import synthetic_ntu_cache
    assert sys.modules["models.search.ntu_searchable"] is ntu and S.ntu is ntu           # install() rebinds the driver's import
    calls = []

    def fake_train(confs, searchable_type, dataloaders, args, device, state_dict=dict(), **kw):
        assert searchable_type is ntu.Searchable_Skeleton_Image_Net
        calls.append(len(confs))
        return [torch.tensor(0.5 + 0.001 * i, dtype=torch.float64) for i, _ in enumerate(confs)]

    monkeypatch.setattr(ntu, "train_sampled_models", fake_train)
    torch.manual_seed(0); np.random.seed(0)
    searcher = b200.cached_ntu_searcher(S, _search_args(num_outputs=60), torch.device("cpu"), synthetic_ntu_cache(32, 1),
                                        synthetic_ntu_cache(16, 2))
    searcher.search()
    assert calls[0] == 32 and all(c == 3 for c in calls[1:]) and len(calls) == 4


def test_batched_surrogate_evaluation_equals_the_reference_loop(ref_searchable):
    """mfas_b200.search_tools.predict_accuracies_with_surrogate (one surrogate forward per depth) against the reference's
    one-by-one loop (models/search/tools.py:22-30 over surrogate.eval_model) on the reference's own surrogate module."""
    S, b200 = ref_searchable
    import importlib
    import models.search.surrogate as surr
    from mfas_b200.search_tools import predict_accuracies_with_surrogate as batched
    tools = importlib.import_module("models.search.tools")
    assert tools.predict_accuracies_with_surrogate is batched                      # install() rebinds it
    torch.manual_seed(3)
    surrogate = surr.SimpleRecurrentSurrogate(100, 3, 100)
    rng = np.random.RandomState(0)
    confs = [rng.randint(0, 4, size=(L, 3)) for L in (1, 2, 2, 3, 1, 2, 4, 3, 2)] + [rng.randint(0, 4, size=(2, 3)) for _ in range(96)]
    want = [surrogate.eval_model(c, torch.device("cpu")) for c in confs]
    got = batched(confs, surrogate, torch.device("cpu"))
    assert len(got) == len(want) and all(isinstance(g, np.floating) or np.ndim(g) == 0 for g in got)
    assert np.abs(np.array(got, np.float64) - np.array(want, np.float64)).max() < 1e-6
    assert np.array(got).shape == (len(confs),)                                     # np.array(accs) as the driver does (tools.py:47)


@pytest.mark.parametrize("script,fixtures", [("gen_golden_pooling.py", ["pooling.npz"]), ("gen_golden_mmimdb.py", ["mmimdb_head.npz"]),
                                             ("gen_golden_mmimdb_path.py", ["mmimdb_path.npz"]), ("gen_golden_found.py", ["found_mt.npz"]),
                                             ("gen_golden_avmnist.py", ["avmnist.npz"]),
                                             ("gen_golden.py", ["cfg1.npz", "cfg2.npz", "mixL.npz", "alph.npz", "wsh.npz"])])
def test_committed_fixtures_are_what_executing_the_reference_produces(script, fixtures, tmp_path):
    """Provenance of the golden vectors: re-run the committed generator (it executes the reference's own classes / loop)
    into a scratch directory and compare with the committed .npz, array by array."""
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    env = dict(os.environ, MFAS_GOLDEN_OUT=str(tmp_path))
    subprocess.run([sys.executable, os.path.join(here, script)], check=True, env=env, capture_output=True, timeout=600)
    for fixture in fixtures:
        new, old = np.load(tmp_path / fixture), np.load(os.path.join(here, fixture))
        assert sorted(new.files) == sorted(old.files), fixture
        for k in old.files:
            a, b = np.asarray(new[k]), np.asarray(old[k])
            assert a.shape == b.shape and a.dtype == b.dtype, (fixture, k)
            if a.dtype.kind == "f":      # same machine, same torch: bit-identical in practice; allow the last ulps of a threaded reduction
                assert np.allclose(a, b, rtol=1e-6, atol=1e-9, equal_nan=True), (fixture, k)
            else:
                assert np.array_equal(a, b), (fixture, k)
