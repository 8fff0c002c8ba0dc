"""Plug the B200 path into the *unmodified* reference tree.

The reference driver binds the hot path by module attribute
(`import models.search.ntu_searchable as ntu`, /root/reference/models/searchable.py:23, used at :256-260)
and `main_found_ntu.py` by `import models.search.train_searchable.ntu as tr` (:19).  `install()` makes those
imports resolve to this package; nothing in the reference is edited.
"""
from __future__ import annotations

import sys
import types


def install(shim_broken_imports: bool = True):
    """Call once, after the reference root is on sys.path and before `import models.searchable`."""
    import importlib
    from . import avmnist_searchable, ntu_searchable, scheduler, train_ntu
    if shim_broken_imports:
        # the reference as shipped does not import: matplotlib is an unused import of
        # models/utils.py:61, `models.aux` / `models.train` are dangling names in loops we never call
        try:                                       # a real matplotlib wins: the stub is only for boxes that have none
            importlib.import_module("matplotlib.pyplot")
        except ImportError:
            for n in ("matplotlib", "matplotlib.pyplot"):
                sys.modules.setdefault(n, types.ModuleType(n))
        try:
            ref_sched = importlib.import_module("models.auxiliary.scheduler")   # the reference's own class is fine:
        except ImportError:                                                       # train_ntu duck-types it
            ref_sched = scheduler
        for n in ("models.aux", "models.train"):
            pkg = types.ModuleType(n)
            pkg.scheduler = ref_sched
            sys.modules.setdefault(n, pkg)
            sys.modules.setdefault(n + ".scheduler", ref_sched)
    for name, mod in (("models.search.ntu_searchable", ntu_searchable),
                      ("models.search.train_searchable.ntu", train_ntu),
                      ("models.search.avmnist_searchable", avmnist_searchable),          # the AV-MNIST twin: module + its loops
                      ("models.search.train_searchable.avmnist", avmnist_searchable)):
        sys.modules[name] = mod
        parent, _, leaf = name.rpartition(".")
        try:
            setattr(importlib.import_module(parent), leaf, mod)
        except ImportError:
            pass
    try:                                       # K x 32 surrogate evaluations per search step as one batch (tools.py:22-30)
        from . import search_tools
        importlib.import_module("models.search.tools").predict_accuracies_with_surrogate = search_tools.predict_accuracies_with_surrogate
    except ImportError:
        pass
    return ntu_searchable


def cached_ntu_searcher(S, args, device, train_cache, dev_cache, seed=0):
    """An `NTUSearcher` (models/searchable.py:233-260) whose dataloaders walk a feature cache instead of decoding
    NTU videos; `.search()` is the reference's own `_epnas` loop.  `S` is the imported `models.searchable`."""
    from .cache import FeatureCacheLoader

    class CachedNTUSearcher(S.NTUSearcher):
        def __init__(self):
            S.ModelSearcher.__init__(self, args)
            self.device = device
            self.dataloaders = {"train": FeatureCacheLoader(train_cache, args.batchsize, True, seed),
                                "dev": FeatureCacheLoader(dev_cache, args.batchsize, True, seed + 50000)}

    return CachedNTUSearcher()


def cached_mmimdb_searcher(S, args, device, train_cache, dev_cache, seed=0):
    """The searcher the reference lacks for MM-IMDB, by analogy with `NTUSearcher`
    (models/searchable.py:233-260): a `ModelSearcher` whose `search()` hands
    `mmimdb_searchable.train_sampled_models` / `get_possible_layer_configurations` / `Searchable_Text_Image_Net` to the
    reference's own `_epnas` loop.  `train_cache` / `dev_cache`: `mmimdb_searchable.text_image_cache(...)` splits."""
    from . import mmimdb_searchable as mm

    class CachedMMIMDBSearcher(S.ModelSearcher):
        def __init__(self):
            S.ModelSearcher.__init__(self, args)
            self.device = device
            self.dataloaders = {"train": mm.TextImageCacheLoader(train_cache, args.batchsize, True, seed),
                                "dev": mm.TextImageCacheLoader(dev_cache, args.batchsize, True, seed + 50000)}

        def search(self):
            surrogate = S.surr.SimpleRecurrentSurrogate(100, 3, 100)          # as NTUSearcher.search, :253-260
            surrogate.to(self.device)
            surrogate_dict = {'model': surrogate, 'criterion': S.torch.nn.MSELoss()}
            searchmethods = {'train_sampled_fun': mm.train_sampled_models,
                             'get_layer_confs': mm.get_possible_layer_configurations}
            return self._epnas(mm.Searchable_Text_Image_Net, surrogate_dict, self.dataloaders, searchmethods, self.device)

    return CachedMMIMDBSearcher()
