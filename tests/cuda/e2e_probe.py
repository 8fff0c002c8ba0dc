import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
os.environ["MFAS_TIMING"] = "1"
import numpy as np, torch
from helpers import make_args
import mfas_b200.ntu_searchable as ntu
from mfas_b200.cache import FeatureCacheLoader, synthetic_ntu_cache
CONF4 = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]]
ht, hd = synthetic_ntu_cache(10240, 1).pin(), synthetic_ntu_cache(5120, 2).pin()
loaders = {"train": FeatureCacheLoader(ht, 64, True, 100), "dev": FeatureCacheLoader(hd, 64, True, 200)}
args = make_args(128, 64, 3, bn=True, drpt=0.0, Ti=1)
args.init_on_device = os.environ.get("PROBE_DEVICE_INIT") == "1"
confs = [np.array(CONF4) for _ in range(int(os.environ.get("PROBE_M", "148")))]
for it in range(3):
    ht.drop_device_copies(); hd.drop_device_copies()
    t0 = time.perf_counter()
    accs = ntu.train_sampled_models(confs, ntu.Searchable_Skeleton_Image_Net, loaders, args, torch.device("cuda:0"))
    torch.cuda.synchronize()
    sys.stderr.write(f"== call {it}: {(time.perf_counter() - t0) * 1e3:.1f} ms\n")
