"""Diagnostic (GPU box): clock stamps of the roles (loader, converter, MMA issuer, epilogue / Adam) of k_tc_fwd_small (CTA 0,
k-blocks 32..47) and k_tc_bwd_small (CTA 0, fills 16..31), in SM clocks.  The stamps are compiled in with -DMFAS_KSTAMPS only:
the same sources are built into /tmp and loaded through MFAS_LIB_PATH (as profiles/run_gpu_sanitize.sh does)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["MFAS_CHAIN_TIMELINE"] = "1"
LIB = "/tmp/_mfas_kstamps.so"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--cudart", "static",
                "-DMFAS_KSTAMPS", "-w", "-o", LIB, os.path.join(ROOT, "mfas_b200/csrc/mfas_abi.cu"), os.path.join(ROOT, "mfas_b200/csrc/host_init.cpp")], check=True)
os.environ["MFAS_LIB_PATH"] = LIB
import numpy as np, torch
from mfas_b200 import _lib
from mfas_b200.cache import synthetic_ntu_cache
from mfas_b200.engine import CandidateGroup
rows32 = [[i, j, k] for i in range(4) for j in range(4) for k in range(2)]
parents = [[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0], [2, 2, 0], [0, 0, 1], [3, 2, 0], [2, 3, 1]]
M, H = int(os.environ.get("TL_M", "256")), int(os.environ.get("TL_H", "16"))
confs = [np.array([p, r]) for p in parents for r in rows32][:M]
if H > 32:                                            # cfg2: the stamps are then those of k_tc_fwd_ws (the backward table stays empty)
    confs = [np.array([[3, 1, 1], [1, 3, 0], [1, 1, 1], [3, 3, 0]])] * M
B = 64
train = synthetic_ntu_cache(4096, 3).to("cuda:0")
g = CandidateGroup(confs, H, 60, _lib.FLAG_BN, "cuda:0", batch_max=B)
g.set_adam(0.9, 0.999, 1e-8, 1e-4); g.params.uniform_(-0.03, 0.03); g.bufs.fill_(1.0)
rows = torch.stack([torch.randperm(4096)[:B] for _ in range(M)]).to("cuda:0", torch.int32)
for it in range(6):
    g.train_step(train, rows, 1e-4)
out = np.zeros((M, 16), np.int64)
_lib.check(_lib.lib().mfas_group_chain_timeline(g._h, out.ctypes.data, M))
t = out.reshape(-1)[:128].reshape(16, 8)
t0 = t[0, 0]
names = ["ld issued", "-", "cv saw landed", "cv split done", "mma saw split", "mma issued", "(fwd_ws: cv saw lofree)", "ld saw free"]
print("k-block | " + " | ".join(names[:8]) + "   (cycles since k-block 32 was issued)")
for i in range(16):
    print("%7d | " % (32 + i) + " | ".join("%8d" % (t[i, e] - t0) for e in range(8)))
d = np.diff(t[:, 5])
print("mma issue interval: median %d cycles; loader issue interval: median %d" % (np.median(d), np.median(np.diff(t[:, 0]))))

t = out.reshape(-1)[128:256].reshape(16, 8)
t0 = t[0, 0]
names = ["ld issued", "cv saw landed", "cv arrived lofull", "mma saw lofull", "mma issued", "adam saw tfull", "adam done", "ld saw rawfree"]
if H > 32: names = ["stage stored (full)", "mma saw tempty", "-", "mma saw full", "mma issued", "adam saw tfull", "adam done", "stagers saw empty"]
print(("k_tc_bwd_ws" if H > 32 else "k_tc_bwd_small") + ", CTA 0, fills 16..31 (cycles since fill 16 was issued / stored)")
print("   fill | " + " | ".join(names))
for i in range(16):
    print("%7d | " % (16 + i) + " | ".join("%8d" % (t[i, e] - t0) for e in range(8)))
print("mma issue interval: median %d cycles; loader issue interval: median %d; adam done interval: median %d" % (np.median(np.diff(t[:, 4])), np.median(np.diff(t[:, 0])), np.median(np.diff(t[:, 6]))))
