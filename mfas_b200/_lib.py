"""ctypes binding of the C ABI declared in include/mfas_b200.h.

There is no CPU fallback: if the shared library is missing this module raises, and every
compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# MFAS_LIB_PATH: another build of the SAME sources (profiles/run_gpu_sanitize.sh: barrier-wait bounds raised for compute-sanitizer)
LIB_PATH = os.environ.get("MFAS_LIB_PATH") or os.path.join(HERE, "_mfas_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "mfas_abi.cu"), os.path.join(HERE, "csrc", "host_init.cpp")]
HEADERS = [os.path.join(HERE, "csrc", f) for f in ("common.cuh", "kernels_ffma.cuh", "kernels_tc.cuh", "kernels_pool.cuh", "umma.cuh")] + [
    os.path.join(ROOT, "include", "mfas_b200.h")]

ABI_VERSION = 5          # MFAS_ABI_VERSION of include/mfas_b200.h
MAX_LAYERS, MAX_BATCH, MAX_HIDDEN, MAX_CLASSES, NUM_TAPS = 8, 128, 256, 64, 8
FLAG_BN, FLAG_DROPOUT, FLAG_ALPHAS, FLAG_MULTITASK, FLAG_MULTILABEL, FLAG_PLAIN = 1, 2, 4, 8, 16, 32
ERRORS = {0: "MFAS_OK", -1: "MFAS_ERR_INVALID", -2: "MFAS_ERR_CUDA", -3: "MFAS_ERR_UNSUPPORTED",
          -4: "MFAS_ERR_NOMEM", -5: "MFAS_ERR_UNBOUND"}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "--cudart", "static"]


class MfasError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Layout(C.Structure):
    _fields_ = [("L", C.c_int32), ("H", C.c_int32), ("C", C.c_int32), ("flags", C.c_int32),
                ("conf", (C.c_int32 * 3) * MAX_LAYERS),
                ("d_ske", C.c_int32 * MAX_LAYERS), ("d_rgb", C.c_int32 * MAX_LAYERS), ("K", C.c_int32 * MAX_LAYERS),
                ("off_W", C.c_int64 * MAX_LAYERS), ("off_b", C.c_int64 * MAX_LAYERS),
                ("off_gamma", C.c_int64 * MAX_LAYERS), ("off_beta", C.c_int64 * MAX_LAYERS),
                ("off_alpha", C.c_int64 * MAX_LAYERS),
                ("off_Wc", C.c_int64), ("off_bc", C.c_int64), ("n_params", C.c_int64),
                ("off_rm", C.c_int64 * MAX_LAYERS), ("off_rv", C.c_int64 * MAX_LAYERS), ("n_bufs", C.c_int64)]


class CacheDesc(C.Structure):
    _fields_ = [("n_rows", C.c_int64),
                ("ske", C.c_void_p * NUM_TAPS), ("rgb", C.c_void_p * NUM_TAPS),
                ("ske_ld", C.c_int64 * NUM_TAPS), ("rgb_ld", C.c_int64 * NUM_TAPS),
                ("d_ske", C.c_int32 * NUM_TAPS), ("d_rgb", C.c_int32 * NUM_TAPS),
                ("labels", C.c_void_p), ("logit_rgb", C.c_void_p), ("logit_ske", C.c_void_p),
                ("targets", C.c_void_p), ("pos_weight", C.c_void_p)]


class Arenas(C.Structure):
    _fields_ = [("params", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p), ("grad", C.c_void_p),
                ("bufs", C.c_void_p), ("nbt", C.c_void_p)]


class AdamHParams(C.Structure):
    _fields_ = [("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double)]


class RunArgs(C.Structure):
    _fields_ = [("n_epochs", C.c_int32), ("batch", C.c_int32),
                ("perm_train", C.c_void_p), ("perm_dev", C.c_void_p),
                ("step_size", C.c_void_p), ("bc2_sqrt", C.c_void_p),
                ("adam_t0", C.c_int64), ("stats", C.c_void_p), ("best_acc", C.c_void_p), ("best_epoch", C.c_void_p),
                ("best_acc_init", C.c_void_p)]


# every symbol include/mfas_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "mfas_abi_version": (C.c_int, []),
    "mfas_last_error": (C.c_char_p, []),
    "mfas_plan_layout": (C.c_int, [C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(Layout)]),
    "mfas_algorithmic_counts": (C.c_int, [C.POINTER(Layout), C.c_int32, C.POINTER(C.c_double)]),
    "mfas_group_create": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(Layout), C.c_int32, C.c_float, C.c_uint32,
                                    C.POINTER(C.c_int32), C.POINTER(_P)]),
    "mfas_group_destroy": (C.c_int, [_P]),
    "mfas_release_cached_memory": (C.c_int, []),
    "mfas_group_bind": (C.c_int, [_P, C.c_int32, C.POINTER(Arenas)]),
    "mfas_group_set_adam": (C.c_int, [_P, C.POINTER(AdamHParams)]),
    "mfas_group_init_params": (C.c_int, [_P, C.c_uint64, _P]),
    "mfas_group_num_launches": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "mfas_group_engine": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "mfas_group_status": (C.c_int, [_P]),
    "mfas_forward": (C.c_int, [_P, C.POINTER(CacheDesc), _P, C.c_int64, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, _P]),
    "mfas_train_step": (C.c_int, [_P, C.POINTER(CacheDesc), _P, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_int64,
                                  _P, _P, _P, _P]),
    "mfas_train_run": (C.c_int, [_P, C.POINTER(CacheDesc), C.POINTER(CacheDesc), C.POINTER(RunArgs), _P]),
    "mfas_eval_pass": (C.c_int, [_P, C.POINTER(CacheDesc), _P, C.c_int32, _P, _P]),
    "mfas_group_set_profiling": (C.c_int, [_P, C.c_int32]),
    "mfas_group_last_step_ms": (C.c_int, [_P, _P]),
    "mfas_group_chain_timeline": (C.c_int, [_P, _P, C.c_int32]),
    "mfas_plan_bwd_tiles": (C.c_int, [C.POINTER(Layout), C.c_int32, C.c_int32, _P, C.c_int64, _P, _P]),
    "mfas_global_pool": (C.c_int, [C.c_int32, _P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, _P]),
    "mfas_host_uniform_fill": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, _P, C.c_int32]),
}


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building " + LIB_PATH)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library (raises if it has not been built -- no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(mfas_b200 has no CPU / PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.mfas_abi_version() != ABI_VERSION:
            raise RuntimeError("ABI version mismatch; rebuild the library")
        _lib = l
    return _lib


def check(code: int):
    if code != 0:
        raise MfasError(code, lib().mfas_last_error().decode())
