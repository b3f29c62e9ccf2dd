"""ctypes binding of libkucd.so (include/kucd.h) and the in-tree build recipe.

The library is the product: there is no Python or CPU implementation behind it.  `load()` raises
when the shared object is missing or was built without its symbols, and every engine call raises
when the C side reports an error (KUCD_ERR_* -> ValueError for argument/shape errors, mirroring
the ValueErrors of /root/reference/ku/ebm/dbn.py:29,48; RuntimeError otherwise).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libkucd.so")
SRC = os.path.join(_HERE, "csrc", "kucd.cu")
HEADER = os.path.join(ROOT, "include", "kucd.h")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC",
]

# ---- kucd.h mirrored ---------------------------------------------------------------------------
KUCD_OK = 0
ERR_INVALID_ARG, ERR_SHAPE_MISMATCH, ERR_UNSUPPORTED_DTYPE = -1, -2, -3
ERR_CUDA, ERR_NCCL, ERR_NOT_SM100, ERR_NOT_BUILT = -4, -5, -6, -7
DEV_CPU, DEV_CUDA, DEV_CUDA_HOST = 1, 2, 3
DT_INT, DT_UINT, DT_FLOAT, DT_BFLOAT = 0, 1, 2, 4
MODE_VISIBLE_BERNOULLI, MODE_VISIBLE_GAUSSIAN = 0, 1
COMPUTE_BF16, COMPUTE_F32X3 = 0, 1
UPDATE_W, UPDATE_C, UPDATE_B, UPDATE_ALL = 1, 2, 4, 7
MAX_K = 32


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device_type", C.c_int32), ("device_id", C.c_int32),
                ("dtype_code", C.c_int32), ("bits", C.c_int32), ("shape", C.c_int64 * 2),
                ("strides", C.c_int64 * 2)]


class HParams(C.Structure):
    _fields_ = [("lr", C.c_float), ("k", C.c_int32), ("persistent", C.c_int32), ("momentum", C.c_float),
                ("weight_decay", C.c_float), ("normalize", C.c_int32), ("update_mask", C.c_int32),
                ("want_stats", C.c_int32)]


class Inject(C.Structure):
    _fields_ = [("u_h", C.POINTER(Tensor) * MAX_K), ("u_v", C.POINTER(Tensor) * MAX_K),
                ("u_hc", C.POINTER(Tensor))]


class StepStats(C.Structure):
    _fields_ = [("score", C.c_float), ("recon_err", C.c_float), ("fe_mean", C.c_float), ("rows", C.c_int32)]


class EpochStats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("rows", C.c_int64), ("device_ms", C.c_float), ("last_score", C.c_float),
                ("last_recon_err", C.c_float)]


class Timings(C.Structure):
    _fields_ = [("gemm_launches", C.c_int64), ("chain_launches", C.c_int64), ("chain_dw_launches", C.c_int64),
                ("aux_launches", C.c_int64),
                ("graph_launches", C.c_int64),
                ("graph_kernel_launches", C.c_int64), ("allreduce_calls", C.c_int64), ("fused_reduce_steps", C.c_int64), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("proj_timed", C.c_int64), ("dw_timed", C.c_int64),
                ("proj_ms", C.c_float), ("dw_ms", C.c_float), ("last_gemm_ms", C.c_float),
                ("unit_steps", C.c_int64), ("unit_exchanges", C.c_int64), ("xchg_timed", C.c_int64),
                ("upd_timed", C.c_int64), ("xchg_ms", C.c_float), ("upd_ms", C.c_float)]


_P = C.c_void_p
_TP = C.POINTER(Tensor)
# name -> (restype, argtypes); one entry per declaration in include/kucd.h
SIGNATURES = {
    "kucd_abi_version": (C.c_int, []),
    "kucd_last_error": (C.c_char_p, []),
    "kucd_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_uint64]),
    "kucd_ctx_destroy": (C.c_int, [_P]),
    "kucd_sync": (C.c_int, [_P]),
    "kucd_get_timings": (C.c_int, [_P, C.POINTER(Timings), C.c_int]),
    "kucd_ctx_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "kucd_ctx_set_profile": (C.c_int, [_P, C.c_int]),
    "kucd_comm_unique_id": (C.c_int, [_P]),
    "kucd_ctx_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "kucd_rbm_create": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "kucd_rbm_destroy": (C.c_int, [_P]),
    "kucd_rbm_set_params": (C.c_int, [_P, _TP, _TP, _TP]),
    "kucd_rbm_get_params": (C.c_int, [_P, _TP, _TP, _TP]),
    "kucd_rbm_peer_export": (C.c_int, [_P, _P]),
    "kucd_rbm_peer_attach": (C.c_int, [_P, _P]),
    "kucd_rbm_peer_detach": (C.c_int, [_P]),
    "kucd_rbm_set_seed": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "kucd_rbm_get_counters": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]),
    "kucd_rbm_get_draw_counters": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "kucd_rbm_set_draw_counters": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "kucd_rbm_get_momentum": (C.c_int, [_P, _TP, _TP, _TP, C.POINTER(C.c_int)]),
    "kucd_rbm_set_momentum": (C.c_int, [_P, _TP, _TP, _TP]),
    "kucd_rbm_transform": (C.c_int, [_P, _TP, _TP, _TP, _TP]),
    "kucd_rbm_inv_transform": (C.c_int, [_P, _TP, _TP, _TP, _TP]),
    "kucd_rbm_free_energy": (C.c_int, [_P, _TP, _TP]),
    "kucd_rbm_cd_step": (C.c_int, [_P, _TP, C.POINTER(HParams), C.POINTER(Inject), C.c_int64,
                                    C.POINTER(StepStats)]),
    "kucd_rbm_score": (C.c_int, [_P, _TP, _TP, _TP, C.POINTER(C.c_float)]),
    "kucd_rbm_last_stats": (C.c_int, [_P, _TP, _TP, _TP, _TP, _TP, _TP]),
    "kucd_rbm_delta_rule": (C.c_int, [_P, C.c_int, _TP, _TP, C.c_float, C.c_int]),
    "kucd_rbm_set_chains": (C.c_int, [_P, _TP]),
    "kucd_rbm_get_chains": (C.c_int, [_P, _TP]),
    "kucd_dataset_create": (C.c_int, [_P, _TP, C.c_int, C.POINTER(_P)]),
    "kucd_dataset_destroy": (C.c_int, [_P]),
    "kucd_dataset_shape": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kucd_dataset_read": (C.c_int, [_P, _TP]),
    "kucd_dataset_shuffle": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.POINTER(_P)]),
    "kucd_rbm_fit_epoch": (C.c_int, [_P, _P, C.c_int64, C.POINTER(HParams), C.c_int64, C.POINTER(EpochStats)]),
    "kucd_rbm_fit_range": (C.c_int, [_P, _P, C.c_int64, C.POINTER(HParams), C.c_int64, C.c_int64, C.c_int64,
                                      C.POINTER(EpochStats)]),
    "kucd_rbm_fit_host": (C.c_int, [_P, _TP, C.c_int64, C.POINTER(HParams), C.c_int64, C.POINTER(C.c_float),
                                     C.POINTER(EpochStats)]),
    "kucd_tensor_from_dlpack": (C.c_int, [_P, _TP]),
    "kucd_rbm_transform_dataset": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "kucd_rbm_inv_transform_dataset": (C.c_int, [_P, _P, C.POINTER(_P)]),
}


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into keras_unsupervised_b200/libkucd.so (nvcc cross-compiles without a GPU).
    The library is several translation units - kucd.cu (host side, HBM-bound kernels) and the explicit instantiations of the
    tensor-core kernels per epilogue / chain variant (inst_*.cu) - compiled side by side and linked: about a minute on
    eight cores instead of four.  KUCD_BUILD_JOBS=1 compiles them one after the other.  Skipped when the library is newer
    than every source."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [HEADER]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    units = [SRC] + sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.startswith("inst_") and f.endswith(".cu"))
    objdir = os.path.join(os.path.dirname(_HERE), "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + ["-DKUCD_SPLIT_BUILD=1", "-c"]

    def compile_unit(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + compile_flags + ["-o", obj, src]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        t0 = time.time()
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (os.path.basename(src), res.stdout, res.stderr))
        if verbose:
            print("  %s: %.0f s" % (os.path.basename(src), time.time() - t0), file=sys.stderr)
        return obj

    jobs = int(os.environ.get("KUCD_BUILD_JOBS", "0")) or min(len(units), os.cpu_count() or 1)
    if jobs > 1:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=jobs) as pool:
            objs = list(pool.map(compile_unit, units))
    else:
        objs = [compile_unit(u) for u in units]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs + [
        "-lcuda", "-ldl"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    for o in objs:  # every unit is recompiled by the next build anyway; the tree travels to the GPU box
        os.remove(o)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load libkucd.so and attach the prototypes.  Fails loudly: no library, no engine."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "keras_unsupervised_b200 has no CPU or pure-Python path.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError when the build lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.kucd_abi_version() != 1:
        raise RuntimeError("libkucd.so ABI version mismatch")
    _lib = lib
    return lib


class KucdError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc == KUCD_OK:
        return
    msg = load().kucd_last_error().decode("utf-8", "replace")
    if rc in (ERR_INVALID_ARG, ERR_SHAPE_MISMATCH, ERR_UNSUPPORTED_DTYPE):
        raise ValueError(msg)
    raise KucdError(f"[kucd {rc}] {msg}")


# ---- arrays -> kucd_tensor ----------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _is_capsule(x) -> bool:
    return type(x).__name__ == "PyCapsule"


def tensor_of(x, keep: list) -> Tensor:
    """Describe a numpy array or a torch tensor (CPU, pinned or CUDA) as a kucd_tensor without copying
    when its layout allows it.  Objects that must outlive the call are appended to `keep`."""
    import numpy as np

    if type(x).__name__ == "PackedBits" and hasattr(x, "n_cols"):  # data.PackedBits: one bit per unit
        t = tensor_of(x.data, keep)  # the uint8 byte matrix
        rows, nbytes = t.shape[0], t.shape[1]
        pitch_bits = t.strides[0] * 8 if rows > 1 else max(nbytes, 1) * 8
        return Tensor(t.data, t.device_type, t.device_id, DT_UINT, 1, (C.c_int64 * 2)(rows, int(x.n_cols)),
                      (C.c_int64 * 2)(pitch_bits, 1))

    if _is_torch(x):
        import torch

        t = x
        if t.dim() == 1:
            t = t.reshape(-1, 1)
        if t.dim() != 2:
            raise ValueError(f"expected a 2-D array, got {tuple(x.shape)}")
        if t.dtype == torch.float32:
            code, bits = DT_FLOAT, 32
        elif t.dtype == torch.bfloat16:
            code, bits = DT_BFLOAT, 16
        elif t.dtype == torch.uint8:
            code, bits = DT_UINT, 8
        elif t.dtype == torch.bool:
            t, code, bits = t.view(torch.uint8), DT_UINT, 8
        else:
            t, code, bits = t.to(torch.float32), DT_FLOAT, 32
        if t.shape[1] > 1 and t.stride(1) != 1:
            t = t.contiguous()
        keep.append(t)
        if t.is_cuda:
            dev, dev_id = DEV_CUDA, t.device.index or 0
        else:
            dev, dev_id = (DEV_CUDA_HOST if t.is_pinned() else DEV_CPU), 0
        s0 = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
        return Tensor(t.data_ptr(), dev, dev_id, code, bits, (C.c_int64 * 2)(*t.shape), (C.c_int64 * 2)(s0, 1))

    if _is_capsule(x) or (hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray)):
        # any other DLPack producer (TensorFlow: tf.experimental.dlpack.to_dlpack(t), jax, cupy ...) or a raw "dltensor"
        # capsule: the library reads the DLManagedTensor itself (kucd_tensor_from_dlpack) - no torch in between.  The
        # capsule stays unconsumed in `keep`; when it is dropped after the call its own destructor runs the deleter.
        cap = x if _is_capsule(x) else x.__dlpack__()
        C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
        C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
        ptr = C.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")
        t = Tensor()
        check(load().kucd_tensor_from_dlpack(C.c_void_p(ptr), C.byref(t)))
        keep.append(cap)
        keep.append(x)
        return t

    a = np.asarray(x)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError(f"expected a 2-D array, got {a.shape}")
    if a.dtype == np.float32:
        code, bits = DT_FLOAT, 32
    elif a.dtype == np.uint8:
        code, bits = DT_UINT, 8
    elif a.dtype == np.bool_:
        a, code, bits = a.view(np.uint8), DT_UINT, 8
    else:
        a, code, bits = a.astype(np.float32), DT_FLOAT, 32
    if a.shape[1] > 1 and a.strides[1] != a.itemsize or any(s < 0 for s in a.strides) or \
            (a.shape[0] > 1 and a.strides[0] % a.itemsize != 0):
        a = np.ascontiguousarray(a)
    keep.append(a)
    s0 = a.strides[0] // a.itemsize if a.shape[0] > 1 else max(a.shape[1], 1)
    return Tensor(a.ctypes.data, DEV_CPU, 0, code, bits, (C.c_int64 * 2)(*a.shape), (C.c_int64 * 2)(s0, 1))
