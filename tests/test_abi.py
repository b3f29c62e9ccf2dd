"""The C-ABI library loads and exports every symbol include/kucd.h declares; the ctypes mirror of its
structs has the layout the C compiler gives them.  No compute call is made (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kucd.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kucd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from keras_unsupervised_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), "libkucd.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes prototypes and include/kucd.h disagree"
    assert lib.kucd_abi_version() == 1


def test_header_is_plain_c_and_struct_layouts_match_ctypes():
    from keras_unsupervised_b200 import _lib

    src = r'''
#include <stdio.h>
#include "kucd.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(kucd_tensor), sizeof(kucd_hparams), sizeof(kucd_inject),
         sizeof(kucd_step_stats), sizeof(kucd_epoch_stats), sizeof(kucd_timings));
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(t) for t in (_lib.Tensor, _lib.HParams, _lib.Inject, _lib.StepStats, _lib.EpochStats, _lib.Timings)]
    assert sizes == want


def test_no_gpu_means_an_error_not_a_fallback():
    """Without a visible sm_100 device the engine refuses to start; nothing is computed elsewhere."""
    from keras_unsupervised_b200 import _lib

    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.kucd_ctx_create(C.byref(h), 0, 1)
    if rc == 0:   # a GPU is present (the suite is also run on the GPU box)
        lib.kucd_ctx_destroy(h)
        pytest.skip("a GPU is visible")
    assert rc in (_lib.ERR_CUDA, _lib.ERR_NOT_SM100)
    assert b"no CPU path" in lib.kucd_last_error() or b"sm_100a only" in lib.kucd_last_error()
    with pytest.raises(RuntimeError):
        from keras_unsupervised_b200.engine import Context
        Context(device=0)
    assert lib.kucd_rbm_destroy(None) == 0 and lib.kucd_ctx_destroy(None) == 0
    assert lib.kucd_sync(None) == _lib.ERR_INVALID_ARG


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "keras_unsupervised_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "oracle/" not in text, f


def test_dlpack_capsules_are_described_without_torch_or_a_gpu():
    """kucd_tensor_from_dlpack: the library reads a producer's DLManagedTensor itself (TF's
    tf.experimental.dlpack.to_dlpack hands out the same capsule), so the TF -> DLPack -> ctypes path needs no torch.
    numpy is the producer here: data pointer, shape, element strides, dtype and device map field by field; unsupported
    layouts and dtypes are refused with the library's own message."""
    import numpy as np

    from keras_unsupervised_b200 import _lib as L

    a = np.arange(6 * 10, dtype=np.float32).reshape(6, 10)
    keep = []
    t = L.tensor_of(a.__dlpack__(), keep)                      # a raw "dltensor" capsule
    assert t.data == a.ctypes.data and (t.shape[0], t.shape[1]) == (6, 10) and (t.strides[0], t.strides[1]) == (10, 1)
    assert (t.device_type, t.dtype_code, t.bits) == (L.DEV_CPU, L.DT_FLOAT, 32)
    view = a[1:5, 2:9]                                         # row pitch 10, offset data pointer

    class Producer:                                            # anything with __dlpack__ that is neither numpy nor torch
        def __dlpack__(self):
            return view.__dlpack__()

    t = L.tensor_of(Producer(), keep)
    assert t.data == view.ctypes.data and (t.shape[0], t.shape[1]) == (4, 7) and t.strides[0] == 10
    v = np.arange(5, dtype=np.uint8)
    t = L.tensor_of(v.__dlpack__(), keep)                      # 1-D -> (n, 1)
    assert (t.shape[0], t.shape[1], t.dtype_code, t.bits) == (5, 1, L.DT_UINT, 8)
    with pytest.raises(ValueError, match="innermost stride"):
        L.tensor_of(a.T.__dlpack__(), keep)
    with pytest.raises(ValueError, match="not float32"):
        L.tensor_of(a.astype(np.float64).__dlpack__(), keep)
    with pytest.raises(ValueError, match="dimensions"):
        L.tensor_of(np.zeros((2, 2, 2), np.float32).__dlpack__(), keep)
