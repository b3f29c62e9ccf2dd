"""Host-side logic that needs no GPU: the reference-facing classes' argument handling, the DBN stack
rules (dbn.py:14-32,47-48) and the data-parallel row layout."""
import numpy as np
import pytest

from keras_unsupervised_b200.ebm import DBN, RBM, MODE_VISIBLE_BERNOULLI, MODE_VISIBLE_GAUSSIAN, MODE_COMPLEX
from keras_unsupervised_b200.parallel import local_batch, shard_rows


def test_constants_and_constructor_signature():
    assert (MODE_VISIBLE_BERNOULLI, MODE_VISIBLE_GAUSSIAN, MODE_COMPLEX) == (0, 1, 2)     # rbm.py:14-16
    hps = {"batch_size": 128, "epochs": 1, "lr": 1e-3}
    rbm = RBM(hps, 500, name="rbm")                                                      # rbm.py:22
    assert rbm.mode == MODE_VISIBLE_GAUSSIAN and rbm.hps is hps and rbm.output_dim == 500 and rbm.name == "rbm"
    assert rbm.compute_output_shape((None, 784)) == (None, 500)                           # rbm.py:94-95
    cfg = rbm.get_config()
    assert cfg == {"hps": hps, "output_dim": 500, "name": "rbm", "mode": MODE_VISIBLE_GAUSSIAN}
    with pytest.raises(ValueError):
        RBM(hps, 10, mode=MODE_COMPLEX)
    assert not rbm.built
    with pytest.raises(ValueError):
        rbm.inv_transform(np.zeros((2, 500), np.float32))


def test_dbn_stack_rules():
    dbn = DBN()
    for call in (dbn.fit, dbn.transform, dbn.inv_transform):
        with pytest.raises(ValueError, match="Any rbm layer doesn't exist."):                # dbn.py:47-48
            call(np.zeros((2, 4), np.float32))
    hps = {"batch_size": 2, "epochs": 1, "lr": 1e-3}

    class Built:  # a layer whose dimensions are known, without touching the engine
        def __init__(self, i, o):
            self.input_shape, self.output_shape, self.output_dim, self.name = (None, i), (None, o), o, "l"

    dbn.add_stack(Built(8, 6))
    with pytest.raises(ValueError, match="output dimension"):                                  # dbn.py:27-30
        dbn.add_stack(Built(7, 3))
    dbn.add_stack(Built(6, 3))
    dbn.add_stack(RBM(hps, 2, name="unbuilt"))        # dimension unknown until data arrives: accepted
    assert len(dbn._rbm_layers) == 3


def test_shard_rows_layout():
    V = np.arange(10 * 3).reshape(10, 3)
    assert shard_rows(V, 4, 0, 1)[0] is V
    # batch 4 over 2 ranks, 10 rows: minibatches [0:4], [4:8], remainder [8:10]
    l0, b0, g0 = shard_rows(V, 4, 0, 2)
    l1, b1, g1 = shard_rows(V, 4, 1, 2)
    assert (b0, g0, b1, g1) == (2, 0, 2, 2)
    assert l0[:, 0].tolist() == [0, 3, 12, 15, 24] and l1[:, 0].tolist() == [6, 9, 18, 21, 27]
    # every row is owned exactly once
    assert sorted(np.concatenate([l0, l1])[:, 0].tolist()) == V[:, 0].tolist()
    with pytest.raises(ValueError):
        shard_rows(V, 5, 0, 2)
    with pytest.raises(ValueError):
        shard_rows(np.zeros((9, 3)), 4, 0, 2)     # remainder of 1 row does not split over 2 ranks
    assert local_batch(4096, 8192, 8) == 512


def test_tensor_descriptors_for_numpy_and_torch_inputs():
    """kucd_tensor = flattened DLTensor: what the shim hands to the C ABI for the array kinds callers pass."""
    import torch

    from keras_unsupervised_b200 import _lib as L

    keep = []
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    t = L.tensor_of(a, keep)
    assert (t.device_type, t.dtype_code, t.bits) == (L.DEV_CPU, L.DT_FLOAT, 32)
    assert tuple(t.shape) == (3, 4) and tuple(t.strides) == (4, 1) and t.data == a.ctypes.data   # no copy
    sl = L.tensor_of(a[:, :2], keep)                     # row-strided view: still no copy
    assert tuple(sl.shape) == (3, 2) and tuple(sl.strides) == (4, 1) and sl.data == a.ctypes.data
    tr = L.tensor_of(a.T, keep)                          # innermost stride != 1: copied
    assert tuple(tr.shape) == (4, 3) and tuple(tr.strides) == (3, 1) and tr.data != a.ctypes.data
    v = L.tensor_of(np.zeros(5, np.float32), keep)       # vectors travel as (n, 1) with unit stride
    assert tuple(v.shape) == (5, 1) and tuple(v.strides) == (1, 1)
    u8 = L.tensor_of(np.ones((2, 3), np.uint8), keep)
    assert (u8.dtype_code, u8.bits) == (L.DT_UINT, 8)
    bl = L.tensor_of(np.ones((2, 3), np.bool_), keep)
    assert (bl.dtype_code, bl.bits) == (L.DT_UINT, 8)
    f64 = L.tensor_of(np.ones((2, 3), np.float64), keep) # anything else is converted to float32 (K.floatx())
    assert (f64.dtype_code, f64.bits) == (L.DT_FLOAT, 32)
    with pytest.raises(ValueError):
        L.tensor_of(np.zeros((2, 2, 2), np.float32), keep)
    tt = torch.arange(6, dtype=torch.bfloat16).reshape(2, 3)
    tb = L.tensor_of(tt, keep)
    assert (tb.device_type, tb.dtype_code, tb.bits) == (L.DEV_CPU, L.DT_BFLOAT, 16) and tb.data == tt.data_ptr()
    ti = L.tensor_of(torch.ones(2, 3, dtype=torch.int64), keep)
    assert (ti.dtype_code, ti.bits) == (L.DT_FLOAT, 32)


def test_error_codes_map_to_python_exceptions():
    from keras_unsupervised_b200 import _lib as L

    lib = L.load()
    assert lib.kucd_sync(None) == L.ERR_INVALID_ARG
    with pytest.raises(ValueError, match="NULL"):
        L.check(lib.kucd_sync(None))
    L.check(0)


def test_config_round_trip():
    hps = {"batch_size": 64, "epochs": 3, "lr": 0.01, "k": 2}
    rbm = RBM(hps, 32, name="x", mode=MODE_VISIBLE_BERNOULLI)
    twin = RBM.from_config(rbm.get_config())
    assert twin.get_config() == rbm.get_config() and twin.mode == MODE_VISIBLE_BERNOULLI and not twin.built
    legacy = RBM.from_config({"hps": hps, "output_dim": 32, "name": "y"})   # the reference's config has no mode
    assert legacy.mode == MODE_VISIBLE_GAUSSIAN                              # rbm.py:22 default


def test_reference_import_lines_resolve_to_the_engine():
    """install_as_ku(): `from ku.ebm.rbm import RBM` (examples/rbm/rbm_softmax_mnist.py:24) and `from ku.ebm import
    RBM, DBN` (ku/ebm/__init__.py:1-2) work unchanged; run in a subprocess so the alias does not leak."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import keras_unsupervised_b200 as kucd\n"
        "kucd.install_as_ku()\n"
        "from ku.ebm.rbm import RBM\n"
        "from ku.ebm import RBM as R2, DBN\n"
        "from ku.ebm.rbm import MODE_VISIBLE_BERNOULLI, MODE_VISIBLE_GAUSSIAN, MODE_COMPLEX\n"
        "from ku.ebm.dbn import DBN as D2\n"
        "import ku.ebm\n"
        "assert RBM is R2 is kucd.ebm.RBM and DBN is D2 is kucd.ebm.DBN\n"
        "r = RBM({'batch_size': 128, 'epochs': 1, 'lr': 0.001}, 128, name='rbm')\n"
        "assert r.mode == MODE_VISIBLE_GAUSSIAN and (MODE_VISIBLE_BERNOULLI, MODE_COMPLEX) == (0, 2)\n"
        "try:\n"
        "    import ku.backprop\n"
        "    raise SystemExit('out-of-scope subpackage resolved')\n"
        "except ImportError:\n"
        "    pass\n"
        "kucd.install_as_ku()\n"
        "print('ok')\n")
    res = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, PYTHONPATH=root))
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout + res.stderr
