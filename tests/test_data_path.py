"""CPU tests of the step before the CD path: CSV ingestion as the reference's example does
it, binarisation, bit-packed data, the shuffling permutation.  The device helpers that are plain functions
(make_feistel_key / feistel_permute / bits_to_bf16x8 / bf16x8_to_bits in csrc/aux_kernels.cuh) are compiled for the
host by nvcc (tools/data_path_host.cu) and compared with the oracle's restatement - no GPU needed."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from keras_unsupervised_b200 import _lib as L
from keras_unsupervised_b200.data import PackedBits, binarize, concat_rows, load_csv
from keras_unsupervised_b200.parallel import shard_rows
from oracle import cd_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the reference's loader, restated row by row (examples/rbm/rbm_softmax_mnist.py:96-111, 129-139) --------------
def _reference_load_training_data(path):
    import pandas as pd

    train_df = pd.read_csv(path)
    V, gt = [], []
    for i in range(train_df.shape[0]):
        V.append(train_df.iloc[i, 1:].values / 255)
        t_gt = np.zeros(shape=(10,))
        t_gt[train_df.iloc[i, 0]] = 1.
        gt.append(t_gt)
    return np.asarray(V, dtype=np.float32), np.asarray(gt, dtype=np.float32)


def _reference_load_test_data(path):
    import pandas as pd

    test_df = pd.read_csv(path)
    V = [test_df.iloc[i, :].values / 255 for i in range(test_df.shape[0])]
    return np.asarray(V, dtype=np.float32)


def test_load_csv_equals_the_reference_loader(tmp_path):
    rng = np.random.default_rng(3)
    pix = rng.integers(0, 256, (57, 784))
    lab = rng.integers(0, 10, 57)
    lab[0] = 9
    header = ",".join(["label"] + ["pixel%d" % i for i in range(784)])
    train = tmp_path / "train.csv"
    np.savetxt(train, np.column_stack([lab, pix]), fmt="%d", delimiter=",", header=header, comments="")
    test = tmp_path / "test.csv"
    np.savetxt(test, pix, fmt="%d", delimiter=",", header=header.split(",", 1)[1], comments="")
    V_ref, gt_ref = _reference_load_training_data(train)
    V, gt = load_csv(train, label_col=0, n_classes=10)
    assert V.dtype == np.float32 and gt.dtype == np.float32
    assert np.array_equal(V, V_ref) and np.array_equal(gt, gt_ref)
    Vt, none = load_csv(test)
    assert none is None and np.array_equal(Vt, _reference_load_test_data(test))
    with pytest.raises(ValueError):
        load_csv(train, label_col=0, n_classes=5)


def test_binarize():
    x = np.array([[0.0, 0.5, 0.50001, 1.0]], np.float32)
    assert binarize(x).tolist() == [[False, False, True, True]]          # strict >
    rng = np.random.default_rng(0)
    p = np.full((200, 500), 0.1307, np.float32)
    s = binarize(p, "sample", rng)
    assert s.dtype == np.bool_ and abs(s.mean() - 0.1307) < 0.005
    assert not binarize(np.zeros((4, 4)), "sample").any() and binarize(np.ones((4, 4)), "sample").all()


@pytest.mark.parametrize("cols", [1, 7, 8, 9, 64, 130, 784])
def test_packed_bits_round_trip(cols):
    rng = np.random.default_rng(cols)
    x = (rng.random((37, cols)) < 0.4).astype(np.float32)
    p = PackedBits.from_dense(x)
    assert p.shape == (37, cols) and p.data.shape == (37, (cols + 7) // 8) and len(p) == 37
    assert np.array_equal(p.data, O.pack_bits(x))                          # oracle layout: bit j % 8 of byte j // 8
    assert np.array_equal(p.to_dense(), x) and np.array_equal(O.unpack_bits(p.data, cols), x)
    assert np.array_equal(p[5:11].to_dense(), x[5:11]) and p[3].shape == (1, cols)
    assert np.array_equal(concat_rows([p[:10], p[10:]]).to_dense(), x)
    assert p.nbytes * 32 >= x.nbytes >= p.nbytes * 4


def test_packed_bits_refuses_what_it_cannot_hold():
    with pytest.raises(ValueError):
        PackedBits.from_dense(np.array([[0.0, 0.3]]))
    with pytest.raises(ValueError):
        PackedBits(np.zeros((2, 3), np.uint8), 25)
    with pytest.raises(ValueError):
        PackedBits(np.zeros((2, 3), np.uint8), 16)
    with pytest.raises(ValueError):
        PackedBits(np.zeros((2, 3), np.float32), 24)
    with pytest.raises(IndexError):
        PackedBits(np.zeros((2, 3), np.uint8), 24)[0, 1]


def test_packed_tensor_descriptor():
    """include/kucd.h: dtype_code = UINT, bits = 1, shape[1] = columns, strides[0] = row pitch in bits."""
    p = PackedBits.from_dense(np.ones((5, 130), np.float32))
    keep = []
    t = L.tensor_of(p, keep)
    assert (t.dtype_code, t.bits) == (L.DT_UINT, 1)
    assert tuple(t.shape) == (5, 130) and tuple(t.strides) == (17 * 8, 1)
    assert t.data == p.data.ctypes.data and t.device_type == L.DEV_CPU
    wide = np.zeros((5, 40), np.uint8)
    view = PackedBits(wide[:, :17], 130)                                   # rows 40 bytes apart
    t = L.tensor_of(view, keep)
    assert tuple(t.strides) == (40 * 8, 1)
    one = L.tensor_of(p[2], keep)
    assert tuple(one.shape) == (1, 130)


def test_sharding_packed_rows_equals_packing_sharded_rows():
    rng = np.random.default_rng(9)
    x = (rng.random((1000, 77)) < 0.5).astype(np.float32)                   # 3 full minibatches of 256 + 232
    p = PackedBits.from_dense(x)
    for rank in range(4):
        dense, b, row0 = shard_rows(x, 256, rank, 4)
        packed, b2, row02 = shard_rows(p, 256, rank, 4)
        assert (b, row0) == (b2, row02) == (64, 64 * rank)
        assert isinstance(packed, PackedBits) and np.array_equal(packed.to_dense(), dense)


# ---- shuffling permutation -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 17, 128, 1000, 4097, 60000])
def test_feistel_permutation_is_a_bijection(n):
    p = O.feistel_permutation(n, seed=42, epoch=3)
    assert p.shape == (n,) and np.array_equal(np.sort(p), np.arange(n))


def test_feistel_permutation_mixes_and_depends_on_its_key():
    n = 60000
    a, b, c = (O.feistel_permutation(n, 42, 0), O.feistel_permutation(n, 42, 1), O.feistel_permutation(n, 43, 0))
    assert np.array_equal(a, O.feistel_permutation(n, 42, 0))
    for p in (a, b, c):
        assert (p == np.arange(n)).mean() < 1e-3                             # ~1/n fixed points expected
        assert abs(np.corrcoef(p, np.arange(n))[0, 1]) < 0.02                # no trend left
        d = np.diff(p)
        assert (np.abs(d) == 1).mean() < 1e-3                                # neighbours are torn apart
    assert (a == b).mean() < 1e-3 and (a == c).mean() < 1e-3
    # every minibatch of 128 draws from the whole data set: its mean source index is near n / 2
    means = a[: n // 128 * 128].reshape(-1, 128).mean(axis=1)
    assert abs(means.mean() - n / 2) < 0.02 * n and means.std() < 0.05 * n
    assert O.feistel_keys(42, 0) == [0xe70fdbbb, 0x2abed76b, 0x15b1d043, 0xe77bb78c, 0x58becc7c, 0x04a83e35]


@pytest.fixture(scope="module")
def host_tool():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc is needed to compile the host-side check of csrc/aux_kernels.cuh")
    out = os.path.join(ROOT, "build", "data_path_host")
    src = os.path.join(ROOT, "tools", "data_path_host.cu")
    dep = os.path.join(ROOT, "keras_unsupervised_b200", "csrc", "aux_kernels.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.run(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, src],
                       check=True)
    return out


@pytest.mark.parametrize("n,seed,epoch", [(1, 0, 0), (10, 42, 0), (1000, 42, 7), (60000, 2**40 + 5, 2**33 + 1)])
def test_engine_permutation_equals_the_oracle(host_tool, n, seed, epoch):
    """The functions permute_rows_kernel runs (compiled for the host) against oracle/cd_oracle.py:feistel_permutation."""
    txt = subprocess.run([host_tool, "perm", str(n), str(seed), str(epoch)], check=True, capture_output=True, text=True)
    got = np.array(txt.stdout.split(), dtype=np.int64)
    assert np.array_equal(got, O.feistel_permutation(n, seed, epoch))


def test_engine_bit_expansion_equals_the_oracle(host_tool):
    """bits_to_bf16x8 / bf16x8_to_bits (what ingest_bits_kernel / export_bits_kernel run per byte) for every byte value
    and every tail length, against the oracle's unpack_bits + bf16 encoding of 0 and 1."""
    txt = subprocess.run([host_tool, "bits"], check=True, capture_output=True, text=True)
    rows = np.array(txt.stdout.split(), dtype=np.int64).reshape(-1, 7)
    assert rows.shape[0] == 256 * 9
    b, valid, words, back = rows[:, 0], rows[:, 1], rows[:, 2:6], rows[:, 6]
    units = O.unpack_bits(b.astype(np.uint8)[:, None], 8)                   # (n, 8) floats in {0, 1}
    units = units * (np.arange(8)[None, :] < valid[:, None])
    enc = (O.bf16_round(units.astype(np.float32)).view(np.uint32) >> 16).astype(np.int64)   # bf16 bit patterns
    expect = enc[:, 0::2] | (enc[:, 1::2] << 16)
    assert np.array_equal(words, expect)
    assert np.array_equal(back, (b & ((1 << valid) - 1)))


def test_formats_match_the_committed_fixture():
    """tests/golden/data_path.json (made by tests/golden/make_data_path_fixtures.py): the packed layout and the
    shuffling stream must not drift - checkpoints and packed data sets written by one round are read by the next."""
    import json

    with open(os.path.join(ROOT, "tests", "golden", "data_path.json")) as f:
        gold = json.load(f)
    dense = np.array(gold["pack"]["dense"], dtype=np.uint8)
    assert O.pack_bits(dense).tolist() == gold["pack"]["packed"]
    assert PackedBits.from_dense(dense).data.tolist() == gold["pack"]["packed"]
    for k in gold["keys"]:
        assert O.feistel_keys(k["seed"], k["epoch"]) == k["keys"]
    for p in gold["perm"]:
        perm = O.feistel_permutation(p["rows"], p["seed"], p["epoch"])
        assert perm[:16].tolist() == p["head"]
        assert int((perm * (np.arange(p["rows"]) + 1)).sum() % (2**61 - 1)) == p["checksum"]
