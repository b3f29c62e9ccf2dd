"""The step before the CD path: getting a data matrix to the engine.

The reference's only ingestion code is the example's Kaggle-MNIST loader
(/root/reference/examples/rbm/rbm_softmax_mnist.py:96-111 train.csv -> V/255 + one-hot labels;
:129-139 test.csv -> V/255): a header row, the label in column 0 of train.csv, 784 pixel columns, a
Python loop over DataFrame rows.  `load_csv` reads the same files in one vectorised pass and returns the
same arrays; `binarize` and `PackedBits` are what a Bernoulli RBM wants to be fed with: 0/1 data stored one
bit per unit, which crosses PCIe and HBM 32x smaller than float32 and is expanded to the engine's bf16
operand planes by one HBM-bound kernel (csrc/aux_kernels.cuh:ingest_bits_kernel).

Host-side preparation only - nothing of the CD arithmetic happens here.
"""
from __future__ import annotations

import numpy as np


class PackedBits:
    """A (rows, n_cols) 0/1 matrix stored as bits: column j of a row is bit (j % 8) of byte (j // 8), i.e.
    numpy.packbits(x, axis=1, bitorder="little").  Accepted by every engine entry point that reads a visible
    or hidden matrix (include/kucd.h: dtype_code = KUCD_DT_UINT, bits = 1)."""

    def __init__(self, data, n_cols: int):
        if type(data).__module__.split(".")[0] == "torch":
            if data.dim() != 2 or str(data.dtype) != "torch.uint8":
                raise ValueError("PackedBits needs a 2-D uint8 tensor")
            if data.shape[1] > 1 and data.stride(1) != 1:
                data = data.contiguous()
        else:
            data = np.asarray(data)
            if data.ndim != 2 or data.dtype != np.uint8:
                raise ValueError("PackedBits needs a 2-D uint8 array")
            if data.shape[1] > 1 and data.strides[1] != 1:
                data = np.ascontiguousarray(data)
        n_cols = int(n_cols)
        if not (data.shape[1] - 1) * 8 < n_cols <= data.shape[1] * 8 and not (n_cols == 0 and data.shape[1] == 0):
            raise ValueError(f"{data.shape[1]} bytes per row cannot hold exactly {n_cols} columns")
        self.data = data
        self.n_cols = n_cols

    @property
    def shape(self):
        return (int(self.data.shape[0]), self.n_cols)

    @property
    def nbytes(self) -> int:
        return int(self.data.shape[0]) * int(self.data.shape[1])

    def __len__(self):
        return int(self.data.shape[0])

    def __getitem__(self, rows):
        """Row slicing only (what the minibatch loop and the data-parallel sharding need)."""
        if isinstance(rows, tuple):
            raise IndexError("PackedBits supports row indexing only")
        sub = self.data[rows]
        if sub.ndim == 1:
            sub = sub.reshape(1, -1)
        return PackedBits(sub, self.n_cols)

    @classmethod
    def from_dense(cls, x) -> "PackedBits":
        """Pack a 0/1 matrix (any numeric or boolean dtype).  Values other than 0 and 1 are refused rather
        than thresholded: binarise explicitly first."""
        a = np.asarray(x)
        if a.ndim != 2:
            raise ValueError(f"expected a 2-D array, got {a.shape}")
        if a.dtype != np.bool_:
            if not np.all((a == 0) | (a == 1)):
                raise ValueError("PackedBits.from_dense needs 0/1 data; call binarize() first")
            a = a != 0
        return cls(np.packbits(a, axis=1, bitorder="little"), a.shape[1])

    def to_dense(self, dtype=np.float32) -> np.ndarray:
        data = self.data.cpu().numpy() if type(self.data).__module__.split(".")[0] == "torch" else self.data
        if data.shape[0] == 0:
            return np.zeros((0, self.n_cols), dtype)
        return np.unpackbits(data, axis=1, count=self.n_cols, bitorder="little").astype(dtype)

    def pin(self) -> "PackedBits":
        """Page-locked copy (a torch uint8 tensor) for asynchronous host -> device streaming."""
        import torch

        t = self.data if type(self.data).__module__.split(".")[0] == "torch" else torch.from_numpy(self.data)
        return PackedBits(t.pin_memory(), self.n_cols)


def concat_rows(parts):
    """Row-wise concatenation of PackedBits (or plain arrays)."""
    if all(isinstance(p, PackedBits) for p in parts):
        datas = [p.data for p in parts]
        if type(datas[0]).__module__.split(".")[0] == "torch":
            import torch

            return PackedBits(torch.cat(datas), parts[0].n_cols)
        return PackedBits(np.concatenate(datas), parts[0].n_cols)
    return np.concatenate(parts)


def binarize(x, threshold=0.5, rng=None) -> np.ndarray:
    """Real-valued intensities in [0, 1] -> 0/1 (bool).  threshold='sample' draws each unit as a Bernoulli
    variable with its intensity as the probability (the usual binarised-MNIST recipe); a number thresholds
    with a strict >."""
    a = np.asarray(x, dtype=np.float32)
    if isinstance(threshold, str):
        if threshold != "sample":
            raise ValueError("threshold must be a number or 'sample'")
        rng = rng if rng is not None else np.random.default_rng(0)
        return rng.random(a.shape, dtype=np.float32) < a
    return a > np.float32(threshold)


def load_csv(path, label_col=None, divide_by=255.0, n_classes=None, header=True, dtype=np.float32):
    """A numeric CSV -> (V, gt).

    Mirrors rbm_softmax_mnist.py:96-111 / :129-139: `label_col=0` for Kaggle's train.csv (V = the remaining
    columns / divide_by, gt = one-hot of column 0, float32), `label_col=None` for test.csv (gt is None).  The
    division is done in float64 and then rounded to float32, as `values/255` followed by
    `np.asarray(..., dtype=np.float32)` does there; `n_classes` defaults to max label + 1 (the reference
    hard-codes 10).
    """
    try:
        import pandas as pd

        df = pd.read_csv(path, header=0 if header else None)
        raw = df.to_numpy()
    except ImportError:  # pandas is what the reference uses; numpy reads the same files
        raw = np.loadtxt(path, delimiter=",", skiprows=1 if header else 0, ndmin=2)
    if raw.ndim != 2 or raw.shape[1] == 0:
        raise ValueError(f"{path}: not a numeric table")
    gt = None
    if label_col is not None:
        labels = raw[:, label_col].astype(np.int64)
        raw = np.delete(raw, label_col, axis=1)
        if labels.size and labels.min() < 0:
            raise ValueError("labels must be non-negative integers")
        k = int(n_classes) if n_classes is not None else (int(labels.max()) + 1 if labels.size else 0)
        if labels.size and labels.max() >= k:
            raise ValueError(f"label {int(labels.max())} does not fit {k} classes")
        gt = np.zeros((labels.shape[0], k), dtype=dtype)
        gt[np.arange(labels.shape[0]), labels] = 1
    V = np.asarray(raw, dtype=np.float64) / divide_by
    return np.ascontiguousarray(V, dtype=dtype), gt
