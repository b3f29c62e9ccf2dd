"""CPU study for DESIGN section 10 item 3: float32-grade contractions with TWO term products per weight instead of three.
W = hi + lo with hi = bf16(W) and lo = fp16((W - hi) * 2^s) / 2^s (the MMA takes bf16 and fp16 operands alike; the piecewise
accumulation already sums pieces in registers, where the lo pieces can be scaled back).  Exact accumulation here (float64):
what is measured is the REPRESENTATION error of the weights, as a relative error of sigmoid(v.W + c) against float64 -
the quantity tests/test_gpu_parity.py::test_large_contraction_accuracy_f32 bounds by 1e-5."""
import numpy as np


def bf16(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split3(W):
    hi = bf16(W)
    mid = bf16(W - hi)
    lo = bf16(W - hi - mid)
    return hi.astype(np.float64) + mid + lo


def split2(W, shift):
    hi = bf16(W)
    lo = ((W - hi).astype(np.float64) * 2.0 ** shift).astype(np.float16).astype(np.float64) / 2.0 ** shift
    return hi.astype(np.float64) + lo


def split2h(W, shift):
    """both terms fp16: hi = fp16(W) (11 significant bits, but a narrow exponent range), lo = the scaled residual"""
    hi = W.astype(np.float16).astype(np.float32)
    lo = ((W - hi).astype(np.float64) * 2.0 ** shift).astype(np.float16).astype(np.float64) / 2.0 ** shift
    return hi.astype(np.float64) + lo


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


rng = np.random.default_rng(0)
rows, H = 256, 320
print("%-34s %-7s %-12s %-12s %-12s %-12s %-12s" % ("weights", "K", "3 x bf16", "bf16+fp16", "bf16+fp16<<8", "bf16+fp16<<12", "fp16+fp16<<11"))
for K in (4096, 16384):
    for name, W in (("U(-0.05, 0.05) (initial)", rng.uniform(-0.05, 0.05, (K, H))),
                    ("N(0, 0.5 / sqrt(K/64)) (trained-like)", rng.normal(0, 0.5, (K, H)) / np.sqrt(K / 64)),
                    ("N(0, 0.5) (large)", rng.normal(0, 0.5, (K, H)))):
        W = W.astype(np.float32)
        c = rng.uniform(-0.05, 0.05, H)
        for vname, v in (("binary", (rng.random((rows, K)) < 0.5).astype(np.float64)),):
            ref = sigmoid(v @ W.astype(np.float64) + c)
            errs = []
            for Wq in (split3(W), split2(W, 0), split2(W, 8), split2(W, 12), split2h(W, 11)):
                p = sigmoid(v @ Wq + c)
                errs.append(float((np.abs(p - ref) / np.maximum(ref, 1e-30)).max()))
            print("%-34s %-7d %-12.2e %-12.2e %-12.2e %-12.2e %-12.2e" % (name[:34], K, *errs))
