"""Float32-grade mode: how long may a tensor-memory accumulation piece be?  (launch.cuh: kPreciseCH k-blocks of 64.)
Usage: python tools/diag/ch_experiment.py <path to a libkucd.so built with -DKUCD_PRECISE_CH=n>
Prints the accuracy of the K = 4096 contraction against the float64 oracle (the bar of
tests/test_gpu_parity.py::test_large_contraction_accuracy_f32 is 1e-5 relative) and the time of a C3 float32-grade step."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from keras_unsupervised_b200 import _lib as L  # noqa: E402

L.LIB_PATH = os.path.abspath(sys.argv[1])
from keras_unsupervised_b200.engine import Context, Dataset, Machine  # noqa: E402
from oracle import cd_oracle as O  # noqa: E402  (a diagnostic, not the product)

ctx = Context(device=0, seed=0)
rng = np.random.default_rng(83)
for V, H in ((4096, 320), (16384, 320)):
    rows = 256
    m = Machine(ctx, V, H, 0, L.COMPUTE_F32X3, seed=0)
    W, b, c = O.OracleRBM.init_params(V, H, seed=0)
    m.set_params(W, b, c)
    orc = O.OracleRBM(W, b, c, compute="f64")
    v = rng.random((rows, V)).astype(np.float32)
    p = m.transform(v, want_p=True)[1]
    ref = orc.prob_h(v)
    vb = (v < 0.5).astype(np.float32)
    pb = m.transform(vb, want_p=True)[1]
    refb = orc.prob_h(vb)
    # trained-like weights (sigma 0.5): larger pre-activations, the harder case for the aligned accumulation
    W2 = (rng.normal(0, 0.5, (V, H)) / np.sqrt(V / 64)).astype(np.float32)
    m.set_params(W2, b, c)
    orc2 = O.OracleRBM(W2, b, c, compute="f64")
    p2 = m.transform(vb, want_p=True)[1]
    ref2 = orc2.prob_h(vb)
    rel = lambda a, r: float((np.abs(a - r) / np.maximum(np.abs(r), 1e-30)).max())
    print("%s K=%d: max rel err real-valued %.3e, binary %.3e, binary with larger weights %.3e" % (
        os.path.basename(sys.argv[1]), V, rel(p, ref), rel(pb, refb), rel(p2, ref2)), flush=True)
    m.close()

# C3 float32-grade step: 4096 -> 4096, CD-10, 4096 rows, graph replay over a resident data set
V = H = B = 4096
m = Machine(ctx, V, H, 0, L.COMPUTE_F32X3, seed=0)
W, b, c = O.OracleRBM.init_params(V, H, seed=0)
m.set_params(W, b, c)
data = (rng.random((B * 4, V)) < 0.5).astype(np.float32)
ds = Dataset.from_array(ctx, data, L.COMPUTE_F32X3)
hp = Machine.hparams(lr=1e-3, k=10)
m.fit_epoch(ds, B, hp, want_stats=False)
ctx.sync()
t0 = time.perf_counter()
for _ in range(3):
    m.fit_epoch(ds, B, hp, want_stats=False)
ctx.sync()
dt = (time.perf_counter() - t0) / 12
print("%s C3 float32-grade: %.3f ms per step, %.3f M samples/s" % (os.path.basename(sys.argv[1]), dt * 1e3, B / dt / 1e6), flush=True)
