"""A small run that touches every kernel family (projections in all major-ness combinations, chain kernel, dW,
update, reductions, graph replay, free energy, Gaussian epilogues) - meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os
import sys

os.environ["KUCD_CHAIN"] = "2"      # chain kernel at small sizes
os.environ["KUCD_CHAIN_DW"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from keras_unsupervised_b200 import _lib as L  # noqa: E402
from keras_unsupervised_b200.engine import Context, Dataset, Machine  # noqa: E402

rng = np.random.default_rng(0)
ctx = Context(0, 3)
for compute in (L.COMPUTE_BF16, L.COMPUTE_F32X3):
    for mode in (0, 1):
        V, H, B = 200, 136, 300
        m = Machine(ctx, V, H, mode, compute, seed=5)
        m.set_params(rng.uniform(-0.05, 0.05, (V, H)).astype(np.float32), rng.uniform(-0.05, 0.05, V).astype(np.float32),
                     rng.uniform(-0.05, 0.05, H).astype(np.float32))
        x = (rng.random((B, V)) < 0.3).astype(np.float32) if mode == 0 else rng.normal(0, 1, (B, V)).astype(np.float32)
        m.cd_step(x, Machine.hparams(lr=1e-3, k=2, want_stats=1))
        m.last_stats(B)
        h = m.transform(x)
        m.inv_transform(h, want_p=True)
        m.free_energy(x)
        m.score(x)
        ds = Dataset.from_array(ctx, x, compute)
        m.fit_epoch(ds, 128, Machine.hparams(lr=1e-3, k=1))
        m.fit_host(x, 128, Machine.hparams(lr=1e-3, k=1))
        out = m.transform_dataset(ds)
        out.numpy()
        if mode == 0:
            m.set_chains(x[:128])
            m.cd_step(x[:128], Machine.hparams(lr=1e-3, k=1, persistent=True))
        ctx.sync()
print("sanitize_case done", ctx.timings())
