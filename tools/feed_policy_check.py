"""Where the time of a one-epoch fit of a small-minibatch host array goes (C1 shape: 60000 x 784 float32, batch 128):
streamed (kucd_rbm_fit_host) against resident (kucd_dataset_create + graph replay), phase by phase."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from keras_unsupervised_b200 import _lib as L  # noqa: E402
from keras_unsupervised_b200.engine import Context, Dataset, Machine  # noqa: E402

ctx = Context(device=0, seed=1)
N, V, H, B = 60000, 784, 500, 128
rng = np.random.default_rng(0)
X = (rng.random((N, V)) < 0.13).astype(np.float32)
Xp = torch.from_numpy(X).pin_memory()
m = Machine(ctx, V, H, L.MODE_VISIBLE_BERNOULLI, L.COMPUTE_BF16, seed=5)
m.set_params(rng.uniform(-0.05, 0.05, (V, H)).astype(np.float32), np.zeros(V, np.float32), np.zeros(H, np.float32))
hp = Machine.hparams(lr=1e-3, k=1, normalize=True)
res = {}


def timed(name, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        ctx.sync()
        t0 = time.perf_counter()
        out = fn()
        ctx.sync()
        best = min(best, time.perf_counter() - t0)
    res[name] = round(1e3 * best, 3)
    return out


for label, arr in (("pageable", X), ("pinned", Xp)):
    timed(f"fit_host_{label}_ms", lambda: m.fit_host(arr, B, hp))
    ds = timed(f"dataset_create_{label}_ms", lambda: Dataset.from_array(ctx, arr, L.COMPUTE_BF16), reps=1)
    timed(f"fit_epoch_first_{label}_ms", lambda: m.fit_epoch(ds, B, hp, want_stats=False), reps=1)
    timed(f"fit_epoch_again_{label}_ms", lambda: m.fit_epoch(ds, B, hp, want_stats=False))
    hp.want_stats = 1
    timed(f"fit_epoch_stats_{label}_ms", lambda: m.fit_epoch(ds, B, hp, want_stats=True))
    hp.want_stats = 0
    timed(f"dataset_close_{label}_ms", lambda: ds.close(), reps=1)
    ds2 = timed(f"dataset_create_2nd_{label}_ms", lambda: Dataset.from_array(ctx, arr, L.COMPUTE_BF16), reps=1)
    ds2.close()
print(json.dumps(res))
