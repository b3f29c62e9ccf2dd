"""Diagnostic for tests/test_gpu_api.py::test_one_epoch_fit_streamed_or_resident (one hidden bias 1.9e-6 apart on one box):
repeat the streamed / resident / streamed / resident one-epoch fits and print where the parameters differ."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keras_unsupervised_b200.engine import Context  # noqa: E402
from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI  # noqa: E402

ctx = Context(device=0, seed=0)
rng = np.random.default_rng(41)
X = (rng.random((1000, 200)) < 0.2).astype(np.float32)


def fit(stream):
    hps = {"batch_size": 128, "epochs": 1, "lr": 1e-3, "dtype": "bf16", "seed": 9}
    if stream is not None:
        hps["stream"] = stream
    r = RBM(hps, 96, name="p", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    r.fit(X, verbose=0)
    return np.array(r.rbm_weight), np.array(r.hidden_bias), np.array(r.visible_bias)


ref = fit(False)
for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    for stream in (None, False):
        got = fit(stream)
        d = [np.abs(a - b) for a, b in zip(got, ref)]
        msg = "trial %2d %-8s  max|dW| %.3e  max|dc| %.3e  max|db| %.3e" % (trial, "streamed" if stream is None else "resident",
                                                                          d[0].max(), d[1].max(), d[2].max())
        if d[1].max() > 5e-7 or d[2].max() > 5e-7 or d[0].max() > 5e-7:
            j = int(d[1].argmax())
            msg += "   c[%d]: %.9g vs %.9g; W[:, %d] max diff %.3e; #W>5e-7: %d; #b>5e-7: %d" % (
                j, got[1][j], ref[1][j], j, d[0][:, j].max(), int((d[0] > 5e-7).sum()), int((d[2] > 5e-7).sum()))
        print(msg, flush=True)
        if d[0].max() > 5e-7 or d[1].max() > 5e-7:
            sd = [a.astype(np.float64) - b.astype(np.float64) for a, b in zip(got, ref)]
            print("      c diffs beyond 5e-8:", {int(q): float("%.4g" % sd[1][q]) for q in np.argwhere(np.abs(sd[1]) > 5e-8)[:, 0]})
            cols = np.argwhere((np.abs(sd[0]) > 5e-8).any(0))[:, 0]
            for q in cols[:4]:
                rows_ = np.argwhere(np.abs(sd[0][:, q]) > 5e-8)[:, 0]
                print("      W[:, %d]: %d entries beyond 5e-8; signed diffs (first 12): %s" % (
                    q, len(rows_), [float("%.4g" % sd[0][i, q]) for i in rows_[:12]]))
            print("      W columns touched:", [int(q) for q in cols])
