"""Generate golden vectors by EXECUTING the reference's own ku/ebm/rbm.py and ku/ebm/dbn.py.

Run in the build container (needs /root/reference; TensorFlow is absent, so the keras-backend calls are
served by tests/golden/kshim.py):

    python tests/golden/make_reference_fixtures.py

Writes tests/golden/ref_rbm_{bernoulli,gaussian}.npz (64 -> 64, batch 16, 4 minibatches), ref_rbm_{bernoulli,gaussian}_128.npz
(128 -> 128, batch 32, 8 minibatches, other seeds and learning rate) and ref_dbn.json.  The tests that consume
them (tests/test_oracle_vs_reference.py) run anywhere: they never read /root/reference.

What can be executed of the reference, and how:
  * RBM.build() as written.  rbm.py:46 draws the hidden uniforms with shape (batch_size, n_visible) and
    compares them with a (rows, n_hidden) tensor (SURVEY.md defect D3), so the fixtures use
    n_visible == n_hidden and rows == batch_size, where the line is well-formed.
  * RBM.fit() as written, up to the exception it always raises on the last minibatch of the first epoch
    (rbm.py:169: int(x, base); defect D1).  Every earlier minibatch runs the full reference schedule:
    rbm_weight_update_func, hidden_bias_update_func, visible_bias_update_func, cal_free_energy,
    sample_first_visible, cal_free_energy (rbm.py:221-233).  The shim records each call's random draws
    and outputs.
  * transform_func / inv_transform_func / free_energy_func directly (the `transform` / `inv_transform`
    METHODS are shadowed by tensors of the same name after build(), defect D5).
  * DBN: dbn.py is import-free; its add_stack / fit / transform / inv_transform are driven with recording
    mock layers to document the defects D7 that the new DBN resolves.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import kshim  # noqa: E402


# the second size widens the pin: a larger layer, a larger minibatch, twice as many minibatches, other seeds and rate
SIZES = {"": dict(V=64, B=16, N=80, lr=0.01, seed=7, param_seed=0, data_seed=1234),
         "_128": dict(V=128, B=32, N=288, lr=0.003, seed=11, param_seed=5, data_seed=4321)}


def run_rbm(mode_name, out_dir=HERE, size=""):
    sz = SIZES[size]
    kshim.REC = kshim.Recorder(seed=sz["seed"], param_seed=sz["param_seed"])
    kshim.Function._count = 0
    kshim.install("/root/reference")
    for name in [n for n in sys.modules if n.startswith("ku.ebm.")]:
        del sys.modules[name]
    import importlib

    rbm_mod = importlib.import_module("ku.ebm.rbm")
    assert rbm_mod.__file__ == "/root/reference/ku/ebm/rbm.py", rbm_mod.__file__
    mode = rbm_mod.MODE_VISIBLE_BERNOULLI if mode_name == "bernoulli" else rbm_mod.MODE_VISIBLE_GAUSSIAN
    V = H = sz["V"]
    B, N = sz["B"], sz["N"]
    hps = {"batch_size": B, "epochs": 2, "lr": sz["lr"]}
    rbm = rbm_mod.RBM(hps, H, name="rbm", mode=mode)
    rbm.build((None, V))
    W0, c0, b0 = rbm.rbm_weight.value.copy(), rbm.hidden_bias.value.copy(), rbm.visible_bias.value.copy()

    data_rng = np.random.default_rng(sz["data_seed"])
    if mode_name == "bernoulli":
        X = (data_rng.random((N, V)) < 0.3).astype(np.float32)
    else:
        X = data_rng.standard_normal((N, V)).astype(np.float32)

    # --- inference functions before training (rbm.py:48,54,76) ---
    REC = kshim.REC
    n0 = len(REC.calls)
    h = rbm.transform_func([X[:B]])[0]
    Hin = (data_rng.random((B, H)) < 0.5).astype(np.float32)
    v = rbm.inv_transform_func([Hin])[0]
    fe = rbm.free_energy_func([X[:B]])[0]
    infer = dict(x=X[:B], h=h, u_h=REC.calls[n0]["draws"][0][2], h_in=Hin, v=v, u_v=REC.calls[n0 + 1]["draws"][0][2],
                 fe=fe)

    # --- fit (rbm.py:100-234) until the reference's own exception ---
    n1 = len(REC.calls)
    out = io.StringIO()
    err = None
    with contextlib.redirect_stdout(out):
        try:
            rbm.fit(X, verbose=1)
        except Exception as e:  # noqa: BLE001  (D1: TypeError from int(x, base))
            err = "%s: %s" % (type(e).__name__, e)
    calls = REC.calls[n1:]
    assert err is not None and "int()" in err, err
    assert len(calls) % 6 == 0
    steps = len(calls) // 6
    assert steps == N // B - 1, steps
    # the hidden draw is the random node created in build(); the visible draw the one created in fit()
    transform_rand_id = infer_id = REC.calls[n0]["draws"][0][0]
    arrays = {}
    scores = []
    for s in range(steps):
        group = calls[6 * s:6 * s + 6]
        for j, tag in zip((0, 1, 2, 4), "ABCD"):   # runs A, B, C (updates) and D (sample_first_visible)
            draws = {nid: arr for nid, _, arr in group[j]["draws"]}
            assert len(draws) == 2 and transform_rand_id in draws
            arrays["s%d_%s_uh" % (s, tag)] = draws[transform_rand_id]
            arrays["s%d_%s_uv" % (s, tag)] = [a for nid, a in draws.items() if nid != transform_rand_id][0]
        arrays["s%d_W_after_A" % s] = group[0]["outputs"][0]
        arrays["s%d_c_after_B" % s] = group[1]["outputs"][0]
        arrays["s%d_b_after_C" % s] = group[2]["outputs"][0]
        arrays["s%d_fe" % s] = group[3]["outputs"][0]
        arrays["s%d_v_neg_D" % s] = group[4]["outputs"][0]
        arrays["s%d_fe_p" % s] = group[5]["outputs"][0]
        scores.append(float(np.mean(np.abs(group[3]["outputs"][0] - group[5]["outputs"][0]))))  # rbm.py:233
    printed = [float(l.split("score:")[1]) for l in out.getvalue().splitlines() if "score:" in l]
    assert len(printed) == steps and np.allclose(printed, scores, atol=1e-6)
    np.savez_compressed(
        os.path.join(out_dir, "ref_rbm_%s%s.npz" % (mode_name, size)), X=X, W0=W0, b0=b0, c0=c0, lr=np.float32(hps["lr"]),
        batch=np.int64(B), steps=np.int64(steps), scores=np.array(scores, np.float64),
        W_final=rbm.rbm_weight.value, b_final=rbm.visible_bias.value, c_final=rbm.hidden_bias.value,
        fit_error=np.array(err), config=np.array(json.dumps(rbm.get_config(), default=str)),
        **{"infer_" + k: v for k, v in infer.items()}, **arrays)
    print(mode_name + size, "steps executed:", steps, "| reference raised:", err, "| scores:", np.round(scores, 4))


def run_dbn(out_dir=HERE):
    """dbn.py at HEAD with recording mock layers: documents D7."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_dbn", "/root/reference/ku/ebm/dbn.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class Mock:
        def __init__(self, name, i, o):
            self.name, self.input_shape, self.output_shape = name, (None, i), (None, o)

        def fit(self, V):
            pass

        def transform(self, V):
            return V

        def inv_transform(self, Hs):
            return Hs + 1

    res = {}
    dbn = mod.DBN()
    try:
        dbn.fit(np.zeros((2, 4)))
    except ValueError as e:
        res["fit_empty"] = "ValueError: %s" % e
    try:
        dbn.transform(np.zeros((2, 4)))
    except ValueError as e:
        res["transform_empty"] = "ValueError: %s" % e
    dbn.add_stack(Mock("a", 4, 3))
    try:
        dbn.add_stack(Mock("b", 3, 2))
        res["second_add_stack"] = "ok"
    except AttributeError as e:
        res["second_add_stack"] = "AttributeError: %s" % e
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            dbn.fit(np.zeros((2, 4)))
            res["fit"] = "ok"
        except AttributeError as e:
            res["fit"] = "AttributeError: %s" % e
    x = np.ones((2, 3))
    res["inv_transform_is_identity"] = bool(np.array_equal(dbn.inv_transform(x), x))
    res["constants"] = [mod.MODE_VISIBLE_BERNOULLI, mod.MODE_VISIBLE_GAUSSIAN, mod.MODE_COMPLEX]
    with open(os.path.join(out_dir, "ref_dbn.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("dbn:", res)


if __name__ == "__main__":
    for size in SIZES:
        run_rbm("bernoulli", size=size)
        run_rbm("gaussian", size=size)
    run_dbn()
