"""Up-down fine-tuning of a stack (the reference stops at greedy pretraining, dbn.py:34-55).

CPU part: (1) the oracle's restatement learns - the delta rule raises the likelihood it ascends, an up-down pass lowers
the stack's reconstruction error; (2) DBN.fine_tune drives the engine layer in the order and with the operands
OracleDBN.up_down_step prescribes - checked by putting an oracle-backed stand-in behind the Machine interface (same
Philox draw ids as include/kucd.h documents) and comparing the parameters with the oracle's own replay.

GPU part (bottom, `-m gpu`): kucd_rbm_delta_rule against the oracle, and DBN.fine_tune against the oracle replay."""
import os

import numpy as np
import pytest

from keras_unsupervised_b200 import _lib as L
from keras_unsupervised_b200 import engine as E
from keras_unsupervised_b200.ebm import dbn as D
from keras_unsupervised_b200.ebm import rbm as R
from oracle import cd_oracle as O


# ---------------------------------------------------------------------------------------------------------------
# the oracle's restatement
# ---------------------------------------------------------------------------------------------------------------
def _bce(p, t):
    p = np.clip(p.astype(np.float64), 1e-9, 1 - 1e-9)
    return float(-np.mean(t * np.log(p) + (1 - t) * np.log(1 - p)))


@pytest.mark.parametrize("forward", [True, False])
@pytest.mark.parametrize("compute", ["f64", "bf16"])
def test_delta_rule_ascends_the_log_likelihood(forward, compute):
    rng = np.random.default_rng(3)
    V, H, B = 20, 12, 96
    W, b, c = O.OracleRBM.init_params(V, H, seed=5)
    r = O.OracleRBM(W, b, c, compute=compute)
    n_in, n_out = (V, H) if forward else (H, V)
    x = (rng.random((B, n_in)) < 0.5).astype(np.float32)
    t = np.concatenate([x, x], axis=1)[:, 3:3 + n_out].copy()       # targets are copies of inputs: learnable
    prob = r.prob_h if forward else r.prob_v
    before = _bce(prob(x), t)
    for _ in range(150):
        p = r.delta_rule(forward, x, t, 0.1, scale=1.0 / B)
    assert p.shape == (B, n_out)
    assert _bce(prob(x), t) < 0.6 * before
    # the untouched bias stays untouched
    assert np.array_equal(r.b if forward else r.c, b if forward else c)


def test_delta_rule_is_the_gradient_of_the_cross_entropy():
    """One step with lr -> the finite-difference gradient of sum log p(t | x) with respect to W and the bias."""
    rng = np.random.default_rng(8)
    V, H, B = 6, 5, 7
    W, b, c = O.OracleRBM.init_params(V, H, seed=1)
    x = (rng.random((B, V)) < 0.5).astype(np.float32)
    t = (rng.random((B, H)) < 0.5).astype(np.float32)

    def loglik(Wm, cm):
        a = x.astype(np.float64) @ Wm + cm
        return float(np.sum(t * a - np.logaddexp(0, a)))

    r = O.OracleRBM(W, b, c, compute="f64")
    r.delta_rule(True, x, t, 1.0)
    gW, gc = r.W - W, r.c - c
    eps = 1e-4
    for (i, j) in ((0, 0), (3, 2), (5, 4)):
        Wp, Wm = W.astype(np.float64).copy(), W.astype(np.float64).copy()
        Wp[i, j] += eps
        Wm[i, j] -= eps
        fd = (loglik(Wp, c.astype(np.float64)) - loglik(Wm, c.astype(np.float64))) / (2 * eps)
        assert abs(fd - gW[i, j]) < 1e-4
    cp, cm = c.astype(np.float64).copy(), c.astype(np.float64).copy()
    cp[1] += eps
    cm[1] -= eps
    assert abs((loglik(W.astype(np.float64), cp) - loglik(W.astype(np.float64), cm)) / (2 * eps) - gc[1]) < 1e-4


def _prototype_data(rng, n, dim, n_proto=8, flip=0.05):
    protos = (rng.random((n_proto, dim)) < 0.5).astype(np.float32)
    x = protos[rng.integers(0, n_proto, n)]
    noise = rng.random((n, dim)) < flip
    return np.where(noise, 1 - x, x).astype(np.float32)


def _draws(seeds, steps, layer_dims, rows, k, infer_counts):
    """The draws one fine-tune minibatch consumes, with the engine's ids (include/kucd.h, kucd_rbm_set_seed): the n-th
    transform / inv_transform call of a model uses 2^63 + n, training step s uses 64 s + phase."""
    rec_seeds, gen_seeds = seeds
    L_ = len(layer_dims) - 1
    up = [O.philox_uniform(rec_seeds[l], O.draw_id("infer", infer_counts["rec"][l]), 0, rows, layer_dims[l + 1])
          for l in range(L_ - 1)]
    down = [O.philox_uniform(gen_seeds[l], O.draw_id("infer", infer_counts["gen"][l]), 0, rows, layer_dims[l])
            for l in range(L_ - 1)]
    top_seed = rec_seeds[L_ - 1]
    u_h = [O.philox_uniform(top_seed, O.draw_id("train", steps, 0 if t == 0 else 2 * t + 1), 0, rows, layer_dims[L_])
           for t in range(k)]
    u_v = [None] + [O.philox_uniform(top_seed, O.draw_id("train", steps, 2 * t), 0, rows, layer_dims[L_ - 1])
                    for t in range(1, k + 1)]
    for l in range(L_ - 1):
        infer_counts["rec"][l] += 1
        infer_counts["gen"][l] += 1
    return {"up": up, "down": down, "top": (u_h, u_v)}


def test_up_down_pass_improves_the_stack():
    """Greedy CD-1 pretraining of a 24-16-12 stack on prototype data, then up-down passes: the top-down reconstruction
    of the data (up through the recognition weights, down through the generative ones) gets better, and the generative
    weights leave the recognition weights they started from."""
    rng = np.random.default_rng(21)
    dims, B, N, lr = [24, 16, 12], 32, 256, 0.05
    X = _prototype_data(rng, N, dims[0])
    dbn = O.OracleDBN()
    seeds = [11, 12]
    cur = X
    for l in range(2):
        r = O.OracleRBM(*O.OracleRBM.init_params(dims[l], dims[l + 1], seed=seeds[l]), compute="f64")
        O.philox_fit(r, cur, B, 30, lr, seeds[l], normalize=True)
        dbn.add_stack(r)
        cur, _ = r.sample_h(cur, O.philox_uniform(seeds[l], O.draw_id("infer", 0), 0, N, dims[l + 1]))

    def recon_err():
        gen = dbn.untie()
        h1 = dbn.layers[0].prob_h(X)
        return float(np.mean((gen[0].prob_v((h1 > 0.5).astype(np.float32)) - X) ** 2))

    before = recon_err()
    W_rec0 = dbn.layers[0].W.copy()
    counts = {"rec": [1, 0], "gen": [0]}
    step = 30 * (N // B)
    for epoch in range(20):
        for lo in range(0, N, B):
            d = _draws((seeds, [seeds[0] + 500009]), step, dims, B, 1, counts)
            s, t = dbn.up_down_step(X[lo:lo + B], d, lr, k=1, scale=1.0 / B)
            step += 1
    assert len(s) == 2 and len(t) == 2 and s[1].shape == (B, 16) and t[0].shape == (B, 24)
    assert recon_err() < before
    assert np.abs(dbn.gen[0].W - dbn.layers[0].W).max() > 1e-3      # untied
    assert np.abs(dbn.layers[0].W - W_rec0).max() > 1e-4            # and the recognition weights moved too


# ---------------------------------------------------------------------------------------------------------------
# DBN.fine_tune's orchestration, against an oracle-backed stand-in of the engine layer
# ---------------------------------------------------------------------------------------------------------------
class _Ctx:
    rank, world = 0, 1

    def sync(self):
        pass


class OracleMachine:
    """The Machine interface DBN.fine_tune uses, computed by the oracle with the engine's Philox draw ids."""

    hparams = staticmethod(E.Machine.hparams)

    def __init__(self, ctx, V, H, mode, compute, seed=None):
        self.ctx, self.V, self.H, self.mode, self.compute = ctx, V, H, mode, compute
        self.seed, self.infer, self.step, self.last, self.orc = seed, 0, 0, None, None
        self.log = []

    def set_seed(self, seed, step_count=0):
        self.seed, self.step = seed, step_count

    def counters(self):
        return {"seed": self.seed, "step_count": self.step, "n_chains": 0}

    def set_params(self, W=None, b=None, c=None):
        if self.orc is None:
            self.orc = O.OracleRBM(W, b, c, mode=self.mode, compute="f64")
        else:
            self.orc.W = np.asarray(W, np.float32) if W is not None else self.orc.W
            self.orc.b = np.asarray(b, np.float32) if b is not None else self.orc.b
            self.orc.c = np.asarray(c, np.float32) if c is not None else self.orc.c

    def get_params(self):
        return self.orc.W.copy(), self.orc.b.copy(), self.orc.c.copy()

    def transform(self, v, u=None, want_p=False, out_dtype=None):
        u = O.philox_uniform(self.seed, O.draw_id("infer", self.infer), 0, v.shape[0], self.H)
        self.infer += 1
        self.log.append("transform")
        return self.orc.sample_h(v, u)[0]

    def inv_transform(self, h, u=None, want_p=False, out_dtype=None):
        u = O.philox_uniform(self.seed, O.draw_id("infer", self.infer), 0, h.shape[0], self.V)
        self.infer += 1
        self.log.append("inv_transform")
        return self.orc.sample_v(h, u)[0]

    def cd_step(self, v, hp, **kw):
        rows, k = v.shape[0], hp.k
        u_h = [O.philox_uniform(self.seed, O.draw_id("train", self.step, 0 if t == 0 else 2 * t + 1), 0, rows, self.H)
               for t in range(k)]
        u_v = [None] + [O.philox_uniform(self.seed, O.draw_id("train", self.step, 2 * t), 0, rows, self.V)
                        for t in range(1, k + 1)]
        self.last = self.orc.fused_step(v, u_h, u_v, hp.lr, k=k, scale=(1.0 / rows if hp.normalize else 1.0))
        self.step += 1
        self.log.append("cd_step")

    def last_stats(self, rows, states=True, grads=True):
        assert states and not grads                      # fine_tune asks for the states only
        return {"v_neg": self.last["v_neg"], "h_pos": self.last["h_pos"], "h_neg": self.last["h_neg"]}

    def delta_rule(self, forward, x, target, lr, normalize=False):
        self.orc.delta_rule(forward, x, target, lr, scale=(1.0 / x.shape[0] if normalize else 1.0))
        self.log.append("delta_fwd" if forward else "delta_bwd")


@pytest.fixture
def oracle_engine(monkeypatch):
    monkeypatch.setattr(R, "Machine", OracleMachine)
    monkeypatch.setattr(E, "Machine", OracleMachine)
    return _Ctx()


def _stack(ctx, dims, **hps):
    base = {"batch_size": 16, "epochs": 1, "lr": 0.05, "normalize": "mean", "seed": 7}
    base.update(hps)
    dbn = D.DBN()
    for l in range(len(dims) - 1):
        dbn.add_stack(R.RBM(base, dims[l + 1], name="l%d" % l, mode=R.MODE_VISIBLE_BERNOULLI, context=ctx,
                            input_dim=dims[l]))
    return dbn


def test_fine_tune_drives_the_engine_as_the_oracle_prescribes(oracle_engine):
    rng = np.random.default_rng(2)
    dims, B, N, k = [20, 14, 10, 8], 16, 40, 2            # three layers; 2 full minibatches + a remainder of 8
    X = _prototype_data(rng, N, dims[0])
    dbn = _stack(oracle_engine, dims)
    # the oracle's own replay starts from the same parameters and stream positions
    ref = O.OracleDBN()
    for layer in dbn._rbm_layers:
        ref.add_stack(O.OracleRBM(*layer._machine.get_params(), compute="f64"))
    rec_seeds = [layer.seed for layer in dbn._rbm_layers]
    gen_seeds = [s + 500009 for s in rec_seeds[:-1]]
    out = dbn.fine_tune(X, epochs=2, lr=0.05, k=k)
    assert out is dbn
    counts = {"rec": [0, 0, 0], "gen": [0, 0]}
    step = 0
    for epoch in range(2):
        for lo in range(0, N, B):
            rows = min(B, N - lo)
            d = _draws((rec_seeds, gen_seeds), step, dims, rows, k, counts)
            ref.up_down_step(X[lo:lo + rows], d, 0.05, k=k, scale=1.0 / rows)
            step += 1
    for l, layer in enumerate(dbn._rbm_layers):
        W, b, c = layer._machine.get_params()
        np.testing.assert_array_equal(W, ref.layers[l].W)
        np.testing.assert_array_equal(b, ref.layers[l].b)
        np.testing.assert_array_equal(c, ref.layers[l].c)
    for l, g in enumerate(dbn._gen):
        Wg, bg, cg = g.get_params()
        np.testing.assert_array_equal(Wg, ref.gen[l].W)
        np.testing.assert_array_equal(bg, ref.gen[l].b)
    # per minibatch: up through the recognition models, CD on the top one, down through the generative ones, updates
    per_step = ["transform", "delta_fwd"]
    assert dbn._rbm_layers[0]._machine.log == per_step * 6 and dbn._rbm_layers[1]._machine.log == per_step * 6
    assert dbn._rbm_layers[2]._machine.log == ["cd_step"] * 6
    assert dbn._gen[0].log == ["inv_transform", "delta_bwd"] * 6
    # after untying, the stack goes down through the generative weights
    before = dbn._gen[1].infer
    v = dbn.inv_transform(np.zeros((4, dims[-1]), np.float32))
    assert v.shape == (4, dims[0]) and dbn._gen[1].infer == before + 1 and dbn._gen[1].log[-1] == "inv_transform"


def test_fine_tune_argument_errors(oracle_engine):
    with pytest.raises(ValueError):
        D.DBN().fine_tune(np.zeros((4, 4), np.float32))                       # empty stack (dbn.py:47-48)
    one = _stack(oracle_engine, [6, 4])
    with pytest.raises(ValueError):
        one.fine_tune(np.zeros((4, 6), np.float32))                           # nothing below the top RBM
    mixed = D.DBN()
    mixed.add_stack(R.RBM({"batch_size": 4, "epochs": 1, "lr": 0.1}, 4, name="g", mode=R.MODE_VISIBLE_GAUSSIAN,
                          context=oracle_engine, input_dim=6))
    mixed.add_stack(R.RBM({"batch_size": 4, "epochs": 1, "lr": 0.1}, 3, name="t", mode=R.MODE_VISIBLE_BERNOULLI,
                          context=oracle_engine, input_dim=4))
    with pytest.raises(ValueError):
        mixed.fine_tune(np.zeros((4, 6), np.float32))                         # Gaussian visibles: not implemented


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------


@pytest.fixture(scope="module")
def ctx():
    return E.Context(device=0, seed=42)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("f32", 1e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("forward", [True, False])
def test_gpu_delta_rule_matches_the_oracle(ctx, dtype, tol, forward):
    rng = np.random.default_rng(31)
    V, H, B = 333, 130, 200                                   # ragged sizes, a partial row tile
    compute = L.COMPUTE_F32X3 if dtype == "f32" else L.COMPUTE_BF16
    W, b, c = O.OracleRBM.init_params(V, H, seed=9)
    m = E.Machine(ctx, V, H, L.MODE_VISIBLE_BERNOULLI, compute, seed=5)
    m.set_params(W, b, c)
    orc = O.OracleRBM(W, b, c, compute="f64" if dtype == "f32" else "bf16")
    n_in, n_out = (V, H) if forward else (H, V)
    for step in range(3):
        x = (rng.random((B, n_in)) < 0.3).astype(np.float32)
        t = (rng.random((B, n_out)) < 0.4).astype(np.float32)
        m.delta_rule(forward, x, t, 1e-2, normalize=(step == 2))
        orc.delta_rule(forward, x, t, 1e-2, scale=(1.0 / B if step == 2 else 1.0))
    Wd, bd, cd = m.get_params()
    scale = np.abs(orc.W - W).max()
    assert np.abs(Wd - orc.W).max() <= tol * max(scale, 1e-3) + 1e-6
    np.testing.assert_allclose(bd, orc.b, rtol=tol, atol=tol * 1e-2)
    np.testing.assert_allclose(cd, orc.c, rtol=tol, atol=tol * 1e-2)
    assert np.array_equal(bd, b) if forward else np.array_equal(cd, c)      # the other bias is not touched


@pytest.mark.gpu
def test_gpu_fine_tune_matches_the_oracle_replay(ctx):
    rng = np.random.default_rng(2)
    dims, B, N, k = [200, 96, 64], 64, 160, 1               # 2 full minibatches + a remainder of 32
    X = _prototype_data(rng, N, dims[0])
    dbn = _stack(ctx, dims, batch_size=B, dtype="float32")
    ref = O.OracleDBN()
    for layer in dbn._rbm_layers:
        ref.add_stack(O.OracleRBM(*layer._machine.get_params(), compute="f64"))
    rec_seeds = [layer.seed for layer in dbn._rbm_layers]
    gen_seeds = [s + 500009 for s in rec_seeds[:-1]]
    dbn.fine_tune(X, epochs=1, lr=0.05, k=k)
    counts = {"rec": [0, 0], "gen": [0]}
    for step, lo in enumerate(range(0, N, B)):
        rows = min(B, N - lo)
        d = _draws((rec_seeds, gen_seeds), step, dims, rows, k, counts)
        ref.up_down_step(X[lo:lo + rows], d, 0.05, k=k, scale=1.0 / rows)
    # float32-grade mode: probabilities agree to ~1e-6, a sample may flip where u sits inside that gap
    for l, layer in enumerate(dbn._rbm_layers):
        W, b, c = layer._machine.get_params()
        assert np.abs(W - ref.layers[l].W).mean() < 1e-5
    Wg, bg, _ = dbn._gen[0].get_params()
    assert np.abs(Wg - ref.gen[0].W).mean() < 1e-5 and np.abs(bg - ref.gen[0].b).mean() < 1e-4
