"""The reference-facing classes (keras_unsupervised_b200.ebm.RBM / DBN) on the GPU: the calls a user of
ku.ebm makes, with the reference's names, argument meaning, return types and error behaviour."""
import numpy as np
import pytest

from oracle import cd_oracle as O

pytestmark = pytest.mark.gpu


def _structured(rng, n, dim, protos=16, flip=0.05):
    P = (rng.random((protos, dim)) < 0.5)
    x = P[rng.integers(0, protos, n)]
    x ^= rng.random((n, dim)) < flip
    return x.astype(np.float32)


def test_rbm_surface_matches_reference(ctx, capsys):
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(0)
    hps = {"batch_size": 128, "epochs": 2, "lr": 1e-3}
    rbm = RBM(hps, 500, name="rbm", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    V = (rng.random((1000, 784)) < 0.13).astype(np.float32)
    rbm.fit(V, verbose=1)                      # 7 full minibatches + a 104-row remainder (rbm.py:110-111)
    out = capsys.readouterr().out
    assert "1 / 2  epochs" in out and "score:" in out      # rbm.py:115,234
    assert rbm.rbm_weight.shape == (784, 500) and rbm.hidden_bias.shape == (500,) and rbm.visible_bias.shape == (784,)
    h = rbm.transform(V[:100])                 # K.function returns a list of one array (rbm.py:89)
    assert isinstance(h, list) and len(h) == 1 and h[0].shape == (100, 500) and h[0].dtype == np.float32
    assert set(np.unique(h[0])) <= {0.0, 1.0}
    v = rbm.inv_transform(h)                   # and accepts one (dbn.py:55 feeds it straight back)
    assert v[0].shape == (100, 784) and set(np.unique(v[0])) <= {0.0, 1.0}
    fe = rbm.cal_free_energy([V[:100]])
    assert fe[0].shape == (100,)
    assert rbm.compute_output_shape((None, 784)) == (None, 500)
    cfg = rbm.get_config()
    assert cfg["output_dim"] == 500 and cfg["name"] == "rbm" and cfg["hps"] is hps and cfg["mode"] == 0
    assert rbm.call(V[:10]).shape == (10, 500)


def test_rbm_learns_structured_data(ctx):
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(1235)
    V = _structured(rng, 4096, 256)
    for dtype in ("float32", "bf16"):
        rbm = RBM({"batch_size": 128, "epochs": 1, "lr": 0.05, "normalize": "mean", "dtype": dtype, "k": 1}, 128,
                  name="r", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
        rbm.fit(V, verbose=0)
        first = rbm._machine.cd_step(V[:128], rbm._hparams(update_mask=0, want_stats=True))["recon_err"]
        rbm.hps["epochs"] = 15
        rbm.fit(V, verbose=0)
        last = rbm._machine.cd_step(V[:128], rbm._hparams(update_mask=0, want_stats=True))["recon_err"]
        assert last < 0.6 * first, (dtype, first, last)
        assert np.isfinite(rbm.rbm_weight).all()


def test_rbm_reference_schedule_mode(ctx, capsys):
    """compat='reference': three single-parameter runs and a per-step score print (rbm.py:214-234)."""
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(2)
    V = (rng.random((300, 200)) < 0.2).astype(np.float32)
    rbm = RBM({"batch_size": 128, "epochs": 1, "lr": 1e-3, "compat": "reference"}, 64, name="r",
              mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    before = ctx.timings(reset=True)
    rbm.fit(V, verbose=1)
    out = capsys.readouterr().out
    assert out.count("score:") == 3 and "3/3, score:" in out
    t = ctx.timings()
    # per step: 3 runs x (the chain kernel with the 3 projections + the float32-grade dW contraction) + score (2 projections
    # + 2 free energies) = 10 contraction launches
    assert t["gemm_launches"] == 3 * 10 and t["chain_launches"] == 3 * 3


def test_dbn_greedy_pretraining(ctx, capsys):
    from keras_unsupervised_b200.ebm import DBN, RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(3)
    V = _structured(rng, 2048, 784)
    hps = {"batch_size": 256, "epochs": 1, "lr": 1e-3, "dtype": "bf16"}
    dbn = DBN()
    with pytest.raises(ValueError):
        dbn.fit(V)                                   # dbn.py:47-48
    dims = [500, 500, 2000]
    for i, d in enumerate(dims):
        dbn.add_stack(RBM(hps, d, name="rbm%d" % i, mode=MODE_VISIBLE_BERNOULLI, context=ctx))
    dbn.fit(V, verbose=0)
    assert capsys.readouterr().out.count("Train rbm") == 3   # dbn.py:53
    H = dbn.transform(V[:64])
    assert H.shape == (64, 2000) and set(np.unique(H)) <= {0.0, 1.0}
    back = dbn.inv_transform(H)
    assert back.shape == (64, 784)
    bad = RBM(hps, 10, name="bad", mode=MODE_VISIBLE_BERNOULLI, context=ctx, input_dim=123)
    with pytest.raises(ValueError):
        dbn.add_stack(bad)                           # dbn.py:27-30


def test_one_epoch_fit_streamed_or_resident(ctx):
    """A one-epoch RBM.fit of a host array is streamed (kucd_rbm_fit_host) unless hps['stream'] is False (upload, then
    graph replay); both ways train the same parameters."""
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(41)
    X = (rng.random((1000, 200)) < 0.2).astype(np.float32)
    out = []
    for stream in (None, False):
        hps = {"batch_size": 128, "epochs": 1, "lr": 1e-3, "dtype": "bf16", "seed": 9}
        if stream is not None:
            hps["stream"] = stream
        r = RBM(hps, 96, name="p", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
        t0 = ctx.timings()["graph_launches"]
        r.fit(X, verbose=0)
        out.append((r.rbm_weight, r.hidden_bias, ctx.timings()["graph_launches"] - t0))
    assert out[0][2] in (0, 7) and out[1][2] == 8        # 7: KUCD_STREAM_GRAPH=1 replays the full minibatches of a streamed fit
    # same kernels, same draws, commuting column-statistic atomics (rng_math.cuh: stat_grid_round): the same bits
    assert np.array_equal(out[1][0], out[0][0]) and np.array_equal(out[1][1], out[0][1])


def test_training_is_bitwise_reproducible(ctx):
    """Run-to-run determinism.  The only order-dependent arithmetic of a training step used to be the fp32 atomics of the
    column statistics (db, dc): 1e-8 of noise on a bias that, in about one fit of ten of exactly this problem, moved the
    bf16 rounding of one stored probability (hidden unit 58) and with it 78 weights by lr * 2^-10
    (profiles/r02_call8_diag_stream_vs_resident.log).  The partial sums now lie on a grid on which fp32 addition is exact
    (stat_grid_round), so every run gives the same bits: 40 fits of the problem that showed it, both compute modes,
    Bernoulli and Gaussian visibles, CD-1 and CD-2 with momentum."""
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI, MODE_VISIBLE_GAUSSIAN

    rng = np.random.default_rng(41)
    X = (rng.random((1000, 200)) < 0.2).astype(np.float32)
    G = rng.normal(0, 1, (600, 200)).astype(np.float32)
    cases = [(X, MODE_VISIBLE_BERNOULLI, {"dtype": "bf16"}, 40),
             (X, MODE_VISIBLE_BERNOULLI, {"dtype": "float32", "k": 2, "momentum": 0.5}, 6),
             (X, MODE_VISIBLE_BERNOULLI, {"dtype": "bf16", "stream": False, "epochs": 2}, 6),
             (G, MODE_VISIBLE_GAUSSIAN, {"dtype": "bf16", "lr": 1e-4}, 6),
             (G, MODE_VISIBLE_GAUSSIAN, {"dtype": "float32", "lr": 1e-4}, 4)]
    for data, mode, extra, n in cases:
        first = None
        for _ in range(n):
            hps = {"batch_size": 128, "epochs": 1, "lr": 1e-3, "seed": 9}
            hps.update(extra)
            r = RBM(hps, 96, name="p", mode=mode, context=ctx)
            r.fit(data, verbose=0)
            got = (np.array(r.rbm_weight), np.array(r.hidden_bias), np.array(r.visible_bias))
            if first is None:
                first = got
            else:
                for a, b, name in zip(got, first, ("W", "c", "b")):
                    assert np.array_equal(a, b), (extra, name, float(np.abs(a - b).max()))


def test_dbn_generate_top_down(ctx):
    """Gibbs sampling in the top RBM, then the top-down pass of dbn.py:77-96.  With zero Gibbs sweeps
    it is exactly inv_transform of the starting hidden states (same Philox draws); with sweeps, samples of a stack
    trained on 16 prototypes land nearer to a prototype than random bits do."""
    from keras_unsupervised_b200.ebm import DBN, RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(5)
    V = _structured(rng, 4096, 256)
    hps = {"batch_size": 128, "epochs": 8, "lr": 0.1, "dtype": "bf16", "normalize": "mean", "seed": 3}
    dbn = DBN()
    for i, d in enumerate((128, 64)):
        dbn.add_stack(RBM(dict(hps), d, name="g%d" % i, mode=MODE_VISIBLE_BERNOULLI, context=ctx))
    with pytest.raises(ValueError):
        dbn.generate(4)                              # not trained yet
    dbn.fit(V, verbose=0)
    h0 = (rng.random((32, 64)) < 0.5).astype(np.float32)
    for layer in dbn._rbm_layers:                    # rewind the inference draw counters
        layer._machine.set_seed(layer.seed, layer._machine.counters()["step_count"])
    a = dbn.generate(32, gibbs_steps=0, h_init=h0)
    for layer in dbn._rbm_layers:
        layer._machine.set_seed(layer.seed, layer._machine.counters()["step_count"])
    b = dbn.inv_transform(h0)
    assert a.shape == (32, 256) and np.array_equal(a, b)
    g = dbn.generate(256, gibbs_steps=50, seed=1)
    assert g.shape == (256, 256) and set(np.unique(g)) <= {0.0, 1.0}
    protos = V[:512]                                 # every prototype occurs among 512 rows (5 % of the bits flipped)

    def nearest(x):                                  # mean Hamming distance to the nearest of them
        return np.abs(x[:, None, :] - protos[None, :, :]).sum(axis=2).min(axis=1).mean()

    noise = (rng.random((256, 256)) < V.mean()).astype(np.float32)
    # the oracle, trained the same way, generates at ~32 bits from the nearest prototype; random bits sit at ~107
    assert nearest(g) < 0.6 * nearest(noise)


def test_torch_cuda_tensors_zero_copy(ctx):
    """DLPack-style hand-over: device tensors go in and come out without touching the host."""
    import torch

    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rbm = RBM({"batch_size": 64, "epochs": 1, "lr": 1e-3, "dtype": "bf16", "seed": 9}, 256, name="r",
              mode=MODE_VISIBLE_BERNOULLI, context=ctx, return_list=False)
    x = (torch.rand(512, 320, device="cuda") < 0.3).float()
    before = ctx.timings(reset=True)
    h = rbm.transform(x)
    t = ctx.timings()
    assert h.is_cuda and h.shape == (512, 256)
    assert t["h2d_bytes"] == 0 and t["d2h_bytes"] == 0
    rbm._machine.set_seed(9, 0)
    h_np = rbm.transform(x.cpu().numpy())
    assert np.array_equal(h.cpu().numpy(), h_np)
    hb = rbm._machine.transform(x.to(torch.bfloat16), out_dtype=torch.uint8)
    assert hb.dtype == torch.uint8 and hb.shape == (512, 256)


def test_checkpoint_resume_is_exact(ctx, tmp_path):
    """save/load carries parameters (under the reference's variable names), the Philox position and the
    persistent chains: a resumed fit continues exactly where the uninterrupted one goes."""
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(7)
    V = _structured(rng, 1024, 192)
    hps = {"batch_size": 128, "epochs": 2, "lr": 1e-2, "dtype": "bf16", "persistent": True, "k": 2, "seed": 5}
    a = RBM(dict(hps), 96, name="a", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    a.fit(V, verbose=0)
    a.save(tmp_path / "ckpt.npz")
    z = np.load(tmp_path / "ckpt.npz")
    assert {"rbm_weight", "rbm_hidden_bias", "rbm_visible_bias", "chains", "seed", "step_count"} <= set(z.files)
    assert int(z["step_count"]) == 16
    a.fit(V, verbose=0)
    b = RBM(dict(hps), 96, name="b", mode=MODE_VISIBLE_BERNOULLI, context=ctx).load(tmp_path / "ckpt.npz")
    b.fit(V, verbose=0)
    np.testing.assert_allclose(a.rbm_weight, b.rbm_weight, rtol=0, atol=1e-6)
    np.testing.assert_allclose(a.hidden_bias, b.hidden_bias, rtol=0, atol=1e-6)
    assert np.array_equal(a._machine.get_chains(128), b._machine.get_chains(128))


def test_checkpoint_resume_with_momentum_and_score_streams_is_exact(ctx, tmp_path):
    """A checkpoint also carries the momentum buffers and the positions of the inference / score streams (round-1 ADVICE):
    a fit with momentum resumed from a checkpoint equals the uninterrupted one bit for bit, the epoch score it prints
    (its own Philox stream) included, and so does the next transform."""
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(8)
    V = _structured(rng, 900, 160)
    for dtype in ("bf16", "float32"):
        hps = {"batch_size": 128, "epochs": 2, "lr": 1e-3, "dtype": dtype, "momentum": 0.7, "weight_decay": 1e-4,
               "seed": 6, "stream": False}
        a = RBM(dict(hps), 64, name="a", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
        assert a.built is False or a._machine.get_momentum() is None
        a.fit(V, verbose=1)                       # verbose: the epoch score advances the score stream
        a.transform(V[:64])                       # ... and this the inference stream
        a.save(tmp_path / ("m_%s.npz" % dtype))
        z = np.load(tmp_path / ("m_%s.npz" % dtype))
        assert {"momentum_weight", "momentum_visible_bias", "momentum_hidden_bias", "infer_draws", "score_draws"} <= set(z.files)
        assert int(z["infer_draws"]) == 1 and int(z["score_draws"]) == 2 and np.abs(z["momentum_weight"]).max() > 0
        a.fit(V, verbose=1)
        ha = a.transform(V[:64])[0]
        b = RBM(dict(hps), 64, name="b", mode=MODE_VISIBLE_BERNOULLI, context=ctx).load(tmp_path / ("m_%s.npz" % dtype))
        b.fit(V, verbose=1)
        hb = b.transform(V[:64])[0]
        assert np.array_equal(a.rbm_weight, b.rbm_weight) and np.array_equal(a.hidden_bias, b.hidden_bias)
        assert np.array_equal(a.visible_bias, b.visible_bias) and np.array_equal(ha, hb)
        assert [h["last_score"] for h in a.history[2:]] == [h["last_score"] for h in b.history]
        for x, y in zip(a._machine.get_momentum(), b._machine.get_momentum()):
            assert np.array_equal(x, y)


def test_gaussian_visible_rbm_runs_the_default_mode(ctx):
    """MODE_VISIBLE_GAUSSIAN is the constructor default (rbm.py:22) and what examples/rbm uses: relu-threshold
    hiddens, unit-variance Gaussian reconstructions drawn from the engine's Philox stream (Box-Muller)."""
    from keras_unsupervised_b200.ebm import RBM

    rng = np.random.default_rng(11)
    seed = 21
    X = rng.normal(0, 1, (512, 200)).astype(np.float32)
    rbm = RBM({"batch_size": 128, "epochs": 1, "lr": 1e-4, "seed": seed}, 64, name="g", context=ctx)
    assert rbm.mode == 1
    rbm.build((None, 200))
    W, b, c = rbm._machine.get_params()
    orc = O.OracleRBM(W, b, c, mode=O.MODE_VISIBLE_GAUSSIAN)
    hh = (rng.random((512, 64)) < 0.5).astype(np.float32)
    v = rbm.inv_transform(hh)[0]                 # first inference draw: id 2^63 + 0, unit normals
    n = O.philox_normal(seed, O.draw_id("infer", 0), 0, 512, 200)
    v_ref, _ = orc.sample_v(hh, n)
    np.testing.assert_allclose(v, v_ref, rtol=1e-5, atol=2e-5)
    assert abs(float((v - orc.pre_v(hh)).std()) - 1.0) < 0.02
    rbm.fit(X, verbose=0)
    assert np.isfinite(rbm.rbm_weight).all() and not np.array_equal(rbm.rbm_weight, W)


def test_reference_schedule_fit_matches_oracle_replay(ctx):
    """compat='reference' through RBM.fit: W, then c, then b with fresh draws each (rbm.py:214-216) and the printed
    score (rbm.py:227-234), replayed by the oracle with the engine's Philox stream."""
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(31)
    V = (rng.random((300, 200)) < 0.2).astype(np.float32)
    hps = {"batch_size": 128, "epochs": 2, "lr": 1e-3, "compat": "reference", "seed": 77}
    rbm = RBM(hps, 96, name="r", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    rbm.build((None, 200))
    W, b, c = rbm._machine.get_params()
    orc = O.OracleRBM(W, b, c)
    rbm.fit(V, verbose=0)
    scores = O.philox_fit_reference(orc, V, 128, 2, 1e-3, 77)
    got = [h["last_score"] for h in rbm.history]
    assert len(got) == 6
    np.testing.assert_allclose(got, scores, rtol=2e-3)          # a flipped sample moves a score slightly
    assert np.abs(rbm.rbm_weight - orc.W).mean() < 2e-6 and np.abs(rbm.rbm_weight - orc.W).max() < 5e-3
    assert np.abs(rbm.hidden_bias - orc.c).max() < 5e-3 and np.abs(rbm.visible_bias - orc.b).max() < 5e-3


def test_dbn_fit_matches_oracle_replay(ctx):
    """DBN.fit (dbn.py:51-55): layer l+1 trains on the sampled hidden states of layer l over the whole data set.
    Replayed by the oracle layer by layer with each layer's own Philox stream (stack position offsets the seed)."""
    from keras_unsupervised_b200.ebm import DBN, RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(37)
    X = (rng.random((512, 160)) < 0.3).astype(np.float32)
    dims = [160, 96, 64]
    hps = {"batch_size": 128, "epochs": 1, "lr": 1e-3, "seed": 5}
    dbn = DBN()
    layers = [RBM(dict(hps), d, name="l%d" % i, mode=MODE_VISIBLE_BERNOULLI, context=ctx) for i, d in enumerate(dims[1:])]
    for l in layers:
        dbn.add_stack(l)
    assert [l.seed for l in layers] == [5, 5 + 1000003]
    for l, d in zip(layers, dims[:-1]):
        l.build((None, d))
    init = [l._machine.get_params() for l in layers]
    dbn.fit(X, verbose=0)
    cur = X
    for l, (W, b, c) in zip(layers, init):
        orc = O.OracleRBM(W, b, c)
        O.philox_fit(orc, cur, 128, 1, 1e-3, l.seed)
        assert np.abs(l.rbm_weight - orc.W).mean() < 2e-6, l.name
        # dbn.py:55: the next layer sees this layer's sampled states of the whole data set (first inference draw)
        cur, _ = orc.sample_h(cur, O.philox_uniform(l.seed, O.draw_id("infer", 0), 0, cur.shape[0], orc.H))
    H = dbn.transform(X)
    assert H.shape == (512, 64)
    Vb = dbn.inv_transform(H)
    assert Vb.shape == (512, 160) and set(np.unique(Vb)) <= {0.0, 1.0}


def test_dbn_save_load_round_trip(ctx, tmp_path):
    from keras_unsupervised_b200.ebm import DBN, RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(41)
    X = (rng.random((256, 96)) < 0.3).astype(np.float32)
    hps = {"batch_size": 128, "epochs": 1, "lr": 1e-3, "dtype": "bf16", "seed": 3}
    dbn = DBN()
    for i, d in enumerate((64, 48)):
        dbn.add_stack(RBM(dict(hps), d, name="l%d" % i, mode=MODE_VISIBLE_BERNOULLI, context=ctx))
    dbn.fit(X, verbose=0)
    dbn.save(tmp_path / "ck")
    twin = DBN.load(tmp_path / "ck", context=ctx)
    assert len(twin._rbm_layers) == 2 and [l.seed for l in twin._rbm_layers] == [l.seed for l in dbn._rbm_layers]
    for a, b in zip(dbn._rbm_layers, twin._rbm_layers):
        assert np.array_equal(a.rbm_weight, b.rbm_weight) and np.array_equal(a.hidden_bias, b.hidden_bias)
        a._machine.set_seed(a.seed, 0)
        b._machine.set_seed(b.seed, 0)
    assert np.array_equal(dbn.transform(X), twin.transform(X))
    # after fine-tuning the lower layers have untied generative weights: a checkpoint must carry them, or the reloaded
    # stack would generate through the recognition weights
    dbn.fine_tune(X, epochs=1, lr=1e-2)
    dbn.save(tmp_path / "ck2")
    twin2 = DBN.load(tmp_path / "ck2", context=ctx)
    assert len(twin2._gen) == 1
    for g, h in zip(dbn._gen, twin2._gen):
        for x, y in zip(g.get_params(), h.get_params()):
            assert np.array_equal(x, y)
    assert not np.array_equal(dbn._gen[0].get_params()[0], dbn._rbm_layers[0].rbm_weight)   # they did move apart


def test_refitting_with_fresh_arrays_never_replays_a_stale_graph(ctx):
    """Every RBM.fit(array) builds a new resident data set; the captured step graph must follow it (the allocator
    is free to hand a released data set's address out again)."""
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(51)
    hps = {"batch_size": 128, "epochs": 2, "lr": 1e-2, "dtype": "bf16", "seed": 4}
    data = [(rng.random((512, 192)) < q).astype(np.float32) for q in (0.1, 0.5, 0.9, 0.3)]
    a = RBM(dict(hps), 64, name="a", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    for X in data:
        a.fit(X, verbose=0)
    W, b, c = a._machine.get_params()
    # replay with the oracle: the visible bias follows the data's mean, so stale data would be obvious
    b0 = RBM(dict(hps), 64, name="b", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    b0.build((None, 192))
    orc = O.OracleRBM(*b0._machine.get_params(), compute="bf16")
    step = 0
    for X in data:
        step = O.philox_fit(orc, X, 128, 2, 1e-2, 4, step0=step)
    assert np.abs(W - orc.W).mean() < 5e-5 and np.abs(b - orc.b).max() < 2e-2
