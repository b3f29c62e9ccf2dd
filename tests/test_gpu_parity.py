"""GPU parity: libkucd.so (through the C ABI / ctypes) against the CPU oracle on the same seeded inputs.

Bars: sampled states bit-exact under injected, margin-conditioned uniforms;
probabilities / free energy / statistics within 1e-5 relative in float32 mode (2e-2 in bf16 mode, where
the oracle rounds the same operands to bf16 first, so the observed gap is far smaller).
"""
import numpy as np
import pytest

from oracle import cd_oracle as O

pytestmark = pytest.mark.gpu

RTOL_F32 = 1e-5   # north_star: probabilities, free energy, reconstruction error, float32
RTOL_BF16 = 2e-2  # north_star: bf16


def _machine(ctx, V, H, compute, mode=0, seed=0, pseed=0):
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Machine

    m = Machine(ctx, V, H, mode, L.COMPUTE_F32X3 if compute == "f32" else L.COMPUTE_BF16, seed=seed)
    W, b, c = O.OracleRBM.init_params(V, H, seed=pseed)
    m.set_params(W, b, c)
    orc = O.OracleRBM(W, b, c, mode=mode, compute="f64" if compute == "f32" else "bf16")
    return m, orc


def _data(rng, rows, V, q=0.3):
    return (rng.random((rows, V)) < q).astype(np.float32)


@pytest.mark.parametrize("compute", ["f32", "bf16"])
@pytest.mark.parametrize("rows,V,H", [(128, 784, 500), (96, 784, 500), (300, 200, 333), (1, 64, 64), (257, 1000, 72)])
def test_transform_and_inverse_bit_exact(ctx, compute, rows, V, H):
    rng = np.random.default_rng(rows + V)
    m, orc = _machine(ctx, V, H, compute)
    v = _data(rng, rows, V)
    u = O.lattice_uniform(rng, (rows, H))
    (u,), _, _ = O.condition_margin(orc, v, [u], [None], k=0)
    h, p = m.transform(v, u=u, want_p=True)
    h_ref, p_ref = orc.sample_h(v, u)
    np.testing.assert_allclose(p, p_ref, rtol=RTOL_F32 if compute == "f32" else RTOL_BF16, atol=1e-7)
    assert np.array_equal(h, h_ref)

    hh = _data(rng, rows, H, 0.5)
    u2 = O.lattice_uniform(rng, (rows, V))
    p2_ref = orc.prob_v(hh)
    bad = np.abs(u2 - p2_ref) <= 1e-4
    u2[bad] = np.where(p2_ref[bad] > 0.5, 0.0, 0.999)
    vv, p2 = m.inv_transform(hh, u=u2, want_p=True)
    np.testing.assert_allclose(p2, p2_ref, rtol=RTOL_F32 if compute == "f32" else RTOL_BF16, atol=1e-7)
    assert np.array_equal(vv, (u2 < p2_ref).astype(np.float32))


def test_transform_edge_uniforms(ctx):
    """Strict < (K.less, rbm.py:46): u = 0 fires unless p = 0; u just below 1 never fires for p < 1."""
    rng = np.random.default_rng(5)
    m, orc = _machine(ctx, 128, 64, "f32")
    v = _data(rng, 64, 128)
    h0 = m.transform(v, u=np.zeros((64, 64), np.float32))
    h1 = m.transform(v, u=np.full((64, 64), 1.0 - 2.0 ** -23, np.float32))
    assert h0.min() == 1.0 and h1.max() == 0.0


@pytest.mark.parametrize("compute", ["f32", "bf16"])
def test_real_valued_visibles(ctx, compute):
    """The example feeds grey levels in [0,1], not bits (examples/rbm/rbm_softmax_mnist.py:103)."""
    rng = np.random.default_rng(11)
    rows, V, H = 200, 784, 500
    m, orc = _machine(ctx, V, H, compute)
    v = (rng.integers(0, 256, (rows, V)) / 255.0).astype(np.float32)
    p = m.transform(v, want_p=True)[1]
    np.testing.assert_allclose(p, orc.prob_h(v), rtol=RTOL_F32 if compute == "f32" else RTOL_BF16, atol=1e-7)
    fe = m.free_energy(v)
    np.testing.assert_allclose(fe, orc.free_energy(v), rtol=RTOL_F32 if compute == "f32" else RTOL_BF16)


@pytest.mark.parametrize("compute", ["f32", "bf16"])
@pytest.mark.parametrize("rows,V,H", [(128, 784, 500), (37, 300, 130), (1024, 512, 512)])
def test_free_energy(ctx, compute, rows, V, H):
    rng = np.random.default_rng(7)
    m, orc = _machine(ctx, V, H, compute)
    v = _data(rng, rows, V)
    np.testing.assert_allclose(m.free_energy(v), orc.free_energy(v),
                               rtol=RTOL_F32 if compute == "f32" else RTOL_BF16)


def test_free_energy_matches_partition_sum_on_tiny_rbm(ctx):
    """F(v) = -log sum_h exp(-E(v,h)) checked by brute force over all 2^H hidden states (H = 10)."""
    rng = np.random.default_rng(3)
    V, H, rows = 64, 10, 16
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Machine

    W = rng.normal(0, 0.5, (V, H)).astype(np.float32)
    b = rng.normal(0, 0.5, V).astype(np.float32)
    c = rng.normal(0, 0.5, H).astype(np.float32)
    m = Machine(ctx, V, H, 0, L.COMPUTE_F32X3)
    m.set_params(W, b, c)
    v = _data(rng, rows, V)
    hs = ((np.arange(1 << H)[:, None] >> np.arange(H)) & 1).astype(np.float64)
    energy = -(v.astype(np.float64) @ b)[:, None] - hs @ c - (v.astype(np.float64) @ W) @ hs.T
    brute = -np.log(np.exp(-energy).sum(1))
    np.testing.assert_allclose(m.free_energy(v), brute, rtol=1e-5)


def _inject(rng, rows, V, H, k, persistent=False):
    u_h = [O.lattice_uniform(rng, (rows, H)) for _ in range(k)]
    u_v = [None] + [O.lattice_uniform(rng, (rows, V)) for _ in range(k)]
    u_hc = O.lattice_uniform(rng, (rows, H)) if persistent else None
    return u_h, u_v, u_hc


@pytest.mark.parametrize("compute", ["f32", "bf16"])
@pytest.mark.parametrize("rows,V,H,k", [(128, 784, 500, 1), (96, 784, 500, 1), (64, 256, 192, 3), (200, 130, 70, 2)])
def test_cd_step_matches_oracle(ctx, compute, rows, V, H, k):
    rng = np.random.default_rng(17 + k)
    m, orc = _machine(ctx, V, H, compute)
    v = _data(rng, rows, V, 0.13)
    u_h, u_v, _ = _inject(rng, rows, V, H, k)
    u_h, u_v, _ = O.condition_margin(orc, v, u_h, u_v, k=k)
    from keras_unsupervised_b200.engine import Machine

    st = orc.fused_step(v, u_h, u_v, lr=1e-3, k=k)
    m.cd_step(v, Machine.hparams(lr=1e-3, k=k), u_h=u_h, u_v=u_v)
    got = m.last_stats(rows)
    assert np.array_equal(got["h_pos"], st["h_pos"])
    assert np.array_equal(got["v_neg"], st["v_neg"])
    tol = RTOL_F32 if compute == "f32" else RTOL_BF16
    h_neg_ref = st["h_neg"] if compute == "f32" else O.bf16_round(st["h_neg"])
    np.testing.assert_allclose(got["h_neg"], h_neg_ref, rtol=tol, atol=1e-7)
    scale = max(1.0, float(np.abs(st["dW"]).max()))
    np.testing.assert_allclose(got["dW"], st["dW"], rtol=tol, atol=tol * scale)
    np.testing.assert_allclose(got["dc"], st["dc"], rtol=tol, atol=tol * rows)
    np.testing.assert_allclose(got["db"], st["db"], rtol=tol, atol=1e-6)
    W, b, c = m.get_params()
    np.testing.assert_allclose(W, orc.W, rtol=tol, atol=1e-6)
    np.testing.assert_allclose(b, orc.b, rtol=tol, atol=1e-6)
    np.testing.assert_allclose(c, orc.c, rtol=tol, atol=1e-6)


def test_reference_three_pass_schedule(ctx):
    """rbm.py:214-233: W, then c with fresh draws and the new W, then b, then the score chain."""
    rng = np.random.default_rng(23)
    rows, V, H = 128, 784, 500
    m, orc = _machine(ctx, V, H, "f32")
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Machine

    v = _data(rng, rows, V, 0.13)
    draws = []
    for mask in (1, 2, 4):
        u_h, u_v, _ = _inject(rng, rows, V, H, 1)
        u_h, u_v, _ = O.condition_margin(orc, v, u_h, u_v, k=1)
        draws.append((u_h[0], u_v[1]))
        orc.apply(orc.cd_stats(v, u_h, u_v), 1e-3, mask)
        m.cd_step(v, Machine.hparams(lr=1e-3, update_mask={1: L.UPDATE_W, 2: L.UPDATE_C, 4: L.UPDATE_B}[mask]),
                  u_h=u_h, u_v=u_v)
    u_h, u_v, _ = _inject(rng, rows, V, H, 1)
    u_h, u_v, _ = O.condition_margin(orc, v, u_h, u_v, k=1)
    s_ref = orc.score(v, u_h[0], u_v[1])
    s = m.score(v, u_h=u_h[0], u_v=u_v[1])
    W, b, c = m.get_params()
    np.testing.assert_allclose(W, orc.W, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(b, orc.b, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(c, orc.c, rtol=1e-5, atol=1e-6)
    assert abs(s - s_ref) <= 1e-5 * max(1.0, abs(s_ref))


def test_persistent_chains(ctx):
    rng = np.random.default_rng(31)
    rows, V, H, k = 128, 320, 200, 2
    m, orc = _machine(ctx, V, H, "f32")
    from keras_unsupervised_b200.engine import Machine

    chains = _data(rng, rows, V, 0.5)
    orc.chains = chains.copy()
    m.set_chains(chains)
    for step in range(2):
        v = _data(rng, rows, V, 0.2)
        u_h, u_v, u_hc = _inject(rng, rows, V, H, k, persistent=True)
        u_h, u_v, u_hc = O.condition_margin(orc, v, u_h, u_v, k=k, persistent=True, u_hc=u_hc)
        st = orc.fused_step(v, u_h, u_v, lr=1e-3, k=k, persistent=True, u_hc=u_hc)
        m.cd_step(v, Machine.hparams(lr=1e-3, k=k, persistent=True), u_h=u_h, u_v=u_v, u_hc=u_hc)
        got = m.last_stats(rows)
        assert np.array_equal(got["v_neg"], st["v_neg"])
        assert np.array_equal(m.get_chains(rows), orc.chains)
        np.testing.assert_allclose(got["dW"], st["dW"], rtol=1e-5, atol=1e-4)


def test_momentum_weight_decay_mean(ctx):
    rng = np.random.default_rng(41)
    rows, V, H = 64, 192, 128
    m, orc = _machine(ctx, V, H, "f32")
    from keras_unsupervised_b200.engine import Machine

    for _ in range(3):
        v = _data(rng, rows, V)
        u_h, u_v, _ = _inject(rng, rows, V, H, 1)
        u_h, u_v, _ = O.condition_margin(orc, v, u_h, u_v, k=1)
        orc.fused_step(v, u_h, u_v, lr=0.05, momentum=0.5, weight_decay=1e-3, scale=1.0 / rows)
        m.cd_step(v, Machine.hparams(lr=0.05, momentum=0.5, weight_decay=1e-3, normalize=True), u_h=u_h, u_v=u_v)
    W, b, c = m.get_params()
    np.testing.assert_allclose(W, orc.W, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(b, orc.b, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(c, orc.c, rtol=1e-5, atol=1e-6)


def test_philox_stream_matches_oracle(ctx):
    """Without injection the engine draws Philox4x32-10 keyed by (seed, draw id, global row, column):
    the oracle regenerates the same uniforms, so even 'random' runs are comparable."""
    rng = np.random.default_rng(53)
    rows, V, H, seed = 300, 200, 333, 1234
    m, orc = _machine(ctx, V, H, "f32", seed=seed)
    v = _data(rng, rows, V)
    h, p = m.transform(v, want_p=True)          # first inference draw: id 2^63 + 0
    u = O.philox_uniform(seed, O.draw_id("infer", 0), 0, rows, H)
    near = np.abs(u - p) <= 4 * 2.0 ** -23
    assert np.array_equal(h[~near], (u < orc.prob_h(v)).astype(np.float32)[~near])
    assert near.mean() < 1e-3
    h2 = m.transform(v)                          # second call, next draw id: a different sample
    assert (h2 != h).mean() > 0.1
    # chunk / shard independence: rows 100.. drawn with row0 = 100 equal the tail of the full draw
    m.set_seed(seed, 0)
    full = m.transform(v)
    assert np.array_equal(full, h)


def test_fit_epoch_graph_replay_matches_oracle(ctx):
    """fit_epoch = CUDA-graph replay per minibatch with device-side offsets; includes a remainder step."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset, Machine

    rng = np.random.default_rng(61)
    N, B, V, H, seed = 1000, 128, 784, 500, 77
    m, orc = _machine(ctx, V, H, "f32", seed=seed)
    data = _data(rng, N, V, 0.13)
    ds = Dataset.from_array(ctx, data, L.COMPUTE_F32X3)
    assert ds.shape == (N, V)
    assert np.array_equal(ds.numpy(), data)
    hp = Machine.hparams(lr=1e-3, k=1)
    for _ in range(2):
        st = m.fit_epoch(ds, B, hp)
        assert st["steps"] == 8 and st["rows"] == N
    ctx.sync()
    O.philox_fit(orc, data, B, 2, 1e-3, seed)
    W, b, c = m.get_params()
    # a sample may flip where |u - p| is within rounding of the two sigmoid implementations; each
    # flip moves single entries by lr.  Demand near-equality in the mean and a bounded max.
    assert np.abs(W - orc.W).mean() < 2e-6
    assert np.abs(W - orc.W).max() < 5e-3
    assert np.abs(b - orc.b).max() < 5e-3 and np.abs(c - orc.c).max() < 5e-3
    ds.close()


@pytest.mark.parametrize("mode", ["f32", "bf16"])
@pytest.mark.parametrize("B,N", [(16, 88), (32, 176), (8, 44)])
def test_fit_epoch_tiny_minibatches_match_oracle(ctx, mode, B, N):
    """Minibatches far below one 128-row tile, with a half-size remainder (what a rank of tests/dp_check.py sees at 8 ranks:
    16 rows per step, 8 in the last one), CD-2, both compute modes: the replayed steps must follow the oracle."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset, Machine

    rng = np.random.default_rng(B)
    V, H, seed = 320, 256, 21
    m, orc = _machine(ctx, V, H, mode, seed=seed)
    data = _data(rng, N, V, 0.25)
    ds = Dataset.from_array(ctx, data, L.COMPUTE_F32X3 if mode == "f32" else L.COMPUTE_BF16)
    hp = Machine.hparams(lr=1e-3, k=2)
    for _ in range(2):
        m.fit_epoch(ds, B, hp, want_stats=False)
    ctx.sync()
    O.philox_fit(orc, data, B, 2, 1e-3, seed, k=2)
    W, b, c = m.get_params()
    assert np.abs(W - orc.W).mean() < 2e-6
    assert np.abs(W - orc.W).max() < 5e-3
    assert np.abs(b - orc.b).max() < 5e-3 and np.abs(c - orc.c).max() < 5e-3
    ds.close()


def test_dataset_transform_equals_array_transform(ctx):
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset

    rng = np.random.default_rng(67)
    rows, V, H, seed = 500, 256, 192, 5
    m, orc = _machine(ctx, V, H, "bf16", seed=seed)
    v = _data(rng, rows, V)
    ds = Dataset.from_array(ctx, v, L.COMPUTE_BF16)
    out = m.transform_dataset(ds)     # draw id 2^63 + 0
    m.set_seed(seed, 0)
    h = m.transform(v)                # same id again
    assert np.array_equal(out.numpy(), h)
    back = m.inv_transform_dataset(out)
    assert back.shape == (rows, V)


def test_errors_are_reported_not_fatal(ctx):
    m, _ = _machine(ctx, 128, 64, "f32")
    with pytest.raises(ValueError):
        m.transform(np.zeros((4, 100), np.float32))
    with pytest.raises(ValueError):
        m.free_energy(np.zeros((4, 64), np.float32))
    from keras_unsupervised_b200.engine import Machine

    with pytest.raises(ValueError):
        m.cd_step(np.zeros((4, 128), np.float32), Machine.hparams(k=0))
    with pytest.raises(ValueError):
        m.cd_step(np.zeros((4, 128), np.float32), Machine.hparams(persistent=True))
    # empty input: nothing to do, nothing breaks
    assert m.transform(np.zeros((0, 128), np.float32)).shape == (0, 64)
    m.cd_step(np.zeros((0, 128), np.float32), Machine.hparams())


def test_gaussian_visible_mode(ctx):
    """rbm.py:55-67,139-155: relu-threshold hiddens, unit-variance Gaussian visibles."""
    rng = np.random.default_rng(71)
    rows, V, H = 128, 256, 128
    m, orc = _machine(ctx, V, H, "f32", mode=1)
    from keras_unsupervised_b200.engine import Machine

    v = rng.normal(0, 1, (rows, V)).astype(np.float32)
    u_h = O.lattice_uniform(rng, (rows, H))
    pre = np.maximum(orc.pre_h(v), 0)
    bad = np.abs(u_h - pre) <= 1e-4
    u_h[bad] = 0.999
    n_v = rng.normal(0, 1, (rows, V)).astype(np.float32)
    h, p = m.transform(v, u=u_h, want_p=True)
    h_ref, p_ref = orc.sample_h(v, u_h)
    np.testing.assert_allclose(p, p_ref, rtol=1e-5, atol=1e-6)
    assert np.array_equal(h, h_ref)
    st = orc.fused_step(v, [u_h], [None, n_v], lr=1e-4)
    m.cd_step(v, Machine.hparams(lr=1e-4), u_h=[u_h], u_v=[None, n_v])
    got = m.last_stats(rows)
    np.testing.assert_allclose(got["v_neg"], st["v_neg"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got["dW"], st["dW"], rtol=1e-4, atol=1e-3)
    W, b, c = m.get_params()
    np.testing.assert_allclose(W, orc.W, rtol=1e-5, atol=1e-6)


def test_large_contraction_accuracy_f32(ctx):
    """K = 4096 real-valued inputs: the longest accumulation chain of the BASELINE configs."""
    rng = np.random.default_rng(83)
    rows, V, H = 256, 4096, 320
    m, orc = _machine(ctx, V, H, "f32")
    v = rng.random((rows, V)).astype(np.float32)
    p = m.transform(v, want_p=True)[1]
    p_ref = orc.prob_h(v)
    rel = np.abs(p - p_ref) / np.maximum(np.abs(p_ref), 1e-30)
    print("max rel err K=4096 real-valued:", rel.max())
    np.testing.assert_allclose(p, p_ref, rtol=RTOL_F32, atol=1e-7)
    vb = (v < 0.5).astype(np.float32)
    pb = m.transform(vb, want_p=True)[1]
    relb = np.abs(pb - orc.prob_h(vb)) / np.maximum(orc.prob_h(vb), 1e-30)
    print("max rel err K=4096 binary:", relb.max())
    np.testing.assert_allclose(pb, orc.prob_h(vb), rtol=RTOL_F32, atol=1e-7)


def test_split_minibatch_two_chain_schedule(monkeypatch):
    """Large minibatches run as two row-halves on two streams (their kernels fill each other's tail
    waves).  Forced here at small sizes: injected parity of one step, then graph replay with a remainder
    minibatch that leaves the second chain without valid rows."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Context, Dataset, Machine

    monkeypatch.setenv("KUCD_SPLIT", "2")
    ctx2 = Context(device=0, seed=1)
    rng = np.random.default_rng(91)
    rows, V, H, k = 300, 320, 200, 2
    for compute in ("f32", "bf16"):
        m, orc = _machine(ctx2, V, H, compute, seed=13)
        v = _data(rng, rows, V, 0.2)
        u_h, u_v, _ = _inject(rng, rows, V, H, k)
        u_h, u_v, _ = O.condition_margin(orc, v, u_h, u_v, k=k)
        st = orc.fused_step(v, u_h, u_v, lr=1e-3, k=k)
        m.cd_step(v, Machine.hparams(lr=1e-3, k=k), u_h=u_h, u_v=u_v)
        got = m.last_stats(rows)
        assert np.array_equal(got["h_pos"], st["h_pos"]) and np.array_equal(got["v_neg"], st["v_neg"])
        tol = RTOL_F32 if compute == "f32" else RTOL_BF16
        np.testing.assert_allclose(got["dW"], st["dW"], rtol=tol, atol=tol * max(1.0, np.abs(st["dW"]).max()))
        np.testing.assert_allclose(got["db"], st["db"], rtol=tol, atol=1e-6)
        W, b, c = m.get_params()
        np.testing.assert_allclose(W, orc.W, rtol=tol, atol=1e-6)

    N, B, seed = 1000, 300, 13
    m, orc = _machine(ctx2, V, H, "f32", seed=seed)
    data = _data(rng, N, V, 0.2)
    ds = Dataset.from_array(ctx2, data, L.COMPUTE_F32X3)
    for _ in range(2):
        m.fit_epoch(ds, B, Machine.hparams(lr=1e-3, k=k))
    ctx2.sync()
    O.philox_fit(orc, data, B, 2, 1e-3, seed, k=k)
    W, b, c = m.get_params()
    assert np.abs(W - orc.W).mean() < 2e-6 and np.abs(W - orc.W).max() < 5e-3
    assert np.abs(b - orc.b).max() < 5e-3 and np.abs(c - orc.c).max() < 5e-3
    ctx2.close()


def test_fit_host_streaming_equals_resident_fit(ctx):
    """fit_host (minibatch i+1 copied while minibatch i runs) and fit_epoch (resident data set, graph
    replay) draw the same Philox stream and must train the same model; per-step statistics come back."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset, Machine

    rng = np.random.default_rng(97)
    N, B, V, H, seed = 1000, 128, 300, 200, 3
    data = _data(rng, N, V, 0.2)
    hp = Machine.hparams(lr=1e-3, k=2)
    a, _ = _machine(ctx, V, H, "bf16", seed=seed)
    b, orc = _machine(ctx, V, H, "bf16", seed=seed)
    ds = Dataset.from_array(ctx, data, L.COMPUTE_BF16)
    a.fit_epoch(ds, B, hp)
    before = ctx.timings(reset=True)
    st = b.fit_host(data, B, hp)
    t = ctx.timings()
    assert st["steps"] == 8 and st["step_recon_err"].shape == (8,)
    assert np.all(st["step_recon_err"] > 0) and np.all(st["step_recon_err"] < 1)
    assert t["h2d_bytes"] == data.nbytes and t["d2h_bytes"] == 8 * 4
    Wa, ba, ca = a.get_params()
    Wb, bb, cb = b.get_params()
    np.testing.assert_allclose(Wa, Wb, rtol=0, atol=2e-6)
    np.testing.assert_allclose(ba, bb, rtol=0, atol=2e-6)
    O.philox_fit(orc, data, B, 1, 1e-3, seed, k=2)
    # bf16 mode: probabilities agree to ~1e-6, so over 1 M draws a couple of samples sit inside the rounding
    # gap and flip; each flip moves one row/column of W by lr
    assert np.abs(Wb - orc.W).mean() < 5e-5 and np.abs(Wb - orc.W).max() < 5e-3
    ds.close()


def test_whole_chain_kernel_equals_per_projection_launches(monkeypatch):
    """The persistent chain kernel (all 2k+1 projections of a minibatch in one launch, row-block dataflow
    between stages) must reproduce the launch-per-projection path bit for bit: same Philox draws, same
    arithmetic.  Forced here at small, ragged sizes; CD-3, PCD, and graph replay with a remainder minibatch."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Context, Dataset, Machine

    monkeypatch.setenv("KUCD_CHAIN", "2")
    monkeypatch.setenv("KUCD_CHAIN_DW", "1")   # also carry dW inside the chain kernel (off by default)
    c_chain = Context(device=0, seed=1)
    monkeypatch.setenv("KUCD_CHAIN", "0")
    c_plain = Context(device=0, seed=1)
    rng = np.random.default_rng(101)
    V, H, seed = 600, 520, 17
    ms = []
    for c in (c_chain, c_plain):
        m, orc = _machine(c, V, H, "bf16", seed=seed)
        ms.append(m)
    # one step, CD-3, 700 rows (three 256-row blocks, the last ragged)
    v = _data(rng, 700, V, 0.2)
    hp = Machine.hparams(lr=1e-3, k=3)
    got = []
    for m in ms:
        m.cd_step(v, hp)
        got.append(m.last_stats(700))
    t = c_chain.timings()
    assert t["chain_launches"] == 1 and t["chain_dw_launches"] == 1 and c_plain.timings()["chain_launches"] == 0
    for key in ("h_pos", "v_neg", "h_neg", "dW", "db"):
        assert np.array_equal(got[0][key], got[1][key]), key
    np.testing.assert_allclose(got[0]["dc"], got[1]["dc"], rtol=0, atol=1e-3)
    # against the oracle's regeneration of the same Philox stream
    u_h = [O.philox_uniform(seed, O.draw_id("train", 0, 0 if t_ == 0 else 2 * t_ + 1), 0, 700, H) for t_ in range(3)]
    u_v = [None] + [O.philox_uniform(seed, O.draw_id("train", 0, 2 * t_), 0, 700, V) for t_ in range(1, 4)]
    st = orc.cd_stats(v, u_h, u_v, k=3)
    assert (got[0]["h_pos"] != st["h_pos"]).mean() < 1e-4 and (got[0]["v_neg"] != st["v_neg"]).mean() < 1e-3

    # persistent chains + graph replay (1000 rows in minibatches of 300: remainder of 100)
    chains = _data(rng, 300, V, 0.5)
    data = _data(rng, 1000, V, 0.2)
    hp = Machine.hparams(lr=1e-3, k=2, persistent=True)
    params = []
    for c, m in zip((c_chain, c_plain), ms):
        m.set_chains(chains)
        ds = Dataset.from_array(c, data, L.COMPUTE_BF16)
        for _ in range(2):
            m.fit_epoch(ds, 300, hp)
        c.sync()
        params.append(m.get_params() + (m.get_chains(300),))
        ds.close()
    assert np.array_equal(params[0][3], params[1][3])                      # chains
    np.testing.assert_allclose(params[0][0], params[1][0], rtol=0, atol=1e-6)   # W (dc/db atomics order only)
    np.testing.assert_allclose(params[0][1], params[1][1], rtol=0, atol=1e-6)
    np.testing.assert_allclose(params[0][2], params[1][2], rtol=0, atol=1e-6)
    c_chain.close()
    c_plain.close()


def test_chain_kernel_dataflow_stress(monkeypatch):
    """200 CD-5 minibatches through the chain kernel (cross-CTA row-block flags, TMA reads of rows other CTAs just
    wrote) against the launch-per-projection path, statistics compared exactly at every step: a missed fence or a
    premature flag shows up as a differing dW."""
    from keras_unsupervised_b200.engine import Context, Machine

    monkeypatch.setenv("KUCD_CHAIN", "2")
    c_chain = Context(device=0, seed=1)
    monkeypatch.setenv("KUCD_CHAIN", "0")
    c_plain = Context(device=0, seed=1)
    rng = np.random.default_rng(5)
    V, H, B = 1024, 768, 2048
    a, _ = _machine(c_chain, V, H, "bf16", seed=3)
    b, _ = _machine(c_plain, V, H, "bf16", seed=3)
    hp = Machine.hparams(lr=1e-3, k=5, update_mask=0)
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    for step in range(200):
        x = (torch.rand((B, V), device="cuda", generator=g) < 0.3).to(torch.uint8)
        a.cd_step(x, hp)
        b.cd_step(x, hp)
        if step % 10 == 0 or step > 190:
            sa, sb = a.last_stats(B, states=False), b.last_stats(B, states=False)
            assert np.array_equal(sa["dW"], sb["dW"]) and np.array_equal(sa["db"], sb["db"]), step
    assert c_chain.timings()["chain_launches"] == 200
    c_chain.close()
    c_plain.close()


def test_sample_frequencies_follow_the_probabilities(ctx):
    """Philox mode: over 256 independent draws the empirical frequency of h_j = 1 matches sigmoid(v.W + c)."""
    rng = np.random.default_rng(13)
    rows, V, H, n = 64, 128, 96, 256
    m, orc = _machine(ctx, V, H, "f32", seed=99)
    W = rng.normal(0, 0.3, (V, H)).astype(np.float32)
    m.set_params(W=W)
    orc.W = W
    v = _data(rng, rows, V)
    p = orc.prob_h(v).astype(np.float64)
    freq = np.zeros_like(p)
    for _ in range(n):
        freq += m.transform(v)
    z = (freq - n * p) / np.sqrt(n * p * (1 - p) + 1e-12)
    assert np.abs(z).max() < 5.5 and abs(z.mean()) < 0.05 and 0.9 < z.std() < 1.1


def test_small_chain_variant_equals_per_projection_launches(monkeypatch):
    """Latency-bound sizes (minibatch <= 512): projections AND the dW contraction run as one launch of the 128 x 64-tile
    chain kernel.  Bit-identical states / dW / chains to the launch-per-contraction path."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Context, Dataset, Machine

    monkeypatch.delenv("KUCD_CHAIN", raising=False)
    c_small = Context(device=0, seed=1)
    monkeypatch.setenv("KUCD_CHAIN", "0")
    c_plain = Context(device=0, seed=1)
    rng = np.random.default_rng(103)
    for V, H, rows, k in ((784, 500, 128, 1), (333, 270, 300, 3)):
        ms = [_machine(c, V, H, "bf16", seed=23)[0] for c in (c_small, c_plain)]
        v = _data(rng, rows, V, 0.2)
        hp = Machine.hparams(lr=1e-3, k=k)
        got = []
        for m in ms:
            m.cd_step(v, hp)
            got.append(m.last_stats(rows))
        for key in ("h_pos", "v_neg", "h_neg", "dW", "db"):
            assert np.array_equal(got[0][key], got[1][key]), (V, key)
        np.testing.assert_allclose(got[0]["dc"], got[1]["dc"], rtol=0, atol=1e-3)
        chains = _data(rng, 128, V, 0.5)
        data = _data(rng, 600, V, 0.2)
        hp = Machine.hparams(lr=1e-3, k=2, persistent=True)
        params = []
        for c, m in zip((c_small, c_plain), ms):
            m.set_chains(chains)
            ds = Dataset.from_array(c, data, L.COMPUTE_BF16)
            for _ in range(2):
                m.fit_epoch(ds, 128, hp)          # 4 full minibatches + a remainder of 88 rows
            c.sync()
            params.append(m.get_params() + (m.get_chains(128),))
            ds.close()
        assert np.array_equal(params[0][3], params[1][3])
        for i in range(3):
            np.testing.assert_allclose(params[0][i], params[1][i], rtol=0, atol=1e-6)
    assert c_small.timings()["chain_dw_launches"] > 0 and c_plain.timings()["chain_launches"] == 0
    c_small.close()
    c_plain.close()


def test_mean_normalisation_with_a_remainder_minibatch(ctx):
    """normalize = mean under graph replay: the last, shorter minibatch divides by ITS row count (device-side)."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset, Machine

    rng = np.random.default_rng(107)
    N, B, V, H, seed = 300, 128, 200, 136, 19     # minibatches of 128, 128 and 44 rows
    for compute in ("f32", "bf16"):
        m, orc = _machine(ctx, V, H, compute, seed=seed)
        data = _data(rng, N, V, 0.3)
        ds = Dataset.from_array(ctx, data, L.COMPUTE_F32X3 if compute == "f32" else L.COMPUTE_BF16)
        hp = Machine.hparams(lr=0.1, k=1, normalize=True, momentum=0.5)
        for _ in range(2):
            m.fit_epoch(ds, B, hp)
        ctx.sync()
        O.philox_fit(orc, data, B, 2, 0.1, seed, normalize=True, momentum=0.5)
        W, b, c = m.get_params()
        assert np.abs(W - orc.W).mean() < 2e-5 and np.abs(W - orc.W).max() < 5e-3, compute
        assert np.abs(b - orc.b).max() < 5e-3 and np.abs(c - orc.c).max() < 5e-3
        ds.close()


def test_float32_grade_chain_kernels_equal_per_projection_launches(monkeypatch):
    """Float32-grade mode (the reference's own precision, rbm.py:39) through the chain kernels (chain.cuh, CH > 0: three
    term planes of W as K-segments, accumulation cut every 4 k-blocks, pieces added in registers): same draws, same
    piece boundaries, same order of additions as the launch-per-projection path - states, h_neg and dW bit for bit.
    Small-tile variant (784 -> 500, batch 128; CD-3 at ragged sizes) and the 256 x 256 CTA-pair variant (forced at a
    small size, CD-3, 700 rows), PCD with graph replay and a remainder minibatch; then against the oracle at 1e-5."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Context, Dataset, Machine

    rng = np.random.default_rng(131)
    for force, shapes in ((None, ((784, 500, 128, 1), (333, 270, 300, 3))), ("2", ((600, 520, 700, 3),))):
        if force is None:
            monkeypatch.delenv("KUCD_CHAIN", raising=False)
        else:
            monkeypatch.setenv("KUCD_CHAIN", force)
        c_chain = Context(device=0, seed=1)
        monkeypatch.setenv("KUCD_CHAIN", "0")
        c_plain = Context(device=0, seed=1)
        for V, H, rows, k in shapes:
            ms, orc = [], None
            for c in (c_chain, c_plain):
                m, orc = _machine(c, V, H, "f32", seed=37)
                ms.append(m)
            v = _data(rng, rows, V, 0.2)
            hp = Machine.hparams(lr=1e-3, k=k, update_mask=0)
            got = []
            for m in ms:
                m.cd_step(v, hp)
                got.append(m.last_stats(rows))
            for key in ("h_pos", "v_neg", "h_neg", "dW", "db"):
                assert np.array_equal(got[0][key], got[1][key]), (V, key)
            np.testing.assert_allclose(got[0]["dc"], got[1]["dc"], rtol=0, atol=1e-3)
            # against the oracle's regeneration of the same Philox stream, float32 grade
            u_h = [O.philox_uniform(37, O.draw_id("train", 0, 0 if t == 0 else 2 * t + 1), 0, rows, H) for t in range(k)]
            u_v = [None] + [O.philox_uniform(37, O.draw_id("train", 0, 2 * t), 0, rows, V) for t in range(1, k + 1)]
            st = orc.cd_stats(v, u_h, u_v, k=k)
            same = ~((got[0]["h_pos"] != st["h_pos"]).any(axis=1) | (got[0]["v_neg"] != st["v_neg"]).any(axis=1))
            assert not (~same & (st["row_margin"] > 2e-6)).any() and same.mean() > 0.9
            np.testing.assert_allclose(got[0]["h_neg"][same], st["h_neg"][same], rtol=RTOL_F32, atol=1e-7)
            # persistent chains + graph replay with a remainder minibatch
            b = min(rows, 256)
            chains = _data(rng, b, V, 0.5)
            data = _data(rng, 2 * b + 88, V, 0.2)
            hp = Machine.hparams(lr=1e-3, k=2, persistent=True)
            params = []
            for c, m in zip((c_chain, c_plain), ms):
                m.set_chains(chains)
                ds = Dataset.from_array(c, data, L.COMPUTE_F32X3)
                for _ in range(2):
                    m.fit_epoch(ds, b, hp)
                c.sync()
                params.append(m.get_params() + (m.get_chains(b),))
                ds.close()
            assert np.array_equal(params[0][3], params[1][3])
            for i in range(3):
                np.testing.assert_allclose(params[0][i], params[1][i], rtol=0, atol=1e-6)
        assert c_chain.timings()["chain_launches"] > 0 and c_plain.timings()["chain_launches"] == 0
        c_chain.close()
        c_plain.close()


def test_mid_chain_variant_equals_per_projection_launches(monkeypatch):
    """Minibatches too short for 74 pair tiles per stage but with >= 32 tiles of 128 x 256 (a 512-row shard of a
    4096-wide layer): chain_kernel<256, 1> - the flattened chain on single CTAs.  Bit-identical states and dW to the
    launch-per-projection path; CD-2 at a ragged row count, then graph replay with a remainder minibatch."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Context, Dataset, Machine

    monkeypatch.delenv("KUCD_CHAIN", raising=False)
    c_mid = Context(device=0, seed=1)
    monkeypatch.setenv("KUCD_CHAIN", "0")
    c_plain = Context(device=0, seed=1)
    rng = np.random.default_rng(137)
    V, H, rows = 2048, 2304, 600
    ms = [_machine(c, V, H, "bf16", seed=41)[0] for c in (c_mid, c_plain)]
    v = _data(rng, rows, V, 0.3)
    hp = Machine.hparams(lr=1e-3, k=2, update_mask=0)
    got = []
    for m in ms:
        m.cd_step(v, hp)
        got.append(m.last_stats(rows))
    assert c_mid.timings()["chain_launches"] == 1 and c_mid.timings()["chain_dw_launches"] == 0
    for key in ("h_pos", "v_neg", "h_neg", "dW", "db"):
        assert np.array_equal(got[0][key], got[1][key]), key
    np.testing.assert_allclose(got[0]["dc"], got[1]["dc"], rtol=0, atol=2e-3)
    data = _data(rng, 2 * 512 + 200, V, 0.3)
    hp = Machine.hparams(lr=1e-3, k=1)
    params = []
    for c, m in zip((c_mid, c_plain), ms):
        ds = Dataset.from_array(c, data, L.COMPUTE_BF16)
        m.fit_epoch(ds, 512, hp)
        c.sync()
        params.append(m.get_params())
        ds.close()
    for i in range(3):
        np.testing.assert_allclose(params[0][i], params[1][i], rtol=0, atol=1e-6)
    assert c_plain.timings()["chain_launches"] == 0
    c_mid.close()
    c_plain.close()
