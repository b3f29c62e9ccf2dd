"""GPU tests of the data formats either side of the CD path through the C ABI: bit-packed 0/1
matrices in and out, the per-epoch shuffle, and the Gaussian-visible chain kernels.  Bit-exact bars: packed and dense
forms of the same data must give the same states and the same parameters; a shuffled data set must be the oracle's
permutation of its rows."""
import numpy as np
import pytest

from oracle import cd_oracle as O

pytestmark = pytest.mark.gpu


def _machine(ctx, V, H, compute="bf16", mode=0, seed=0, pseed=0):
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Machine

    m = Machine(ctx, V, H, mode, L.COMPUTE_F32X3 if compute == "f32" else L.COMPUTE_BF16, seed=seed)
    W, b, c = O.OracleRBM.init_params(V, H, seed=pseed)
    m.set_params(W, b, c)
    return m


def _data(rng, rows, V, q=0.3):
    return (rng.random((rows, V)) < q).astype(np.float32)


@pytest.mark.parametrize("rows,V", [(100, 784), (1, 8), (37, 130), (300, 77), (65, 1)])
def test_packed_data_set_equals_dense_data_set(ctx, rows, V):
    """ingest_bits_kernel / export_bits_kernel against the oracle's unpack / pack, ragged column counts included."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.data import PackedBits
    from keras_unsupervised_b200.engine import Dataset

    rng = np.random.default_rng(rows * 1000 + V)
    x = _data(rng, rows, V, 0.5)
    p = PackedBits(O.pack_bits(x), V)
    ds = Dataset.from_array(ctx, p, L.COMPUTE_BF16)
    assert ds.shape == (rows, V)
    assert np.array_equal(ds.numpy(), x)
    back = ds.packed()
    assert np.array_equal(back.data, O.pack_bits(x))
    ds.close()
    dense = Dataset.from_array(ctx, x, L.COMPUTE_F32X3)                      # float32 in, bits out
    assert np.array_equal(dense.packed().data, O.pack_bits(x))
    dense.close()


@pytest.mark.parametrize("V", [784, 130, 8, 5])
def test_uint8_and_bool_ingest(ctx, V):
    """ingest_kernel<uint8_t>: the eight-bytes-per-load path (rows of 784 start 8-byte aligned) and the byte path
    (rows of 130 or 5 do not), bool arrays, and a row-strided view."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset

    rng = np.random.default_rng(V)
    x = rng.integers(0, 2, (203, V), dtype=np.uint8)
    for form in (x, x.astype(bool), np.ascontiguousarray(np.pad(x, ((0, 0), (0, 3))))[:, :V]):
        ds = Dataset.from_array(ctx, form, L.COMPUTE_BF16)
        assert np.array_equal(ds.numpy(), x.astype(np.float32))
        ds.close()
    big = rng.integers(0, 256, (64, V), dtype=np.uint8)                      # any byte value, exactly (<= 8 bits)
    ds = Dataset.from_array(ctx, big, L.COMPUTE_BF16)
    assert np.array_equal(ds.numpy(), big.astype(np.float32))
    ds.close()


def test_packed_rows_with_a_pitch(ctx):
    """strides[0] is the row pitch in bits: rows cut out of a wider byte matrix."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.data import PackedBits
    from keras_unsupervised_b200.engine import Dataset

    rng = np.random.default_rng(5)
    x = _data(rng, 50, 130, 0.5)
    wide = rng.integers(0, 256, (50, 40), dtype=np.uint8)
    wide[:, :17] = O.pack_bits(x)
    ds = Dataset.from_array(ctx, PackedBits(wide[:, :17], 130), L.COMPUTE_BF16)
    assert np.array_equal(ds.numpy(), x)                                     # junk bits beyond column 130 are masked
    ds.close()


@pytest.mark.parametrize("compute", ["bf16", "f32"])
def test_packed_input_and_output_of_transform(ctx, compute):
    """transform / inv_transform with packed input and out_dtype='bits' give the states of the dense call (same Philox
    draws: the seed and the inference draw counter are reset before each call)."""
    from keras_unsupervised_b200.data import PackedBits

    rng = np.random.default_rng(7)
    rows, V, H = 300, 333, 130
    m = _machine(ctx, V, H, compute, seed=3)
    v = _data(rng, rows, V)
    m.set_seed(3, 0)
    h_dense = m.transform(v)
    m.set_seed(3, 0)
    h_bits = m.transform(PackedBits.from_dense(v), out_dtype="bits")
    assert isinstance(h_bits, PackedBits) and h_bits.shape == (rows, H)
    assert np.array_equal(h_bits.to_dense(), h_dense)
    m.set_seed(3, 0)
    v_dense = m.inv_transform(h_dense)
    m.set_seed(3, 0)
    v_bits = m.inv_transform(h_bits, out_dtype="bits")
    assert np.array_equal(v_bits.to_dense(), v_dense)
    # the softplus row sums are accumulated across column tiles with atomics: equal up to the order of fp32 adds
    np.testing.assert_allclose(m.free_energy(PackedBits.from_dense(v)), m.free_energy(v), rtol=2e-6, atol=0)
    with pytest.raises(ValueError):                                          # probabilities are not bits
        keep = []
        from keras_unsupervised_b200 import _lib as L
        import ctypes as C
        tin = L.tensor_of(v, keep)
        out = PackedBits(np.zeros((rows, (H + 7) // 8), np.uint8), H)
        L.check(ctx.lib.kucd_rbm_transform(m.handle, C.byref(tin), None, C.byref(L.tensor_of(out, keep)), None))
    m.close()


def test_training_on_packed_data_equals_training_on_float32(ctx):
    """cd_step, fit_epoch (resident) and fit_host (streamed, 1/32 of the bytes) on packed data leave exactly the
    parameters the float32 form of the same data leaves."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.data import PackedBits
    from keras_unsupervised_b200.engine import Dataset, Machine

    rng = np.random.default_rng(11)
    V, H, N, B = 200, 96, 1000, 128                                          # 7 full minibatches + 104 rows
    x = _data(rng, N, V, 0.25)
    p = PackedBits.from_dense(x)
    hp = Machine.hparams(lr=1e-3, k=2)
    results = []
    for form in (x, p):
        m = _machine(ctx, V, H, "bf16", seed=9)
        m.cd_step(form[:B], hp)
        ds = Dataset.from_array(ctx, form, L.COMPUTE_BF16)
        m.fit_epoch(ds, B, hp)
        ds.close()
        t0 = ctx.timings()["h2d_bytes"]
        st = m.fit_host(form, B, hp)
        moved = ctx.timings()["h2d_bytes"] - t0
        results.append((m.get_params(), st["step_recon_err"].copy(), moved))
        m.close()
    (Wa, ba, ca), ra, moved_f32 = results[0]
    (Wb, bb, cb), rb, moved_bits = results[1]
    assert np.array_equal(ra, rb)
    np.testing.assert_allclose(Wa, Wb, rtol=0, atol=1e-6)                    # dc / db accumulate with atomics
    np.testing.assert_allclose(ba, bb, rtol=0, atol=1e-6)
    np.testing.assert_allclose(ca, cb, rtol=0, atol=1e-6)
    assert moved_f32 == N * V * 4 and moved_bits == N * (V // 8)


@pytest.mark.parametrize("compute,real", [("bf16", False), ("f32", True)])
def test_shuffled_data_set_is_the_oracle_permutation(ctx, compute, real):
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset

    rng = np.random.default_rng(13)
    N, V = 1000, 130
    x = rng.random((N, V)).astype(np.float32) if real else _data(rng, N, V)   # real values: three term planes
    ds = Dataset.from_array(ctx, x, L.COMPUTE_F32X3 if compute == "f32" else L.COMPUTE_BF16)
    base = ds.numpy()
    np.testing.assert_allclose(base, x, rtol=1e-6, atol=0)
    s0 = ds.shuffled(seed=42, epoch=0)
    assert np.array_equal(s0.numpy(), base[O.feistel_permutation(N, 42, 0)])
    s1 = ds.shuffled(seed=42, epoch=1, into=s0)                              # reuse the buffer
    assert s1 is s0 and np.array_equal(s0.numpy(), base[O.feistel_permutation(N, 42, 1)])
    assert np.array_equal(ds.numpy(), base)                                  # the source is untouched
    with pytest.raises(ValueError):
        ds.shuffled(seed=1, epoch=0, into=ds)
    other = Dataset.from_array(ctx, x[:10], L.COMPUTE_BF16)
    with pytest.raises(ValueError):
        ds.shuffled(seed=1, epoch=0, into=other)
    for d in (other, s0, ds):
        d.close()


def test_fit_with_shuffle_equals_fit_on_the_permuted_rows(ctx):
    """hps['shuffle']: epoch e trains on rows perm_e of the data.  Replayed by the oracle (same Philox stream) on the
    explicitly permuted rows, and compared with an engine run that is fed the permuted rows epoch by epoch."""
    from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI

    rng = np.random.default_rng(17)
    N, V, H, B, seed = 600, 160, 64, 128, 31
    x = _data(rng, N, V, 0.3)
    hps = {"batch_size": B, "epochs": 2, "lr": 1e-3, "dtype": "float32", "seed": seed, "shuffle": True}
    a = RBM(dict(hps), H, name="a", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    a.build((None, V))
    W0, b0, c0 = a._machine.get_params()
    a.fit(x, verbose=0)

    b = RBM(dict(hps, shuffle=False, epochs=1, stream=False), H, name="b", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    b.build((None, V))
    orc = O.OracleRBM(W0, b0, c0)
    steps = (N + B - 1) // B
    for e in range(2):
        xe = x[O.feistel_permutation(N, seed, e)]
        b.fit(xe, verbose=0)
        O.philox_fit(orc, xe, B, 1, 1e-3, seed, step0=e * steps)
    np.testing.assert_allclose(a.rbm_weight, b.rbm_weight, rtol=0, atol=1e-6)
    np.testing.assert_allclose(a.hidden_bias, b.hidden_bias, rtol=0, atol=1e-6)
    assert np.abs(a.rbm_weight - orc.W).mean() < 2e-6 and np.abs(a.rbm_weight - orc.W).max() < 5e-3
    # and it is not the unshuffled fit
    c = RBM(dict(hps, shuffle=False), H, name="c", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
    c.fit(x, verbose=0)
    assert np.abs(c.rbm_weight - a.rbm_weight).max() > 1e-4


def test_gaussian_visible_chain_kernels_equal_per_projection_launches(monkeypatch):
    """Gaussian-visible mode in bf16 now runs through the chain kernels too (chain_kernel<.., GAUSS = true>: relu-threshold
    hiddens, v = mean + Box-Muller normal, final h the sigmoid).  Same Philox draws, same epilogue code: states and dW
    must equal the launch-per-projection path bit for bit - small-tile variant (projections + dW in one launch) and
    the 256 x 256 CTA-pair variant (forced at a small size), CD-1 and CD-3, ragged shapes, graph replay with a remainder."""
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Context, Dataset, Machine

    rng = np.random.default_rng(19)
    for force, (V, H, rows, k) in ((None, (200, 96, 128, 1)), (None, (333, 270, 300, 3)), ("2", (600, 520, 700, 2))):
        if force is None:
            monkeypatch.delenv("KUCD_CHAIN", raising=False)
        else:
            monkeypatch.setenv("KUCD_CHAIN", force)
        c_chain = Context(device=0, seed=1)
        monkeypatch.setenv("KUCD_CHAIN", "0")
        c_plain = Context(device=0, seed=1)
        ms = [_machine(c, V, H, "bf16", mode=1, seed=29) for c in (c_chain, c_plain)]
        v = rng.normal(0, 1, (rows, V)).astype(np.float32)
        hp = Machine.hparams(lr=1e-4, k=k)
        got = []
        for m in ms:
            m.cd_step(v, hp)
            got.append(m.last_stats(rows))
        assert c_chain.timings()["chain_launches"] == 1 and c_plain.timings()["chain_launches"] == 0
        for key in ("h_pos", "v_neg", "h_neg"):
            assert np.array_equal(got[0][key], got[1][key]), (V, key)
        np.testing.assert_allclose(got[0]["dW"], got[1]["dW"], rtol=0, atol=1e-3)   # real-valued operands
        np.testing.assert_allclose(got[0]["db"], got[1]["db"], rtol=0, atol=2e-3)   # real-valued sums, atomics order
        np.testing.assert_allclose(got[0]["dc"], got[1]["dc"], rtol=0, atol=2e-3)
        # the sampled visibles really are mean + unit normal of the engine's stream (oracle regeneration)
        n1 = O.philox_normal(29, O.draw_id("train", 0, 2 * k), 0, rows, V)
        assert abs(float(np.corrcoef((got[0]["v_neg"]).ravel(), n1.ravel())[0, 1])) > 0.9
        data = rng.normal(0, 1, (5 * rows // 2, V)).astype(np.float32)          # 2 full minibatches + half a one
        hp_fit = Machine.hparams(lr=1e-3, k=k, normalize=True)   # batch means: a sum-updated Gaussian RBM diverges here
        params = []
        for c, m in zip((c_chain, c_plain), ms):
            ds = Dataset.from_array(c, data, L.COMPUTE_BF16)
            for _ in range(2):
                m.fit_epoch(ds, rows, hp_fit)
            c.sync()
            params.append(m.get_params())
            ds.close()
        for i in range(3):
            np.testing.assert_allclose(params[0][i], params[1][i], rtol=1e-5, atol=1e-6)
        assert np.isfinite(params[0][0]).all() and np.abs(params[0][0]).max() < 1.0
        c_chain.close()
        c_plain.close()
