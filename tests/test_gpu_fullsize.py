"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle would
take minutes there): determinism of the counter-based draws, additivity of the statistics over row
shards (the data-parallel identity), checksums of the returned states against the returned statistics,
chunk independence of inference, persistence of untouched chains."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(ctx, V, H, seed=5, compute="bf16", pseed=0):
    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Machine

    m = Machine(ctx, V, H, 0, L.COMPUTE_BF16 if compute == "bf16" else L.COMPUTE_F32X3, seed=seed)
    rng = np.random.default_rng(pseed)
    m.set_params(rng.uniform(-0.05, 0.05, (V, H)).astype(np.float32), rng.uniform(-0.05, 0.05, V).astype(np.float32),
                 rng.uniform(-0.05, 0.05, H).astype(np.float32))
    return m


def test_c3_cd10_step_properties(ctx):
    """RBM 4096 -> 4096, CD-10, batch 4096, bf16 (BASELINE.json configs[2])."""
    import torch

    from keras_unsupervised_b200.engine import Machine

    V = H = B = 4096
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    x = (torch.rand((B, V), device="cuda", generator=g) < 0.5).to(torch.uint8)
    hp0 = Machine.hparams(lr=1e-3, k=10, normalize=True, update_mask=0)   # statistics only
    a, b2 = _mk(ctx, V, H), _mk(ctx, V, H)
    a.cd_step(x, hp0)
    sa = a.last_stats(B)
    b2.cd_step(x, hp0)
    sb = b2.last_stats(B, states=False)
    # determinism: same seed, same step counter -> same chain, bit for bit
    assert np.array_equal(sa["dW"], sb["dW"])
    assert np.array_equal(sa["db"], sb["db"])
    # states are binary; the bias statistics are the column checksums of the returned states
    xs = x.float().cpu().numpy()
    for key in ("h_pos", "v_neg"):
        assert set(np.unique(sa[key])) <= {0.0, 1.0}
    assert 0 < sa["h_neg"].min() and sa["h_neg"].max() <= 1   # bf16 probabilities may round to 1.0
    np.testing.assert_array_equal(sa["db"], xs.sum(0) - sa["v_neg"].sum(0))
    np.testing.assert_allclose(sa["dc"], sa["h_pos"].sum(0) - sa["h_neg"].astype(np.float64).sum(0), rtol=0, atol=0.3)
    # dW is the difference of the two outer products of the returned states (bf16 h_neg, fp32 accumulation)
    ref = xs[:, :64].T @ sa["h_pos"][:, :64] - sa["v_neg"][:, :64].T.astype(np.float64) @ sa["h_neg"][:, :64]
    np.testing.assert_allclose(sa["dW"][:64, :64], ref, rtol=0, atol=2e-2)
    # additivity over row shards with global-row draws: what data parallelism relies on
    lo, hi = _mk(ctx, V, H), _mk(ctx, V, H)
    lo.cd_step(x[:1024], hp0, global_row0=0)
    hi.cd_step(x[1024:], hp0, global_row0=1024)
    s_lo, s_hi = lo.last_stats(1024), hi.last_stats(B - 1024)
    assert np.array_equal(np.concatenate([s_lo["v_neg"], s_hi["v_neg"]]), sa["v_neg"])
    np.testing.assert_allclose(s_lo["dW"] + s_hi["dW"], sa["dW"], rtol=0, atol=1e-2)
    np.testing.assert_array_equal(s_lo["db"] + s_hi["db"], sa["db"])
    # the next step draws differently
    a.cd_step(x, hp0)
    assert not np.array_equal(a.last_stats(B, states=False)["db"], sa["db"])


def test_c5_inference_is_chunk_independent(ctx):
    """4096 -> 4096 transform over 70 000 rows: the array entry point (32 768-row chunks, host round trip)
    and the resident data-set entry point (one launch) return the same states; free energy likewise."""
    import torch

    from keras_unsupervised_b200 import _lib as L
    from keras_unsupervised_b200.engine import Dataset

    V = H = 4096
    n = 70000
    m = _mk(ctx, V, H, seed=8)
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    x = (torch.rand((n, V), device="cuda", generator=g) < 0.5).to(torch.uint8)
    h_chunks = m.transform(x, out_dtype=torch.uint8)            # draw id 2^63 + 0
    ds = Dataset.from_array(ctx, x, L.COMPUTE_BF16)
    m.set_seed(8, 0)
    out = m.transform_dataset(ds)                                # same draw id, one launch
    h_one = torch.from_numpy(out.numpy()).to(torch.uint8)
    assert torch.equal(h_chunks.cpu(), h_one)
    assert 0.3 < h_one.float().mean().item() < 0.7
    fe = m.free_energy(x)
    fe_tail = m.free_energy(x[n - 1000:])
    assert fe.shape == (n,) and torch.isfinite(fe).all()
    torch.testing.assert_close(fe[n - 1000:], fe_tail, rtol=1e-6, atol=1e-3)
    out.close()
    ds.close()


def test_c4_pcd_step_properties(ctx):
    """PCD RBM 16384 -> 8192, one GPU's share of BASELINE.json configs[3]: 1024 rows, 2048 stored chains."""
    import torch

    from keras_unsupervised_b200.engine import Machine

    V, H, B, C = 16384, 8192, 1024, 2048
    m = _mk(ctx, V, H, seed=9)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    x = (torch.rand((B, V), device="cuda", generator=g) < 0.5).to(torch.uint8)
    chains = (torch.rand((C, V), device="cuda", generator=g) < 0.5).to(torch.uint8)
    m.set_chains(chains)
    W0 = m.get_params()[0]
    m.cd_step(x, Machine.hparams(lr=1e-3, k=1, persistent=True, normalize=True))
    st = m.last_stats(B)
    after = m.get_chains(C)
    before = chains.float().cpu().numpy()
    assert np.array_equal(after[:B], st["v_neg"])              # the chain advanced to the new v_neg
    assert np.array_equal(after[B:], before[B:])               # chains beyond the minibatch are untouched
    assert not np.array_equal(after[:B], before[:B])
    W1 = m.get_params()[0]
    np.testing.assert_allclose(W1 - W0, 1e-3 / B * st["dW"], rtol=0, atol=1e-6)
