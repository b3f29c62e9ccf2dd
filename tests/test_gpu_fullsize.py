"""BASELINE.json's full sizes on the GPU.

Against the oracle (rows of a minibatch are independent given W, b, c - rbm.py:119-124 has no cross-row term - so the
oracle replays the engine's own counter-based draws for row slices of the production-size step, and whole minibatches
where that costs seconds): the CD-10 chain of C3 state by state, dW / db / dc of whole minibatches, C4's PCD share,
C5's transform and free energy.  A draw closer to its probability than two sigmoid implementations agree (a few 1e-6)
may legitimately flip and then the rest of that row's chain differs: rows whose smallest |u - p| over the whole chain
is above TIE are demanded bit for bit, every differing row must be explained by such a near-tie, and their number is
bounded.

And through size-independent properties: determinism of the counter-based draws, additivity of the statistics over
row shards (the data-parallel identity), checksums of the returned states against the returned statistics, chunk
independence of inference, persistence of untouched chains."""
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


# ------------------------------------------------------------------------------------------------------------------
# the same sizes against the oracle
# ------------------------------------------------------------------------------------------------------------------
TIE = 1e-5   # a draw this close to its probability may fall either way (bf16 mode: fp32 accumulation order, ex2/rcp)


def _oracle(m, compute="bf16"):
    from oracle import cd_oracle as O

    W, b, c = m.get_params()
    return O.OracleRBM(W, b, c, compute=compute)


def _train_draws(seed, step, row0, rows, V, H, k, pcd=False):
    """The uniforms the engine draws for rows [row0, row0 + rows) of training step `step` (include/kucd.h draw ids)."""
    from oracle import cd_oracle as O

    u_h = [O.philox_uniform(seed, O.draw_id("train", step, 0 if t == 0 else 2 * t + 1), row0, rows, H) for t in range(k)]
    u_v = [None] + [O.philox_uniform(seed, O.draw_id("train", step, 2 * t), row0, rows, V) for t in range(1, k + 1)]
    u_hc = O.philox_uniform(seed, O.draw_id("train", step, 1), row0, rows, H) if pcd else None
    return u_h, u_v, u_hc


def _check_rows(got, st, lo, hi, max_flagged):
    """Engine states of minibatch rows [lo, hi) against the oracle's replay `st` of those rows."""
    clean = st["row_margin"] > TIE
    differs = np.zeros(hi - lo, bool)
    for key in ("h_pos", "v_neg"):
        differs |= (got[key][lo:hi] != st[key]).any(axis=1)
    assert not (differs & clean).any(), "a row with no near-tie differs from the oracle"
    assert differs.sum() <= max_flagged * (hi - lo), (int(differs.sum()), int((~clean).sum()), hi - lo)
    same = ~differs
    # the final hidden term is a probability (rbm.py:124), delivered as bf16
    np.testing.assert_allclose(got["h_neg"][lo:hi][same], st["h_neg"][same], rtol=2e-2, atol=1e-6)
    return same


def _check_grads(got, st, same, x_rows, rows):
    """dW / db / dc of a whole minibatch against the oracle's.  Rows that took the other side of a near-tie are moved
    from one side to the other first: their exact contribution, from the engine's own returned states, replaces the
    oracle's contribution of those rows."""
    from oracle import cd_oracle as O

    f = np.float64
    dW, db, dc = st["dW"].astype(f), st["db"].astype(f), st["dc"].astype(f)
    bad = ~same
    if bad.any():
        hn_o, hn_e = O.bf16_round(st["h_neg"][bad]).astype(f), got["h_neg"][bad].astype(f)
        dW += -(x_rows[bad].T.astype(f) @ st["h_pos"][bad].astype(f) - st["v_neg"][bad].T.astype(f) @ hn_o)
        dW += x_rows[bad].T.astype(f) @ got["h_pos"][bad].astype(f) - got["v_neg"][bad].T.astype(f) @ hn_e
        db += st["v_neg"][bad].sum(0) - got["v_neg"][bad].sum(0)
        dc += -(st["h_pos"][bad].sum(0) - st["h_neg"][bad].astype(f).sum(0)) + \
            (got["h_pos"][bad].sum(0) - got["h_neg"][bad].astype(f).sum(0))
    # entries are differences of two sums of `rows` terms (~ rows / 4 each): 2e-2 relative (north_star's bf16 bar) with
    # an absolute floor for the entries that cancel to ~0
    np.testing.assert_allclose(got["dW"], dW, rtol=2e-2, atol=0.1)
    np.testing.assert_array_equal(got["db"], db.astype(np.float32))      # binary states: integer counts, exact
    np.testing.assert_allclose(got["dc"], dc, rtol=2e-2, atol=0.3)       # (engine sums the bf16-rounded flagged rows)


def test_c3_cd10_chain_matches_oracle(ctx):
    """RBM 4096 -> 4096, CD-10, batch 4096, bf16 (BASELINE.json configs[2]) - the production chain_kernel<256, 2>
    launch (74 CTA pairs, 5376 tiles, cross-CTA row-block flags).  The oracle replays the whole ten-step chain for four
    256-row slices of the minibatch (first / interior, straddling a 256-row tile boundary / last rows)."""
    import torch

    from keras_unsupervised_b200.engine import Machine

    V = H = B = 4096
    seed, k = 21, 10
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    x = (torch.rand((B, V), device="cuda", generator=g) < 0.5).to(torch.uint8)
    m = _mk(ctx, V, H, seed=seed, pseed=3)
    ctx.timings(reset=True)
    m.cd_step(x, Machine.hparams(lr=1e-3, k=k, update_mask=0))
    assert ctx.timings()["chain_launches"] == 1      # the whole-chain kernel ran, not the per-projection fallback
    got = m.last_stats(B)
    xs = x.cpu().numpy().astype(np.float32)
    orc = _oracle(m)
    n_same = 0
    for lo in (0, 1152, 2944, 3840):
        hi = lo + 256
        u_h, u_v, _ = _train_draws(seed, 0, lo, 256, V, H, k)
        st = orc.cd_stats(xs[lo:hi], u_h, u_v, k=k)
        n_same += _check_rows(got, st, lo, hi, max_flagged=0.35).sum()
    assert n_same >= 0.65 * 1024
    # dW of the whole minibatch = the two outer products of the returned states (exact in float64; bf16 h_neg operand)
    f = np.float32
    ref = xs.T @ got["h_pos"] - got["v_neg"].T.astype(f) @ got["h_neg"].astype(f)
    np.testing.assert_allclose(got["dW"], ref, rtol=1e-3, atol=5e-2)
    np.testing.assert_array_equal(got["db"], xs.sum(0) - got["v_neg"].sum(0))


def test_c3_whole_minibatch_statistics_match_oracle(ctx):
    """Same shape, whole minibatch through the oracle (CD-1: three projections + the two outer products of 4096 rows),
    dW / db / dc at north_star's bf16 tolerance; then the update W += lr dW (rbm.py:127-134)."""
    import torch

    from keras_unsupervised_b200.engine import Machine

    V = H = B = 4096
    seed = 22
    g = torch.Generator(device="cuda")
    g.manual_seed(12)
    x = (torch.rand((B, V), device="cuda", generator=g) < 0.5).to(torch.uint8)
    m = _mk(ctx, V, H, seed=seed, pseed=4)
    orc = _oracle(m)
    m.cd_step(x, Machine.hparams(lr=1e-3 / B, k=1))
    got = m.last_stats(B)
    xs = x.cpu().numpy().astype(np.float32)
    u_h, u_v, _ = _train_draws(seed, 0, 0, B, V, H, 1)
    st = orc.cd_stats(xs, u_h, u_v, k=1)
    same = _check_rows(got, st, 0, B, max_flagged=0.1)
    _check_grads(got, st, same, xs, B)
    W1, b1, c1 = m.get_params()
    np.testing.assert_allclose(W1, orc.W + np.float32(1e-3 / B) * got["dW"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(b1, orc.b + np.float32(1e-3 / B) * got["db"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(c1, orc.c + np.float32(1e-3 / B) * got["dc"], rtol=0, atol=1e-7)


def test_c4_pcd_share_matches_oracle(ctx):
    """PCD RBM 16384 -> 8192, one GPU's share of BASELINE.json configs[3] (1024 rows + 1024 stored chains of the 8192),
    at its place in the global minibatch (global_row0 = 3072, i.e. rank 3 of 8): states, the advanced chains and
    dW / db / dc against the oracle."""
    import torch

    from keras_unsupervised_b200.engine import Machine

    V, H, B, row0 = 16384, 8192, 1024, 3072
    seed = 23
    m = _mk(ctx, V, H, seed=seed, pseed=5)
    g = torch.Generator(device="cuda")
    g.manual_seed(13)
    x = (torch.rand((B, V), device="cuda", generator=g) < 0.5).to(torch.uint8)
    chains = (torch.rand((B, V), device="cuda", generator=g) < 0.5).to(torch.uint8)
    m.set_chains(chains)
    m.cd_step(x, Machine.hparams(lr=1e-3, k=1, persistent=True, update_mask=0), global_row0=row0)
    got = m.last_stats(B)
    after = m.get_chains(B)
    xs = x.cpu().numpy().astype(np.float32)
    orc = _oracle(m)
    orc.chains = chains.cpu().numpy().astype(np.float32)
    u_h, u_v, u_hc = _train_draws(seed, 0, row0, B, V, H, 1, pcd=True)
    st = orc.cd_stats(xs, u_h, u_v, k=1, persistent=True, u_hc=u_hc)
    same = _check_rows(got, st, 0, B, max_flagged=0.15)
    assert np.array_equal(after[same], orc.chains[:B][same])     # rbm.py has no PCD: the oracle's extension semantics
    assert np.array_equal(after, got["v_neg"])
    _check_grads(got, st, same, xs, B)


@pytest.mark.parametrize("compute", ["bf16", "f32"])
def test_c5_inference_matches_oracle(ctx, compute):
    """4096 -> 4096 transform, inv_transform and free energy of 4096 rows (BASELINE.json configs[4]) against the oracle:
    states equal except where the draw is within rounding of the probability, probabilities and free energies at
    north_star's tolerance for the precision."""
    from oracle import cd_oracle as O

    V = H = n = 4096
    seed = 24
    m = _mk(ctx, V, H, seed=seed, compute=compute, pseed=6)
    orc = _oracle(m, "bf16" if compute == "bf16" else "f64")
    rng = np.random.default_rng(14)
    v = (rng.random((n, V)) < 0.5).astype(np.float32)
    tol = 2e-2 if compute == "bf16" else 1e-5
    tie = TIE if compute == "bf16" else 2e-6
    h, p = m.transform(v, want_p=True)                       # inference draw 0
    po = orc.prob_h(v)
    np.testing.assert_allclose(p, po, rtol=tol, atol=1e-7)
    u = O.philox_uniform(seed, O.draw_id("infer", 0), 0, n, H)
    near = np.abs(u.astype(np.float64) - po) <= tie
    assert np.array_equal(h[~near], (u < po).astype(np.float32)[~near]) and near.mean() < 1e-4
    vv, pv = m.inv_transform(h, want_p=True)                 # inference draw 1
    pvo = orc.prob_v(h)
    np.testing.assert_allclose(pv, pvo, rtol=tol, atol=1e-7)
    u = O.philox_uniform(seed, O.draw_id("infer", 1), 0, n, V)
    near = np.abs(u.astype(np.float64) - pvo) <= tie
    assert np.array_equal(vv[~near], (u < pvo).astype(np.float32)[~near]) and near.mean() < 1e-4
    np.testing.assert_allclose(m.free_energy(v), orc.free_energy(v), rtol=tol if compute == "bf16" else 1e-5, atol=1e-3)
