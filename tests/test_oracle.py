"""Self-checks of the oracle against closed-form facts ('Oracle self-tests')."""
import numpy as np
import pytest

from oracle import cd_oracle as O


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(*[np.uint64(c) for c in ctr], key[0], key[1])
        assert tuple(int(x) for x in got) == want


def test_philox_uniform_lattice_and_independence_of_sharding():
    u = O.philox_uniform(42, O.draw_id("train", 3, 0), 0, 64, 50)
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1
    assert np.array_equal(u * 2 ** 23, np.round(u * 2 ** 23))          # on the 2^-23 lattice
    tail = O.philox_uniform(42, O.draw_id("train", 3, 0), 40, 24, 50)   # rows 40.. drawn by another rank
    assert np.array_equal(u[40:], tail)
    other = O.philox_uniform(42, O.draw_id("train", 3, 2), 0, 64, 50)
    assert (other != u).mean() > 0.99
    big = O.philox_uniform(1, 0, 0, 512, 512)
    assert abs(big.mean() - 0.5) < 4 / np.sqrt(12 * big.size) + 1e-3


def test_bf16_round_and_split3():
    x = np.array([1.0, 1.00390625, 1.0078125, 1.01171875, -0.3333333, 3.1415927, 1e-20, 65504.0], np.float32)
    r = O.bf16_round(x)
    assert r[0] == 1.0 and r[1] == 1.0 and r[2] == 1.0078125 and r[3] == 1.015625   # ties to even
    assert np.all((r.view(np.uint32) & 0xFFFF) == 0)
    rng = np.random.default_rng(0)
    y = rng.normal(0, 1, 10000).astype(np.float32)
    hi, mid, lo = O.split3(y)
    assert np.array_equal((hi.astype(np.float64) + mid + lo).astype(np.float32), y)


def test_free_energy_is_the_marginal():
    """F(v) = -log sum_h exp(-E(v,h)), E = -v.b - h.c - v.W.h: brute force over 2^H states."""
    rng = np.random.default_rng(1)
    V, H = 7, 9
    W, b, c = rng.normal(0, 1, (V, H)), rng.normal(0, 1, V), rng.normal(0, 1, H)
    orc = O.OracleRBM(W, b, c)
    v = (rng.random((32, V)) < 0.5).astype(np.float32)
    hs = ((np.arange(1 << H)[:, None] >> np.arange(H)) & 1).astype(np.float64)
    Wd, bd, cd = orc.W.astype(np.float64), orc.b.astype(np.float64), orc.c.astype(np.float64)
    e = -(v @ bd)[:, None] - hs @ cd - (v @ Wd) @ hs.T
    np.testing.assert_allclose(orc.free_energy(v), -np.log(np.exp(-e).sum(1)), rtol=1e-5)


def test_conditionals_factorise():
    """P(h_j = 1 | v) from the joint equals sigmoid((vW + c)_j), and likewise for the visibles."""
    rng = np.random.default_rng(2)
    V, H = 5, 6
    orc = O.OracleRBM(rng.normal(0, 1, (V, H)), rng.normal(0, 1, V), rng.normal(0, 1, H))
    Wd, bd, cd = orc.W.astype(np.float64), orc.b.astype(np.float64), orc.c.astype(np.float64)
    hs = ((np.arange(1 << H)[:, None] >> np.arange(H)) & 1).astype(np.float64)
    v = (rng.random((1, V)) < 0.5).astype(np.float32)
    w = np.exp((v @ bd)[:, None] + hs @ cd + (v @ Wd) @ hs.T)[0]
    p = (w[:, None] * hs).sum(0) / w.sum()
    np.testing.assert_allclose(orc.prob_h(v)[0], p, rtol=1e-5)
    vs = ((np.arange(1 << V)[:, None] >> np.arange(V)) & 1).astype(np.float64)
    h = (rng.random((1, H)) < 0.5).astype(np.float32)
    w = np.exp(vs @ bd + (h @ cd) + vs @ Wd @ h[0])
    np.testing.assert_allclose(orc.prob_v(h)[0], (w[:, None] * vs).sum(0) / w.sum(), rtol=1e-5)


def test_dyadic_weights_are_order_independent():
    """Weights on a 2^-10 grid and binary data: every partial sum is exact in fp32, so fp32 / fp64 / bf16
    arithmetic agree bit for bit (the strict tier of the GPU parity tests rests on this)."""
    rng = np.random.default_rng(3)
    V, H, rows = 300, 200, 64
    W = (rng.integers(-51, 52, (V, H)) / 1024.0).astype(np.float32)
    b = (rng.integers(-51, 52, V) / 1024.0).astype(np.float32)
    c = (rng.integers(-51, 52, H) / 1024.0).astype(np.float32)
    v = (rng.random((rows, V)) < 0.4).astype(np.float32)
    pre = [O.OracleRBM(W, b, c, compute=m).pre_h(v) for m in ("f32", "f64", "bf16")]
    assert np.array_equal(pre[0], pre[1]) and np.array_equal(pre[1], pre[2])


def test_batches_remainder_last():
    assert list(O.batches(300, 128)) == [(0, 128), (128, 256), (256, 300)]      # rbm.py:110-111,211,218
    assert list(O.batches(256, 128)) == [(0, 128), (128, 256)]
    assert list(O.batches(0, 128)) == []


def test_cd1_moves_towards_the_data():
    """CD-1 on structured data lowers the free energy of the data relative to its reconstruction."""
    rng = np.random.default_rng(4)
    V, H, B = 32, 16, 64
    protos = (rng.random((4, V)) < 0.5).astype(np.float32)
    X = protos[rng.integers(0, 4, 512)]
    W, b, c = O.OracleRBM.init_params(V, H, seed=0)
    orc = O.OracleRBM(W, b, c)
    def recon(o):
        h, _ = o.sample_h(X[:B], O.lattice_uniform(rng, (B, H)))
        return np.mean((X[:B] - o.prob_v(h)) ** 2)
    before = recon(orc)
    for _ in range(30):
        for lo, hi in O.batches(512, B):
            orc.fused_step(X[lo:hi], [O.lattice_uniform(rng, (hi - lo, H))], [None, O.lattice_uniform(rng, (hi - lo, V))],
                           lr=0.1, scale=1.0 / B)
    assert recon(orc) < 0.3 * before


def test_sharded_statistics_sum_to_the_whole():
    """Data parallelism is exact: chains are row-independent, rows meet only in the batch sums."""
    rng = np.random.default_rng(5)
    V, H, B, n = 96, 80, 64, 4
    W, b, c = O.OracleRBM.init_params(V, H, seed=1)
    v = (rng.random((B, V)) < 0.3).astype(np.float32)
    seed, step = 11, 3
    def stats(rows, row0):
        orc = O.OracleRBM(W, b, c)
        u_h = [O.philox_uniform(seed, O.draw_id("train", step, 0), row0, len(rows), H)]
        u_v = [None, O.philox_uniform(seed, O.draw_id("train", step, 2), row0, len(rows), V)]
        return orc.cd_stats(rows, u_h, u_v)
    whole = stats(v, 0)
    parts = [stats(v[r * B // n:(r + 1) * B // n], r * B // n) for r in range(n)]
    assert np.array_equal(np.concatenate([p["h_pos"] for p in parts]), whole["h_pos"])
    assert np.array_equal(np.concatenate([p["v_neg"] for p in parts]), whole["v_neg"])
    np.testing.assert_allclose(sum(p["dW"].astype(np.float64) for p in parts), whole["dW"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(sum(p["db"] for p in parts), whole["db"], atol=1e-6)


def test_dbn_oracle_stack():
    rng = np.random.default_rng(6)
    dbn = O.OracleDBN()
    dbn.add_stack(O.OracleRBM(*O.OracleRBM.init_params(20, 12)))
    with pytest.raises(ValueError):
        dbn.add_stack(O.OracleRBM(*O.OracleRBM.init_params(13, 5)))
    dbn.add_stack(O.OracleRBM(*O.OracleRBM.init_params(12, 5)))
    x = (rng.random((8, 20)) < 0.5).astype(np.float32)
    h = dbn.transform(x, [O.lattice_uniform(rng, (8, 12)), O.lattice_uniform(rng, (8, 5))])
    assert h.shape == (8, 5)
    v = dbn.inv_transform(h, [O.lattice_uniform(rng, (8, 12)), O.lattice_uniform(rng, (8, 20))])
    assert v.shape == (8, 20)


def test_philox_stream_is_frozen():
    """The engine's draw keying (seed, draw id, global row, column) -> value; frozen so that neither side drifts."""
    u = O.philox_uniform(42, O.draw_id("train", 0, 0), 0, 2, 6)
    want = [[0.6129598617553711, 0.46858644485473633, 0.07323169708251953, 0.340861439704895, 0.9877185821533203,
             0.32706332206726074],
            [0.26124143600463867, 0.49120116233825684, 0.18712186813354492, 0.41673290729522705, 0.1328660249710083,
             0.40821826457977295]]
    assert u.tolist() == want
    n = O.philox_normal(42, O.draw_id("infer", 0), 0, 1, 4)
    np.testing.assert_allclose(n, [[0.4989447295665741, -0.22053277492523193, -0.8862244486808777, 0.9142584204673767]],
                               rtol=1e-6)
    assert O.draw_id("train", 3, 5) == 197 and O.draw_id("infer", 2) == (1 << 63) + 2 and O.draw_id("score", 1, 1) == (1 << 62) + 3


def test_wire_shards_model_of_bf16_partial_sums():
    """cd_stats(wire_shards=n): every shard's part of dW is rounded to bf16 and the parts are added in float32.  With
    8 rows per shard the binary positive term is an integer <= 8 (exact in bf16), so the only rounding is that of the
    negative term's partial sums: the result stays within one bf16 ulp per shard of the unsharded dW, equals it when
    the probabilities are dyadic, and equals the explicit per-shard computation."""
    rng = np.random.default_rng(11)
    V, H, B, n = 24, 16, 32, 4
    W, b, c = O.OracleRBM.init_params(V, H, seed=3)
    x = (rng.random((B, V)) < 0.4).astype(np.float32)
    u_h, u_v = O.lattice_uniform(rng, (B, H)), O.lattice_uniform(rng, (B, V))
    plain = O.OracleRBM(W, b, c, compute="bf16").cd_stats(x, [u_h], [None, u_v])
    wired = O.OracleRBM(W, b, c, compute="bf16").cd_stats(x, [u_h], [None, u_v], wire_shards=n)
    for key in ("h_pos", "v_neg", "h_neg", "db", "dc"):
        assert np.array_equal(plain[key], wired[key])
    # explicit per-shard computation
    rb = B // n
    hn = O.bf16_round(plain["h_neg"]).astype(np.float64)
    acc = np.zeros((V, H), np.float32)
    for s in range(n):
        sl = slice(s * rb, (s + 1) * rb)
        part = x[sl].astype(np.float64).T @ plain["h_pos"][sl].astype(np.float64) - \
            plain["v_neg"][sl].astype(np.float64).T @ hn[sl]
        acc = (acc + O.bf16_round(part.astype(np.float32))).astype(np.float32)
    assert np.array_equal(acc, wired["dW"])
    # within n bf16 ulps (2^-8 relative each) of partial sums bounded by rb
    assert np.abs(wired["dW"] - plain["dW"]).max() <= n * rb * 2.0 ** -8
    assert np.abs(wired["dW"] - plain["dW"]).max() > 0  # the rounding is visible at these sizes
    # zero weights and biases: every probability is exactly 1/2, the parts are multiples of 1/2 below 2^8 -> exact
    z = O.OracleRBM(np.zeros_like(W), np.zeros_like(b), np.zeros_like(c), compute="bf16")
    a = z.cd_stats(x, [u_h], [None, u_v])
    z2 = O.OracleRBM(np.zeros_like(W), np.zeros_like(b), np.zeros_like(c), compute="bf16")
    w = z2.cd_stats(x, [u_h], [None, u_v], wire_shards=n)
    assert np.array_equal(a["dW"], w["dW"])
    # bf16 all-reduce model: the running sum is rounded too, so every entry of dW is a bf16 number
    red = O.OracleRBM(W, b, c, compute="bf16").cd_stats(x, [u_h], [None, u_v], wire_shards=n, wire_sum_bf16=True)
    assert np.array_equal(red["dW"], O.bf16_round(red["dW"])) and not np.array_equal(red["dW"], wired["dW"])
    assert np.abs(red["dW"] - plain["dW"]).max() <= 2 * n * rb * 2.0 ** -8
    with pytest.raises(ValueError):
        O.OracleRBM(W, b, c).cd_stats(x[:30], [u_h[:30]], [None, u_v[:30]], wire_shards=4)


def test_grid_rounded_partial_sums_add_exactly_in_any_order():
    """What makes the engine's column statistics order-independent (csrc/rng_math.cuh: stat_grid_round, restated in the
    oracle): partial sums of probabilities (32 rows each) rounded to the grid of the minibatch size, plus integer counts, add
    up to the same fp32 value in every order - while the unrounded partials do not - and the rounding costs less than the
    fp32 spacing at the size of the sums."""
    rng = np.random.default_rng(0)
    for rows in (16, 104, 128, 300, 4096, 32768):
        groups = (rows + 31) // 32
        p = rng.random((groups, 32)).astype(np.float32)
        p.reshape(-1)[rows:] = 0                                  # rows beyond the minibatch contribute nothing
        partial = p.sum(axis=1, dtype=np.float32)                 # one warp's sum
        valid = np.minimum(32, rows - 32 * np.arange(groups))     # rows of each 32-row group that exist
        counts = rng.integers(0, valid + 1).astype(np.float32)    # the positive phase: 0/1 states, integer sums
        grid = O.stat_grid_round(-partial, rows)
        assert np.abs(grid + partial).max() <= 2.0 ** (int(np.ceil(np.log2(max(rows, 2)))) - 25) + 1e-12
        seen_grid, seen_raw = set(), set()
        for _ in range(50):
            order = rng.permutation(2 * groups)
            for terms, seen in ((np.concatenate([counts, grid]), seen_grid), (np.concatenate([counts, -partial]), seen_raw)):
                acc = np.float32(0)
                for i in order:
                    acc = np.float32(acc + terms[i])
                seen.add(float(acc))
        assert len(seen_grid) == 1, (rows, seen_grid)
        exact = float(np.sum(counts.astype(np.float64)) + np.sum(grid.astype(np.float64)))
        assert seen_grid == {exact}                               # ... and that value is the exact sum of the rounded terms
        if rows >= 300:
            assert len(seen_raw) > 1                              # the unrounded partials do depend on the order
