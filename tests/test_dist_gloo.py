"""The N > 1 path on CPU: two gloo ranks shard every global minibatch by rows (parallel.shard_rows),
draw with global-row Philox counters, all-reduce dW / db / dc and apply the same update - and end with
exactly the parameters a single rank computes.  The arithmetic is the oracle's; what is under test is
the host-side sharding / offset / reduction logic the GPU path uses unchanged."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from keras_unsupervised_b200.parallel import shard_rows
from oracle import cd_oracle as O

V, H, B, N, SEED, LR = 48, 40, 16, 72, 5, 0.01     # 4 full minibatches + a remainder of 8 rows


def _data():
    return (np.random.default_rng(1).random((N, V)) < 0.3).astype(np.float32)


def _single():
    orc = O.OracleRBM(*O.OracleRBM.init_params(V, H, seed=2), compute="f64")
    O.philox_fit(orc, _data(), B, 1, LR, SEED)
    return orc


def _worker(rank, world, port, out, wire16=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local, b, row0 = shard_rows(_data(), B, rank, world)
    orc = O.OracleRBM(*O.OracleRBM.init_params(V, H, seed=2), compute="bf16" if wire16 else "f64")
    step = 0
    for lo, hi in O.batches(local.shape[0], b):
        rows = hi - lo
        # on the remainder minibatch a rank owns rows [rank*rows, ...) of the (smaller) global minibatch
        g0 = row0 if rows == b else rank * rows
        u_h = [O.philox_uniform(SEED, O.draw_id("train", step, 0), g0, rows, H)]
        u_v = [None, O.philox_uniform(SEED, O.draw_id("train", step, 2), g0, rows, V)]
        st = orc.cd_stats(local[lo:hi], u_h, u_v)
        if wire16:  # KUCD_WIRE_BF16: this rank's part of dW is rounded to bf16 before it is exchanged, summed in fp32
            st["dW"] = O.bf16_round(st["dW"])
        for key in ("dW", "db", "dc"):
            t = torch.from_numpy(st[key].astype(np.float64))
            dist.all_reduce(t)
            st[key] = t.numpy().astype(np.float32)
        orc.apply(st, LR)
        step += 1
    if rank == 0:
        out.put((orc.W, orc.b, orc.c, step))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_equal_one():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    W, b, c, steps = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _single()
    assert steps == 5
    np.testing.assert_allclose(W, ref.W, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(b, ref.b, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(c, ref.c, rtol=1e-6, atol=1e-7)


@pytest.mark.timeout(120)
def test_two_ranks_with_bf16_partial_sums_equal_the_wire_model():
    """The opt-in exchange with bf16 partial sums (KUCD_WIRE_BF16=1, fused path): every rank rounds its part of dW to
    bf16, the parts are added in float32.  Two gloo ranks doing exactly that end with the parameters of the
    single-process oracle run with wire_shards=2 - the model tests/dp_check.py holds the GPU path to."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    W, b, c, steps = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = O.OracleRBM(*O.OracleRBM.init_params(V, H, seed=2), compute="bf16")
    O.philox_fit(ref, _data(), B, 1, LR, SEED, wire_shards=2)
    plain = O.OracleRBM(*O.OracleRBM.init_params(V, H, seed=2), compute="bf16")
    O.philox_fit(plain, _data(), B, 1, LR, SEED)
    assert steps == 5
    np.testing.assert_allclose(W, ref.W, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(b, ref.b, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(c, ref.c, rtol=1e-6, atol=1e-7)
    assert np.abs(ref.W - plain.W).max() > 0          # the rounding is visible ...
    assert np.abs(ref.W - plain.W).max() < 5 * LR      # ... and small: a few bf16 ulps of a partial sum, times lr, 5 steps
