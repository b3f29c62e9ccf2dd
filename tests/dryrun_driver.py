"""Runs one scenario of the engine's HOST side against the fake CUDA runtime (tools/dryrun/fake_cudart.cpp) and prints a
JSON summary: the recorded launches, the nodes of the last captured graph, allocation counters and every complaint of the
fake (out-of-bounds copies, invalid tensor-map arguments, kernels pointed outside their buffers).  Executed in a process of
its own by tests/test_host_dryrun.py - no torch, no GPU; environment switches of the scenario are set by the caller."""
import ctypes as C
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DRY = os.path.join(ROOT, "build", "dryrun")
os.environ["KUCD_NCCL_LIB"] = os.path.join(DRY, "libfakecudart.so")

import numpy as np  # noqa: E402

from keras_unsupervised_b200 import _lib as L  # noqa: E402

L.LIB_PATH = os.path.join(DRY, "libkucd_dry.so")
from keras_unsupervised_b200.engine import Context, Dataset, Machine  # noqa: E402

fake = C.CDLL(os.path.join(DRY, "libfakecudart.so"), mode=C.RTLD_GLOBAL)
fake.fake_counter.restype = C.c_longlong


def _lines(size_fn, line_fn):
    out, buf = [], C.create_string_buffer(1024)
    for i in range(size_fn()):
        line_fn(i, buf, 1024)
        out.append(buf.value.decode())
    return out


def _kernel(line):
    """'launch _ZN4kucd15update_w_kernelILb0EE... grid=..' -> 'update_w_kernel<0>' style short name"""
    m = re.match(r"launch (\S+)", line)
    if not m:
        return line.split()[0]
    name = m.group(1)
    short = re.sub(r"^_ZN4kucd\d+", "", name)
    base = re.match(r"[a-z_0-9]+", short)
    base = base.group(0) if base else short
    targs = re.findall(r"(?:Li|Lb)(\d+)E", short.split("EEv")[0]) if "I" in short[len(base):len(base) + 1] else []
    return base + ("<" + ",".join(targs) + ">" if targs else "")


def snapshot():
    log = _lines(fake.fake_log_size, fake.fake_log_line)
    return {"log": log, "kernels": [_kernel(x) for x in log if x.startswith("launch") or x.startswith("graph_launch")
                                    or x.startswith("allreduce") or x.startswith("broadcast")],
            "graph": [_kernel(x) for x in _lines(fake.fake_last_graph_size, fake.fake_last_graph_line)
                      if not x.startswith("event_")],
            "graph_raw": _lines(fake.fake_last_graph_size, fake.fake_last_graph_line),
            "errors": _lines(fake.fake_error_count, fake.fake_error_line),
            "decoded": fake.fake_counter(5), "mallocs": fake.fake_counter(0), "frees": fake.fake_counter(1), "live_bytes": fake.fake_counter(2),
            "peak_bytes": fake.fake_counter(3), "async_allocs": fake.fake_counter(6)}


def machine(ctx, V, H, compute=L.COMPUTE_BF16, mode=L.MODE_VISIBLE_BERNOULLI):
    m = Machine(ctx, V, H, mode, compute, seed=3)
    rng = np.random.default_rng(0)
    m.set_params(rng.uniform(-0.05, 0.05, (V, H)).astype(np.float32), np.zeros(V, np.float32), np.zeros(H, np.float32))
    return m


def data(n, V, seed=1):
    return (np.random.default_rng(seed).random((n, V)) < 0.3).astype(np.float32)


def scenario(name):
    out = {}
    if name == "cd_step":
        ctx = Context(device=0, seed=1)
        for compute in (L.COMPUTE_BF16, L.COMPUTE_F32X3):
            m = machine(ctx, 333, 130, compute)
            fake.fake_reset()
            m.cd_step(data(200, 333), Machine.hparams(lr=1e-3, k=2))
            out["bf16" if compute == L.COMPUTE_BF16 else "f32"] = snapshot()
    elif name == "fit_epoch":
        ctx = Context(device=0, seed=1)
        m = machine(ctx, 784, 500)
        ds = Dataset.from_array(ctx, data(1000, 784), L.COMPUTE_BF16)
        fake.fake_reset()
        st = m.fit_epoch(ds, 128, Machine.hparams(lr=1e-3, k=1))
        out = snapshot()
        out["steps"] = st["steps"]
        out["timings"] = ctx.timings()
    elif name == "full_size":  # BASELINE.json's C3 and C4 shapes: descriptors, bounds and kernel choice at full size
        ctx = Context(device=0, seed=1)
        for key, V, H, B, k, pcd in (("c3", 4096, 4096, 4096, 10, False), ("c4", 16384, 8192, 1024, 1, True)):
            m = Machine(ctx, V, H, L.MODE_VISIBLE_BERNOULLI, L.COMPUTE_BF16, seed=3)
            X = np.zeros((2 * B, V), np.uint8)
            ds = Dataset.from_array(ctx, X, L.COMPUTE_BF16)
            if pcd:
                m.set_chains(np.zeros((B, V), np.uint8))
            fake.fake_reset()
            m.fit_epoch(ds, B, Machine.hparams(lr=1e-3, k=k, persistent=pcd, normalize=True), want_stats=False)
            out[key] = snapshot()
            ds.close()
            m.close()
    elif name == "python_surface":  # the reference-facing classes end to end (values are meaningless in a dry run)
        from keras_unsupervised_b200.ebm import DBN, RBM, MODE_VISIBLE_BERNOULLI

        ctx = Context(device=0, seed=1)
        X = data(600, 784)
        hps = {"batch_size": 128, "epochs": 2, "lr": 1e-3, "dtype": "bf16"}
        dbn = DBN()
        for i, h in enumerate((500, 500, 2000)):
            dbn.add_stack(RBM(dict(hps), h, name="rbm%d" % i, mode=MODE_VISIBLE_BERNOULLI, context=ctx))
        fake.fake_reset()
        dbn.fit(X, verbose=0)
        out["fit"] = snapshot()
        fake.fake_reset()
        H = dbn.transform(X[:64])
        V2 = dbn.inv_transform(H)
        out["shapes"] = [list(H.shape), list(V2.shape)]
        out["transform"] = snapshot()
        fake.fake_reset()
        dbn.fine_tune(X[:256], epochs=1, lr=1e-3, k=1)
        out["fine_tune"] = snapshot()
        fake.fake_reset()
        g = dbn.generate(16, gibbs_steps=3)
        out["generate_shape"] = list(g.shape)
        out["generate"] = snapshot()
        one = RBM({"batch_size": 128, "epochs": 1, "lr": 1e-3}, 96, name="one", mode=MODE_VISIBLE_BERNOULLI, context=ctx)
        fake.fake_reset()
        one.fit(data(1000, 200), verbose=0)                                  # float32-grade, one epoch: streamed
        fe = one.cal_free_energy(data(50, 200))[0]
        out["fe_shape"] = list(np.asarray(fe).shape)
        out["one_epoch_f32"] = snapshot()
    elif name == "data_formats":  # packed bits, uint8, shuffle, Gaussian visibles, PCD, injected draws, score
        from keras_unsupervised_b200.data import PackedBits

        ctx = Context(device=0, seed=1)
        V, H = 333, 130
        X = data(500, V)
        P = PackedBits(np.packbits(X.astype(np.uint8), axis=1, bitorder="little"), V)
        m = machine(ctx, V, H)
        fake.fake_reset()
        m.fit_host(P, 128, Machine.hparams(lr=1e-3, k=1))
        m.fit_host(X.astype(np.uint8), 128, Machine.hparams(lr=1e-3, k=1))
        ds = Dataset.from_array(ctx, P, L.COMPUTE_BF16)
        sh = ds.shuffled(seed=5, epoch=0)
        sh = ds.shuffled(seed=5, epoch=1, into=sh)
        back = sh.packed()
        dense = ds.numpy()
        out["shapes"] = [list(back.data.shape), list(dense.shape)]
        h = m.transform(P, out_dtype="bits")
        out["bits_out"] = [list(h.data.shape), h.n_cols]
        out["packed"] = snapshot()
        g = machine(ctx, V, H, mode=L.MODE_VISIBLE_GAUSSIAN)
        fake.fake_reset()
        g.cd_step(np.random.default_rng(2).standard_normal((200, V)).astype(np.float32), Machine.hparams(lr=1e-3, k=2))
        g32 = machine(ctx, V, H, compute=L.COMPUTE_F32X3, mode=L.MODE_VISIBLE_GAUSSIAN)
        g32.cd_step(np.random.default_rng(2).standard_normal((200, V)).astype(np.float32), Machine.hparams(lr=1e-3, k=1))
        out["gaussian"] = snapshot()
        fake.fake_reset()
        m.set_chains(data(128, V, seed=9))
        m.cd_step(X[:128], Machine.hparams(lr=1e-3, k=1, persistent=True))
        rng = np.random.default_rng(4)
        f = machine(ctx, V, H, compute=L.COMPUTE_F32X3)
        u_h = [rng.random((96, H)).astype(np.float32)]
        u_v = [None, rng.random((96, V)).astype(np.float32)]
        st = f.cd_step(X[:96], Machine.hparams(lr=1e-3, k=1, want_stats=True), u_h=u_h, u_v=u_v)
        f.last_stats(96)
        f.score(X[:96])
        out["pcd_inject_score"] = snapshot()
    elif name == "sweep":  # many shapes / options through whatever switches the environment sets; only complaints matter
        from keras_unsupervised_b200.data import PackedBits

        ctx = Context(device=0, seed=1)
        fake.fake_reset()
        runs = 0
        rng = np.random.default_rng(0)
        for V, H in ((784, 500), (333, 130), (72, 1000), (1024, 64)):
            for compute in (L.COMPUTE_BF16, L.COMPUTE_F32X3):
                m = machine(ctx, V, H, compute)
                for N, B in ((1000, 128), (512, 128), (100, 128), (130, 64), (2048, 512)):
                    X = data(N, V)
                    for kw in (dict(k=1), dict(k=3, momentum=0.5, normalize=True), dict(k=1, persistent=True)):
                        if kw.get("persistent"):
                            m.set_chains(data(min(B, N), V, seed=3))
                        hp = Machine.hparams(lr=1e-3, **kw)
                        m.fit_host(X, B, hp)
                        m.fit_host(X[:, ::1].astype(np.uint8), B, hp)
                        wide = np.zeros((N, V + 24), np.float32)      # rows with a pitch: the 2-D copy path
                        wide[:, :V] = X
                        m.fit_host(wide[:, :V], B, hp)
                        m.fit_host(PackedBits(np.packbits(X.astype(np.uint8), axis=1, bitorder="little"), V), B, hp)
                        ds = Dataset.from_array(ctx, X, compute)
                        m.fit_epoch(ds, B, hp, want_stats=bool(runs % 2))
                        m.fit_range(ds, B, hp, 1, -1) if N > B else None
                        ds.close()
                        m.cd_step(X[:B], hp)
                        runs += 7
                m.close()
        out = snapshot()
        out["runs"] = runs
        out.pop("log")
        out["kernels"] = len(out["kernels"])
    elif name == "ranks_sweep":  # DRY_RANKS contexts of this process as the ranks of one group, several shapes
        n = int(os.environ.get("DRY_RANKS", "2"))
        os.environ["FAKE_CUDA_DEVICES"] = str(n)
        ctxs = [Context(device=r, seed=1) for r in range(n)]
        uid = (C.c_char * 128)()
        L.check(ctxs[0].lib.kucd_comm_unique_id(uid))
        for r, c in enumerate(ctxs):
            L.check(c.lib.kucd_ctx_comm_init(c.handle, bytes(uid), r, n))
            c.rank, c.world = r, n
        fused = os.environ.get("KUCD_FUSED_REDUCE", "1") != "0"
        fake.fake_reset()
        runs = 0
        for V, H in ((784, 500), (333, 130), (130, 72), (1024, 256)):
            ms = []
            for c in ctxs:
                m = Machine.__new__(Machine)
                m.ctx, m.V, m.H, m.mode, m.compute = c, V, H, L.MODE_VISIBLE_BERNOULLI, L.COMPUTE_BF16
                h = C.c_void_p()
                L.check(c.lib.kucd_rbm_create(c.handle, V, H, m.mode, m.compute, C.byref(h)))
                m.handle, m.fused_reduce = h, False
                c._children.add(m)
                m.set_params(np.zeros((V, H), np.float32), np.zeros(V, np.float32), np.zeros(H, np.float32))
                ms.append(m)
            if fused:
                handles = []
                for m in ms:
                    buf = (C.c_char * 128)()
                    L.check(m.ctx.lib.kucd_rbm_peer_export(m.handle, buf))
                    handles.append(bytes(buf))
                for m in ms:
                    L.check(m.ctx.lib.kucd_rbm_peer_attach(m.handle, b"".join(handles)))
            for b, steps in ((64, 3), (256, 2)):
                for kw in (dict(k=1), dict(k=2, momentum=0.5, normalize=True)):
                    hp = Machine.hparams(lr=1e-3, **kw)
                    for r, m in enumerate(ms):
                        X = data(b * steps + b // 2, V, seed=r)                  # full minibatches + a remainder
                        ds = Dataset.from_array(m.ctx, X, L.COMPUTE_BF16)
                        m.fit_epoch(ds, b, hp, global_row0=b * r, want_stats=False)
                        m.fit_host(X, b, hp, global_row0=b * r)
                        m.cd_step(X[:b], hp, global_row0=b * r)
                        m.get_params()
                        ds.close()
                        runs += 3
            for m in ms:
                m.close()
        out = snapshot()
        out["runs"] = runs
        out.pop("log")
        out["kernels"] = len(out["kernels"])
        out["timings"] = ctxs[0].timings()
    elif name == "errors":  # argument errors of the C ABI: the message, the exception type, nothing leaked, nothing launched
        ctx = Context(device=0, seed=1)
        m = machine(ctx, 64, 32)
        ds = Dataset.from_array(ctx, data(256, 64), L.COMPUTE_BF16)
        m.fit_epoch(ds, 64, Machine.hparams(lr=1e-3))
        m.cd_step(data(64, 64), Machine.hparams(lr=1e-3))
        fake.fake_reset()
        live0 = fake.fake_counter(4)
        seen = []

        def expect(kind, fn):
            try:
                fn()
                seen.append(("no error", ""))
            except kind as e:  # noqa: PERF203
                seen.append((kind.__name__, str(e)[:90]))
            except Exception as e:  # noqa: BLE001
                seen.append(("other " + type(e).__name__, str(e)[:90]))

        expect(ValueError, lambda: m.cd_step(data(64, 65), Machine.hparams(lr=1e-3)))                 # wrong column count
        expect(ValueError, lambda: m.transform(data(8, 63)))
        expect(ValueError, lambda: m.inv_transform(data(8, 33)))
        expect(ValueError, lambda: m.free_energy(data(8, 1)))
        expect(ValueError, lambda: m.cd_step(data(64, 64), Machine.hparams(lr=1e-3, k=0)))            # k out of range
        expect(ValueError, lambda: m.cd_step(data(64, 64), Machine.hparams(lr=1e-3, k=40)))
        expect(ValueError, lambda: m.cd_step(data(64, 64), Machine.hparams(lr=1e-3, persistent=True)))  # no chains set
        expect(ValueError, lambda: m.fit_epoch(ds, 0, Machine.hparams(lr=1e-3)))                      # batch size
        expect(ValueError, lambda: m.fit_range(ds, 64, Machine.hparams(lr=1e-3), 3, 2))               # empty / reversed range
        expect(ValueError, lambda: m.fit_range(ds, 64, Machine.hparams(lr=1e-3), 0, 99))
        expect(ValueError, lambda: m.set_params(W=np.zeros((64, 31), np.float32)))
        expect(ValueError, lambda: m.cd_step(data(64, 64).astype(np.float64).view(np.int64), Machine.hparams(lr=1e-3))
               if False else m.cd_step(np.zeros((64, 64), np.int16).view(np.uint16), Machine.hparams(lr=1e-3)))
        expect(ValueError, lambda: m.delta_rule(True, data(16, 64), data(16, 31), 0.1))               # target shape
        expect(ValueError, lambda: m.delta_rule(False, data(16, 64), data(16, 64), 0.1))              # input is (rows, H) backward
        other = machine(ctx, 48, 32)
        expect(ValueError, lambda: other.fit_epoch(ds, 64, Machine.hparams(lr=1e-3)))                 # data set of another width
        expect(ValueError, lambda: ds.shuffled(1, 0, into=ds))                                        # in place
        expect(ValueError, lambda: Machine(ctx, 0, 32, L.MODE_VISIBLE_BERNOULLI, L.COMPUTE_BF16))
        expect(ValueError, lambda: Machine(ctx, 16, 32, 2, L.COMPUTE_BF16))                           # MODE_COMPLEX
        expect(ValueError, lambda: Context(device=7))                                                 # 2 fake devices
        m.cd_step(data(0, 64), Machine.hparams(lr=1e-3))                                              # empty minibatch: a no-op
        out = snapshot()
        out["seen"] = seen
        other.close()
        out["leaked"] = fake.fake_counter(4) - live0
    elif name == "oom":  # the n-th cudaMalloc of a whole session fails, for n = 1 .. 45: an error, no crash, no leak
        fake.fake_fail_malloc_in.argtypes = [C.c_long]
        results = []
        X = data(300, 96)
        for n in range(1, 46):
            fake.fake_reset()
            live0 = fake.fake_counter(4)
            fake.fake_fail_malloc_in(n)
            objs, outcome = [], "ok"
            try:
                ctx = Context(device=0, seed=1)
                objs.append(ctx)
                m = machine(ctx, 96, 40, L.COMPUTE_F32X3 if n % 2 else L.COMPUTE_BF16)
                ds = Dataset.from_array(ctx, X, m.compute)
                m.fit_epoch(ds, 128, Machine.hparams(lr=1e-3, momentum=0.5))
                h = m.transform_dataset(ds)
                sh = ds.shuffled(1, 0)
                m.fit_host(X, 128, Machine.hparams(lr=1e-3))
                m.delta_rule(True, X[:64], data(64, 40, seed=2), 0.1)
                m.transform(X[:32])
                m.free_energy(X[:32])
            except L.KucdError as e:
                outcome = "KucdError"
                if "cudaMalloc" not in str(e) and "memory" not in str(e).lower():
                    outcome = "KucdError (other): " + str(e)[:80]
            except Exception as e:  # noqa: BLE001
                outcome = type(e).__name__ + ": " + str(e)[:80]
            fake.fake_fail_malloc_in(0)
            for o in objs:
                o.close()                      # closes the machines / data sets it still tracks, then the context
            results.append((n, outcome, fake.fake_counter(4) - live0, fake.fake_error_count()))
        out = {"results": results}
    elif name == "rbm_options":  # RBM.fit under the optional hps keys (values are meaningless in a dry run)
        from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI, MODE_VISIBLE_GAUSSIAN

        ctx = Context(device=0, seed=1)
        X = data(600, 200)
        for key, hps, mode in (
                ("shuffle", {"epochs": 3, "shuffle": True, "dtype": "bf16"}, MODE_VISIBLE_BERNOULLI),
                ("reference", {"epochs": 1, "compat": "reference"}, MODE_VISIBLE_BERNOULLI),
                ("pcd", {"epochs": 2, "persistent": True, "k": 2, "dtype": "bf16", "momentum": 0.5, "weight_decay": 1e-4,
                         "normalize": "mean"}, MODE_VISIBLE_BERNOULLI),
                ("gaussian_default", {"epochs": 1}, MODE_VISIBLE_GAUSSIAN),
                ("resident_one_epoch", {"epochs": 1, "stream": False, "dtype": "bf16"}, MODE_VISIBLE_BERNOULLI)):
            base = {"batch_size": 128, "lr": 1e-3}
            base.update(hps)
            r = RBM(base, 64, name=key, mode=mode, context=ctx)
            fake.fake_reset()
            r.fit(X, verbose=0)
            out[key] = snapshot()
            out[key]["history"] = len(r.history)
    elif name == "split":  # KUCD_SPLIT=2 KUCD_CHAIN=0: two Gibbs chains on two streams, forked and joined inside the capture
        ctx = Context(device=0, seed=1)
        m = machine(ctx, 784, 500)
        ds = Dataset.from_array(ctx, data(1024, 784), L.COMPUTE_BF16)
        fake.fake_reset()
        m.fit_epoch(ds, 512, Machine.hparams(lr=1e-3, k=2, persistent=False))
        out = snapshot()
    elif name == "fit_host":  # KUCD_STREAM_CHUNK / KUCD_STREAM_GRAPH are read from the environment
        ctx = Context(device=0, seed=1)
        m = machine(ctx, 300, 200)
        X = data(1000, 300)
        fake.fake_reset()
        st = m.fit_host(X, 128, Machine.hparams(lr=1e-3, k=2))
        out = snapshot()
        out["steps"] = st["steps"]
        out["timings"] = ctx.timings()
        fake.fake_reset()
        st = m.fit_host(X, 128, Machine.hparams(lr=1e-3, k=2))          # a second pass reuses the captured step
        out["second"] = snapshot()
    elif name == "transform_loop":  # data-set planes from the stream-ordered pool
        ctx = Context(device=0, seed=1)
        m = machine(ctx, 512, 256)
        ds = Dataset.from_array(ctx, data(1024, 512), L.COMPUTE_BF16)
        m.transform_dataset(ds).close()
        fake.fake_reset()
        for _ in range(10):
            h = m.transform_dataset(ds)
            v = m.inv_transform_dataset(h)
            h.close()
            v.close()
        out = snapshot()
        ds.close()
        ctx.close()
        out["live_after_close"] = fake.fake_counter(4)
    elif name == "slabs":  # KUCD_AR_SLABS (+ KUCD_AR_SLABS_MIN_ELEMS=1), one rank
        ctx = Context(device=0, seed=1)
        m = machine(ctx, 784, 500)
        ds = Dataset.from_array(ctx, data(512, 784), L.COMPUTE_BF16)
        fake.fake_reset()
        m.fit_epoch(ds, 128, Machine.hparams(lr=1e-3, k=1, momentum=0.5))
        out = snapshot()
        fake.fake_reset()
        m.cd_step(data(1024, 784), Machine.hparams(lr=1e-3, k=1))            # large-tile chain + slabs, direct launches
        out["direct"] = snapshot()
        fake.fake_reset()
        m.cd_step(data(100, 784), Machine.hparams(lr=1e-3, k=1, update_mask=L.UPDATE_C | L.UPDATE_B))
        out["no_w"] = snapshot()
    elif name == "momentum":  # checkpoint state in and out of the ABI: momentum buffers (padded rows), draw counters
        ctx = Context(device=0, seed=1)
        m = machine(ctx, 333, 130)                              # ldH = 192: the copies are pitched
        out["absent"] = m.get_momentum() is None
        rng = np.random.default_rng(4)
        mW, mb, mc = (rng.normal(0, 1, sh).astype(np.float32) for sh in ((333, 130), (333,), (130,)))
        m.set_momentum(mW, mb, mc)
        got = m.get_momentum()
        out["round_trip"] = bool(got is not None and all(np.array_equal(a, b) for a, b in zip(got, (mW, mb, mc))))
        m.cd_step(data(200, 333), Machine.hparams(lr=1e-3, k=1, momentum=0.5))   # the buffers it finds are used, not replaced
        out["used"] = m.get_momentum() is not None
        m2 = machine(ctx, 333, 130)
        m2.cd_step(data(200, 333), Machine.hparams(lr=1e-3, k=1, momentum=0.5))
        out["created_by_step"] = m2.get_momentum() is not None
        m.set_draw_counters(7, 9)
        out["draws"] = m.draw_counters()
        m.set_seed(3, 5)
        out["draws_after_set_seed"] = m.draw_counters()
        try:
            m.set_momentum(mW[:, :100], mb, mc)
            out["bad_shape"] = "accepted"
        except ValueError as e:
            out["bad_shape"] = str(e)
        out["snapshot"] = snapshot()
    elif name == "delta_rule":
        ctx = Context(device=0, seed=1)
        for compute in (L.COMPUTE_BF16, L.COMPUTE_F32X3):
            m = machine(ctx, 333, 130, compute)
            key = "bf16" if compute == L.COMPUTE_BF16 else "f32"
            fake.fake_reset()
            m.delta_rule(True, data(200, 333), data(200, 130, seed=2), 1e-2)
            out[key + "_fwd"] = snapshot()
            fake.fake_reset()
            m.delta_rule(False, data(200, 130, seed=2), data(200, 333), 1e-2, normalize=True)
            out[key + "_bwd"] = snapshot()
    elif name == "two_ranks":  # KUCD_WIRE_BF16 / KUCD_AR_SLABS / KUCD_FUSED_REDUCE from the environment
        # two contexts of this one process act as rank 0 and rank 1: the fake NCCL hands out communicators without
        # talking to anybody, the fake IPC passes pointers through, so the exchange buffers really are each other's
        ctxs = [Context(device=r, seed=1) for r in range(2)]
        uid = (C.c_char * 128)()
        L.check(ctxs[0].lib.kucd_comm_unique_id(uid))
        for r, c in enumerate(ctxs):
            L.check(c.lib.kucd_ctx_comm_init(c.handle, bytes(uid), r, 2))
            c.rank, c.world = r, 2
        V, H = int(os.environ.get("DRY_V", "784")), int(os.environ.get("DRY_H", "500"))
        ms = []
        for c in ctxs:
            m = Machine.__new__(Machine)
            m.ctx, m.V, m.H, m.mode, m.compute = c, V, H, L.MODE_VISIBLE_BERNOULLI, L.COMPUTE_BF16
            h = C.c_void_p()
            L.check(c.lib.kucd_rbm_create(c.handle, V, H, m.mode, m.compute, C.byref(h)))
            m.handle, m.fused_reduce = h, False
            c._children.add(m)
            m.set_params(np.zeros((V, H), np.float32), np.zeros(V, np.float32), np.zeros(H, np.float32))
            ms.append(m)
        fused = os.environ.get("KUCD_FUSED_REDUCE", "1") != "0"
        if fused:
            handles = []
            for m in ms:
                buf = (C.c_char * 128)()
                L.check(m.ctx.lib.kucd_rbm_peer_export(m.handle, buf))
                handles.append(bytes(buf))
            # the fake's canary check reads through the "mapped" pointer: fine, it is the same memory
            for m in ms:
                L.check(m.ctx.lib.kucd_rbm_peer_attach(m.handle, b"".join(handles)))
                m.fused_reduce = True
        b = int(os.environ.get("DRY_ROWS", "64"))  # rows of a global minibatch per rank
        dss = [Dataset.from_array(c, data(4 * b, V, seed=5 + r), L.COMPUTE_BF16) for r, c in enumerate(ctxs)]
        fake.fake_reset()
        hp = Machine.hparams(lr=1e-3, k=1)
        for r, m in enumerate(ms):
            m.fit_epoch(dss[r], b, hp, global_row0=b * r, want_stats=False)
        out = snapshot()
        out["timings"] = [c.timings() for c in ctxs]
    elif name == "units":  # KUCD_EXCHANGE=units: the unit-sharded step over DRY_RANKS in-process ranks
        n = int(os.environ.get("DRY_RANKS", "2"))
        os.environ["FAKE_CUDA_DEVICES"] = str(n)
        ctxs = [Context(device=r, seed=1) for r in range(n)]
        uid = (C.c_char * 128)()
        L.check(ctxs[0].lib.kucd_comm_unique_id(uid))
        for r, c in enumerate(ctxs):
            L.check(c.lib.kucd_ctx_comm_init(c.handle, bytes(uid), r, n))
            c.rank, c.world = r, n
        V, H, b = 128 * n * 2, 128 * n, 128
        ms = []
        for c in ctxs:
            m = Machine.__new__(Machine)
            m.ctx, m.V, m.H, m.mode, m.compute = c, V, H, L.MODE_VISIBLE_BERNOULLI, L.COMPUTE_BF16
            h = C.c_void_p()
            L.check(c.lib.kucd_rbm_create(c.handle, V, H, m.mode, m.compute, C.byref(h)))
            m.handle, m.fused_reduce = h, False
            c._children.add(m)
            m.set_params(np.zeros((V, H), np.float32), np.zeros(V, np.float32), np.zeros(H, np.float32))
            ms.append(m)
        handles = []
        for m in ms:
            buf = (C.c_char * 128)()
            L.check(m.ctx.lib.kucd_rbm_peer_export(m.handle, buf))
            handles.append(bytes(buf))
        for m in ms:
            L.check(m.ctx.lib.kucd_rbm_peer_attach(m.handle, b"".join(handles)))
        dss = [Dataset.from_array(c, data(3 * b, V, seed=5 + r), L.COMPUTE_BF16) for r, c in enumerate(ctxs)]
        fake.fake_reset()
        for r, m in enumerate(ms):
            m.fit_epoch(dss[r], b, Machine.hparams(lr=1e-3, k=2), global_row0=b * r, want_stats=False)
        out = snapshot()
        fake.fake_reset()
        for r, m in enumerate(ms):   # persistent chains, momentum, statistics of the last minibatch, then single steps
            m.set_chains(data(b, V, seed=9 + r))
            m.fit_epoch(dss[r], b, Machine.hparams(lr=1e-3, k=1, persistent=True, momentum=0.5, normalize=True),
                        global_row0=b * r, want_stats=True)
            m.cd_step(data(b, V, seed=20 + r), Machine.hparams(lr=1e-3, k=1, persistent=True), global_row0=b * r)
            m.get_chains(b)
            m.get_params()
        out["pcd"] = snapshot()
        fake.fake_reset()
        for r, m in enumerate(ms):   # a range with a remainder minibatch falls back to the data-parallel exchange
            ds = Dataset.from_array(m.ctx, data(2 * b + 64, V, seed=30 + r), L.COMPUTE_BF16)
            m.fit_epoch(ds, b, Machine.hparams(lr=1e-3, k=1), global_row0=b * r, want_stats=False)
            ds.close()
        out["remainder"] = snapshot()
        out["timings"] = [c.timings() for c in ctxs]
    else:
        raise SystemExit("unknown scenario " + name)
    return out


if __name__ == "__main__":
    print(json.dumps(scenario(sys.argv[1])))
