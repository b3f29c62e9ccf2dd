"""Thin object layer over the C ABI (include/kucd.h): Context, Dataset and Machine (an RBM's device
state).  Nothing here computes: every method marshals arrays into kucd_tensor descriptors and calls
libkucd.so.  The reference-facing classes live in keras_unsupervised_b200/ebm.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import _lib as L
from .data import PackedBits


def _out_like(x, rows: int, cols: int, dtype=None):
    """Allocate the output next to the input: numpy in -> numpy out, torch in -> torch out (same device);
    dtype="bits": a host PackedBits (0/1 states, one bit per unit)."""
    if isinstance(dtype, str):
        if dtype != "bits":
            raise ValueError("out_dtype must be a numpy / torch dtype or 'bits'")
        return PackedBits(np.empty((rows, (cols + 7) // 8), dtype=np.uint8), cols)
    if isinstance(x, PackedBits):
        x = x.data
    if L._is_torch(x):
        import torch

        return torch.empty((rows, cols), dtype=dtype or torch.float32, device=x.device)
    return np.empty((rows, cols), dtype=dtype or np.float32)


def _sync_producer(x) -> None:
    """Device tensors are produced on the caller's stream; the engine reads them on its own."""
    if isinstance(x, PackedBits):
        x = x.data
    if L._is_torch(x) and x.is_cuda:
        import torch

        torch.cuda.current_stream(x.device).synchronize()


class Context:
    """One GPU, one stream, optionally one rank of a data-parallel group."""

    _default = None

    def __init__(self, device: int | None = None, seed: int = 42):
        self.lib = L.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        L.check(self.lib.kucd_ctx_create(C.byref(h), int(device), C.c_uint64(seed)))
        self.handle = h
        self.device = int(device)
        self.seed = int(seed)
        self.rank, self.world = 0, 1
        self._children = weakref.WeakSet()  # machines / data sets: they must be destroyed before the context

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = cls()
        return cls._default

    def sync(self) -> None:
        L.check(self.lib.kucd_sync(self.handle))

    def timings(self, reset: bool = False) -> dict:
        t = L.Timings()
        L.check(self.lib.kucd_get_timings(self.handle, C.byref(t), int(reset)))
        return {k: getattr(t, k) for k, _ in L.Timings._fields_}

    def set_profile(self, enable: bool) -> None:
        L.check(self.lib.kucd_ctx_set_profile(self.handle, int(enable)))

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        L.check(self.lib.kucd_ctx_stream(self.handle, C.byref(p)))
        return p.value

    def join_group(self, rank: int | None = None, world: int | None = None) -> None:
        """Attach this context to the data-parallel group of the running torch.distributed job: rank 0
        creates the NCCL id, torch.distributed ships its 128 bytes, every rank opens the communicator."""
        import torch
        import torch.distributed as dist

        if rank is None:
            rank = dist.get_rank()
        if world is None:
            world = dist.get_world_size()
        if world == 1:
            return
        buf = (C.c_char * 128)()
        if rank == 0:
            L.check(self.lib.kucd_comm_unique_id(buf))
        t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda(self.device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
        L.check(self.lib.kucd_ctx_comm_init(self.handle, raw, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def close(self) -> None:
        if self.handle:
            for child in list(self._children):
                child.close()
            self.lib.kucd_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Dataset:
    """A (rows, dim) matrix resident on the GPU in the engine's operand layout (bf16 term planes)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.handle = ctx, handle
        ctx._children.add(self)

    @classmethod
    def from_array(cls, ctx: Context, data, compute: int) -> "Dataset":
        keep: list = []
        _sync_producer(data)
        t = L.tensor_of(data, keep)
        h = C.c_void_p()
        L.check(ctx.lib.kucd_dataset_create(ctx.handle, C.byref(t), int(compute), C.byref(h)))
        return cls(ctx, h)

    @property
    def shape(self):
        r, d = C.c_int64(), C.c_int64()
        L.check(self.ctx.lib.kucd_dataset_shape(self.handle, C.byref(r), C.byref(d)))
        return (r.value, d.value)

    def shuffled(self, seed: int, epoch: int, into: "Dataset | None" = None) -> "Dataset":
        """The rows in the pseudo-random order of (seed, epoch) - include/kucd.h:kucd_dataset_shuffle.  `into`
        (a data set this method returned earlier) is overwritten instead of allocating a new one."""
        h = C.c_void_p(into.handle.value) if into is not None else C.c_void_p()
        L.check(self.ctx.lib.kucd_dataset_shuffle(self.handle, C.c_uint64(seed), C.c_uint64(epoch), C.byref(h)))
        return into if into is not None else Dataset(self.ctx, h)

    def packed(self) -> PackedBits:
        """A 0/1 data set read back one bit per unit."""
        rows, dim = self.shape
        out = PackedBits(np.empty((rows, (dim + 7) // 8), dtype=np.uint8), dim)
        keep: list = []
        t = L.tensor_of(out, keep)
        L.check(self.ctx.lib.kucd_dataset_read(self.handle, C.byref(t)))
        return out

    def numpy(self) -> np.ndarray:
        rows, dim = self.shape
        out = np.empty((rows, dim), dtype=np.float32)
        keep: list = []
        t = L.tensor_of(out, keep)
        L.check(self.ctx.lib.kucd_dataset_read(self.handle, C.byref(t)))
        return out

    def close(self) -> None:
        if self.handle and self.ctx.handle:
            self.ctx.lib.kucd_dataset_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Machine:
    """Device state of one RBM: parameters, operand planes, chains, workspaces, captured step graph."""

    def __init__(self, ctx: Context, n_visible: int, n_hidden: int, mode: int, compute: int, seed: int | None = None):
        self.ctx = ctx
        self.V, self.H, self.mode, self.compute = int(n_visible), int(n_hidden), int(mode), int(compute)
        h = C.c_void_p()
        L.check(ctx.lib.kucd_rbm_create(ctx.handle, self.V, self.H, self.mode, self.compute, C.byref(h)))
        self.handle = h
        ctx._children.add(self)
        if seed is not None:
            self.set_seed(seed, 0)
        self.fused_reduce = False
        if ctx.world > 1 and self.compute == L.COMPUTE_BF16 and ctx.world <= 8 and \
                os.environ.get("KUCD_FUSED_REDUCE", "1") != "0":
            self._attach_peers()

    def _attach_peers(self) -> None:
        """Collective over the data-parallel group: exchange the CUDA IPC handles of the exchange buffers so that
        the dW contraction can store its rows straight into their owners' memory (include/kucd.h, fused reduction)."""
        import torch.distributed as dist

        mine = (C.c_char * 128)()
        ok = self.ctx.lib.kucd_rbm_peer_export(self.handle, mine) == L.KUCD_OK
        everyone = [None] * self.ctx.world
        dist.all_gather_object(everyone, bytes(mine) if ok else b"")
        ok = all(len(h) == 128 for h in everyone)
        if ok:
            ok = self.ctx.lib.kucd_rbm_peer_attach(self.handle, b"".join(everyone)) == L.KUCD_OK
        verdicts = [None] * self.ctx.world
        dist.all_gather_object(verdicts, bool(ok))
        if all(verdicts):
            self.fused_reduce = True
        else:  # some rank could not map its peers (no P2P / IPC): everybody stays on the NCCL all-reduce
            L.check(self.ctx.lib.kucd_rbm_peer_detach(self.handle))

    def set_seed(self, seed: int, step_count: int = 0) -> None:
        L.check(self.ctx.lib.kucd_rbm_set_seed(self.handle, C.c_uint64(seed), C.c_uint64(step_count)))

    def counters(self) -> dict:
        seed, step, nch = C.c_uint64(), C.c_uint64(), C.c_int64()
        L.check(self.ctx.lib.kucd_rbm_get_counters(self.handle, C.byref(seed), C.byref(step), C.byref(nch)))
        return {"seed": seed.value, "step_count": step.value, "n_chains": nch.value}

    def draw_counters(self) -> dict:
        """Positions of the inference and score Philox streams (set_seed resets both to 0)."""
        a, b = C.c_uint64(), C.c_uint64()
        L.check(self.ctx.lib.kucd_rbm_get_draw_counters(self.handle, C.byref(a), C.byref(b)))
        return {"infer_draws": a.value, "score_draws": b.value}

    def set_draw_counters(self, infer_draws: int = 0, score_draws: int = 0) -> None:
        L.check(self.ctx.lib.kucd_rbm_set_draw_counters(self.handle, C.c_uint64(infer_draws), C.c_uint64(score_draws)))

    def get_momentum(self):
        """(mW, mb, mc) of this rank, or None while no momentum step has run (include/kucd.h: kucd_rbm_get_momentum)."""
        present = C.c_int(0)
        L.check(self.ctx.lib.kucd_rbm_get_momentum(self.handle, None, None, None, C.byref(present)))
        if not present.value:
            return None
        mW = np.empty((self.V, self.H), np.float32)
        mb = np.empty((self.V,), np.float32)
        mc = np.empty((self.H,), np.float32)
        keep: list = []
        ts = [L.tensor_of(x, keep) for x in (mW, mb, mc)]
        L.check(self.ctx.lib.kucd_rbm_get_momentum(self.handle, *[C.byref(t) for t in ts], None))
        return mW, mb, mc

    def set_momentum(self, mW, mb, mc) -> None:
        keep: list = []
        ts = [L.tensor_of(np.ascontiguousarray(x, dtype=np.float32), keep) for x in (mW, mb, mc)]
        L.check(self.ctx.lib.kucd_rbm_set_momentum(self.handle, *[C.byref(t) for t in ts]))

    # ---- parameters ----
    def set_params(self, W=None, b=None, c=None) -> None:
        keep: list = []
        ts = [None if x is None else L.tensor_of(np.ascontiguousarray(x, dtype=np.float32), keep) for x in (W, b, c)]
        ptrs = [None if t is None else C.byref(t) for t in ts]
        L.check(self.ctx.lib.kucd_rbm_set_params(self.handle, *ptrs))

    def get_params(self):
        W = np.empty((self.V, self.H), np.float32)
        b = np.empty((self.V,), np.float32)
        c = np.empty((self.H,), np.float32)
        keep: list = []
        tW, tb, tc = (L.tensor_of(x, keep) for x in (W, b, c))
        L.check(self.ctx.lib.kucd_rbm_get_params(self.handle, C.byref(tW), C.byref(tb), C.byref(tc)))
        return W, b, c

    # ---- inference ----
    def _sample(self, fn, x, n_out: int, u, want_p: bool, out_dtype):
        keep: list = []
        _sync_producer(x)
        tin = L.tensor_of(x, keep)
        rows = tin.shape[0]
        out = _out_like(x, rows, n_out, out_dtype)
        tout = L.tensor_of(out, keep)
        p = _out_like(x, rows, n_out) if want_p else None
        tp = L.tensor_of(p, keep) if want_p else None
        tu = L.tensor_of(u, keep) if u is not None else None
        L.check(fn(self.handle, C.byref(tin), C.byref(tout), C.byref(tp) if want_p else None,
                   C.byref(tu) if tu is not None else None))
        return (out, p) if want_p else out

    def transform(self, v, u=None, want_p: bool = False, out_dtype=None):
        return self._sample(self.ctx.lib.kucd_rbm_transform, v, self.H, u, want_p, out_dtype)

    def inv_transform(self, h, u=None, want_p: bool = False, out_dtype=None):
        return self._sample(self.ctx.lib.kucd_rbm_inv_transform, h, self.V, u, want_p, out_dtype)

    def free_energy(self, v):
        keep: list = []
        _sync_producer(v)
        tin = L.tensor_of(v, keep)
        out = _out_like(v, tin.shape[0], 1)
        tout = L.tensor_of(out, keep)
        L.check(self.ctx.lib.kucd_rbm_free_energy(self.handle, C.byref(tin), C.byref(tout)))
        return out.reshape(-1)

    # ---- training ----
    @staticmethod
    def hparams(lr=1e-3, k=1, persistent=False, momentum=0.0, weight_decay=0.0, normalize=False,
                update_mask=L.UPDATE_ALL, want_stats=False) -> L.HParams:
        return L.HParams(float(lr), int(k), int(bool(persistent)), float(momentum), float(weight_decay),
                         int(bool(normalize)), int(update_mask), int(want_stats))

    def cd_step(self, v_batch, hp: L.HParams, u_h=None, u_v=None, u_hc=None, global_row0: int = 0):
        """u_h: list indexed by t (0 = h_pos, t = intermediate h), u_v: list indexed by t (1..k; entry 0
        ignored).  Returns a dict of stats when hp.want_stats."""
        keep: list = []
        _sync_producer(v_batch)
        tv = L.tensor_of(v_batch, keep)
        inj = None
        if u_h is not None or u_v is not None or u_hc is not None:
            inj = L.Inject()
            for t, u in enumerate(u_h or []):
                if u is not None:
                    tt = L.tensor_of(u, keep)
                    keep.append(tt)
                    inj.u_h[t] = C.pointer(tt)
            for t, u in enumerate(u_v or []):
                if u is not None and t >= 1:
                    tt = L.tensor_of(u, keep)
                    keep.append(tt)
                    inj.u_v[t] = C.pointer(tt)
            if u_hc is not None:
                tt = L.tensor_of(u_hc, keep)
                keep.append(tt)
                inj.u_hc = C.pointer(tt)
        st = L.StepStats()
        L.check(self.ctx.lib.kucd_rbm_cd_step(self.handle, C.byref(tv), C.byref(hp), C.byref(inj) if inj else None,
                                              C.c_int64(global_row0), C.byref(st) if hp.want_stats else None))
        if hp.want_stats:
            return {"score": st.score, "recon_err": st.recon_err, "fe_mean": st.fe_mean, "rows": st.rows}
        return None

    def score(self, v_batch, u_h=None, u_v=None) -> float:
        keep: list = []
        _sync_producer(v_batch)
        tv = L.tensor_of(v_batch, keep)
        th = L.tensor_of(u_h, keep) if u_h is not None else None
        tvv = L.tensor_of(u_v, keep) if u_v is not None else None
        out = C.c_float()
        L.check(self.ctx.lib.kucd_rbm_score(self.handle, C.byref(tv), C.byref(th) if th is not None else None,
                                            C.byref(tvv) if tvv is not None else None, C.byref(out)))
        return out.value

    def last_stats(self, rows: int, states: bool = True, grads: bool = True) -> dict:
        keep: list = []
        out = {}
        if grads:
            dW = np.empty((self.V, self.H), np.float32)
            db = np.empty((self.V,), np.float32)
            dc = np.empty((self.H,), np.float32)
            args = [C.byref(L.tensor_of(x, keep)) for x in (dW, db, dc)]
            out.update(dW=dW, db=db, dc=dc)
        else:
            args = [None, None, None]
        if states:
            hp_, vn, hn = (np.empty((rows, self.H), np.float32), np.empty((rows, self.V), np.float32),
                           np.empty((rows, self.H), np.float32))
            args += [C.byref(L.tensor_of(x, keep)) for x in (hp_, vn, hn)]
            out.update(h_pos=hp_, v_neg=vn, h_neg=hn)
        else:
            args += [None, None, None]
        L.check(self.ctx.lib.kucd_rbm_last_stats(self.handle, *args))
        return out

    def delta_rule(self, forward: bool, x, target, lr: float, normalize: bool = False) -> None:
        """One delta-rule step of the directed layer that shares this RBM's parameters (include/kucd.h:
        kucd_rbm_delta_rule): forward - p = sigmoid(x.W + c), W += lr x^T (target - p), c += lr sum (target - p);
        backward - p = sigmoid(x.W^T + b), W += lr (target - p)^T x, b += lr sum (target - p)."""
        keep: list = []
        _sync_producer(x)
        _sync_producer(target)
        tx, tt = L.tensor_of(x, keep), L.tensor_of(target, keep)
        L.check(self.ctx.lib.kucd_rbm_delta_rule(self.handle, int(bool(forward)), C.byref(tx), C.byref(tt),
                                                 C.c_float(lr), int(bool(normalize))))

    def set_chains(self, v) -> None:
        keep: list = []
        _sync_producer(v)
        t = L.tensor_of(v, keep)
        L.check(self.ctx.lib.kucd_rbm_set_chains(self.handle, C.byref(t)))

    def get_chains(self, n: int) -> np.ndarray:
        out = np.empty((n, self.V), np.float32)
        keep: list = []
        t = L.tensor_of(out, keep)
        L.check(self.ctx.lib.kucd_rbm_get_chains(self.handle, C.byref(t)))
        return out

    def fit_epoch(self, ds: Dataset, batch: int, hp: L.HParams, global_row0: int = 0, want_stats: bool = True) -> dict:
        st = L.EpochStats()
        L.check(self.ctx.lib.kucd_rbm_fit_epoch(self.handle, ds.handle, C.c_int64(batch), C.byref(hp),
                                                C.c_int64(global_row0), C.byref(st) if want_stats else None))
        return {k: getattr(st, k) for k, _ in L.EpochStats._fields_}

    def fit_range(self, ds: Dataset, batch: int, hp: L.HParams, step_begin: int, step_end: int,
                  global_row0: int = 0, want_stats: bool = False) -> dict:
        st = L.EpochStats()
        L.check(self.ctx.lib.kucd_rbm_fit_range(self.handle, ds.handle, C.c_int64(batch), C.byref(hp),
                                                C.c_int64(global_row0), C.c_int64(step_begin), C.c_int64(step_end),
                                                C.byref(st) if want_stats else None))
        return {k: getattr(st, k) for k, _ in L.EpochStats._fields_}

    def fit_host(self, V, batch: int, hp: L.HParams, global_row0: int = 0, want_recon: bool = True) -> dict:
        """One pass over a host array, copies overlapped with compute; returns epoch stats + per-step recon."""
        keep: list = []
        t = L.tensor_of(V, keep)
        steps = (t.shape[0] + batch - 1) // batch
        recon = np.zeros(max(steps, 1), np.float32)
        st = L.EpochStats()
        L.check(self.ctx.lib.kucd_rbm_fit_host(self.handle, C.byref(t), C.c_int64(batch), C.byref(hp),
                                               C.c_int64(global_row0),
                                               recon.ctypes.data_as(C.POINTER(C.c_float)) if want_recon else None,
                                               C.byref(st)))
        out = {k: getattr(st, k) for k, _ in L.EpochStats._fields_}
        out["step_recon_err"] = recon[:steps]
        return out

    def transform_dataset(self, ds: Dataset) -> Dataset:
        h = C.c_void_p()
        L.check(self.ctx.lib.kucd_rbm_transform_dataset(self.handle, ds.handle, C.byref(h)))
        return Dataset(self.ctx, h)

    def inv_transform_dataset(self, ds: Dataset) -> Dataset:
        h = C.c_void_p()
        L.check(self.ctx.lib.kucd_rbm_inv_transform_dataset(self.handle, ds.handle, C.byref(h)))
        return Dataset(self.ctx, h)

    def close(self) -> None:
        if self.handle and self.ctx.handle:
            self.ctx.lib.kucd_rbm_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
