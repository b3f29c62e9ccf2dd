"""CPU oracle for the RBM/DBN contrastive-divergence path of keras_unsupervised.  TEST INFRASTRUCTURE.

This file is the checker, not the product: only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import it.  keras_unsupervised_b200 never does; it fails loudly
when libkucd.so is missing.

What it restates (file:line are relative to /root/reference):
  * ku/ebm/rbm.py:30-40      parameter shapes              -> OracleRBM.__init__
  * ku/ebm/rbm.py:45-48      transform  h = 1[u < s(vW+c)] -> sample_h
  * ku/ebm/rbm.py:51-54      inv_transform                 -> sample_v
  * ku/ebm/rbm.py:55-67      Gaussian-visible variants     -> sample_h / sample_v (mode = 1)
  * ku/ebm/rbm.py:73-76      free energy                   -> free_energy
  * ku/ebm/rbm.py:119-134    CD-1 statistics and updates   -> cd_stats / apply
  * ku/ebm/rbm.py:214-233    three sequential single-parameter runs + score -> reference_step
  * ku/ebm/rbm.py:110-111,163,211,218  minibatch slicing, remainder last, no shuffle -> batches / fit
  * ku/ebm/dbn.py:14-32,44-55,65-75,85-96  stacking (with the defect fixes of SURVEY.md 2.3) -> OracleDBN

PINNING.  The reference ships no tests or golden vectors and its arithmetic runs inside TensorFlow,
which is not installable here, so TensorFlow's own kernels are unpinned.  What IS pinned: the op
sequence of ku/ebm/rbm.py itself.  tests/golden/make_reference_fixtures.py imports the unmodified
/root/reference/ku/ebm/rbm.py on top of a numpy stand-in for the handful of keras-backend calls it
makes (tests/golden/kshim.py), runs RBM.build / RBM.fit / transform / inv_transform / cal_free_energy with
recorded random draws, and freezes inputs + outputs in tests/golden/ref_rbm_*.npz;
tests/test_oracle_vs_reference.py replays the same draws through this oracle and demands equality.

Every random draw is an explicit argument.  `philox_uniform` reproduces the engine's counter-based
stream bit for bit, so non-injected runs are comparable as well.
"""
from __future__ import annotations

import numpy as np

MODE_VISIBLE_BERNOULLI = 0  # rbm.py:14
MODE_VISIBLE_GAUSSIAN = 1   # rbm.py:15

F32 = np.float32


# --------------------------------------------------------------------------------------------
# number formats
# --------------------------------------------------------------------------------------------
def bf16_round(x) -> np.ndarray:
    """float32 -> nearest bfloat16 (ties to even), returned as float32."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    bits = a.view(np.uint32).astype(np.uint64)
    rounded = ((bits + 0x7FFF + ((bits >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    nan = np.isnan(a)
    if nan.any():
        out[nan] = np.nan
    return out.reshape(a.shape)


def split3(x):
    """x = hi + mid + lo, each a bfloat16: how the engine's f32x3 mode carries fp32 operands."""
    x = np.asarray(x, dtype=np.float32)
    hi = bf16_round(x)
    r1 = (x - hi).astype(np.float32)
    mid = bf16_round(r1)
    lo = bf16_round((r1 - mid).astype(np.float32))
    return hi, mid, lo


def stat_grid_round(s, bound: int) -> np.ndarray:
    """The engine's rounding of a partial column sum before its fp32 atomicAdd (csrc/rng_math.cuh: stat_grid_round): to the
    grid 2^(ceil(log2 bound) - 24), the fp32 spacing just below `bound`.  Every sum of grid values whose magnitude stays
    within `bound` is exact in fp32, so such additions commute (tests/test_oracle.py).  The oracle's own statistics are exact
    sums; this restatement exists to check that claim and the size of the rounding."""
    e = int(np.ceil(np.log2(max(int(bound), 2))))
    g = np.float32(2.0 ** (e - 24))
    return (np.rint(np.asarray(s, np.float32) / g) * g).astype(np.float32)


def lattice_uniform(rng: np.random.Generator, shape) -> np.ndarray:
    """float32 uniforms on the 2^-23 lattice in [0,1): the values tf.random.uniform can produce."""
    return (rng.integers(0, 1 << 23, size=shape, dtype=np.uint32).astype(np.float32) * F32(2.0 ** -23)).astype(F32)


# --------------------------------------------------------------------------------------------
# Philox4x32-10, as keras_unsupervised_b200/csrc/rng_math.cuh keys it
# --------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _MASK, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_bits(seed: int, draw: int, row0: int, rows: int, cols: int) -> np.ndarray:
    """(rows, cols) uint32 words: word (r, c) = component c % 4 of philox(counter = (c // 4, row0 + r,
    draw lo, draw hi), key = seed)."""
    ncol4 = (cols + 3) // 4
    c0 = np.arange(ncol4, dtype=np.uint64)[None, :]
    c1 = (np.arange(rows, dtype=np.uint64) + np.uint64(row0))[:, None]
    x = philox4x32_10(c0, c1, np.uint64(draw & 0xFFFFFFFF), np.uint64((draw >> 32) & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.stack(x, axis=-1).reshape(rows, ncol4 * 4)[:, :cols]
    return out.astype(np.uint32)


def philox_uniform(seed: int, draw: int, row0: int, rows: int, cols: int) -> np.ndarray:
    """The engine's uniforms: 23 high-order random bits on the 2^-23 lattice."""
    return ((philox_bits(seed, draw, row0, rows, cols) >> np.uint32(9)).astype(np.float32) * F32(2.0 ** -23)).astype(F32)


def philox_normal(seed: int, draw: int, row0: int, rows: int, cols: int) -> np.ndarray:
    """The engine's unit normals (Gaussian-visible mode): Box-Muller on word pairs (x,y) and (z,w)."""
    cols4 = (cols + 3) // 4 * 4
    bits = philox_bits(seed, draw, row0, rows, cols4).reshape(rows, cols4 // 2, 2)
    u1 = ((bits[..., 0] >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0 ** -23
    u2 = (bits[..., 1] >> np.uint32(9)).astype(np.float64) * 2.0 ** -23
    rad = np.sqrt(-2.0 * np.log(u1))
    n = np.stack([rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)], axis=-1)
    return n.reshape(rows, cols4)[:, :cols].astype(np.float32)


def draw_id(kind: str, n: int = 0, phase: int = 0) -> int:
    """Draw ids of include/kucd.h (kucd_rbm_set_seed)."""
    if kind == "train":
        return 64 * n + phase
    if kind == "infer":
        return (1 << 63) + n
    if kind == "score":
        return (1 << 62) + 2 * n + phase
    raise ValueError(kind)


# --------------------------------------------------------------------------------------------
# epoch shuffling (an extension; the reference never shuffles, rbm.py:218) and bit-packed data
# --------------------------------------------------------------------------------------------
def _fmix32(h):
    """murmur3's 32-bit finaliser on uint64 arrays holding 32-bit values."""
    h = h & _MASK
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _MASK
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _MASK
    h ^= h >> np.uint64(16)
    return h


def feistel_keys(seed: int, epoch: int):
    """Six round keys: Philox4x32-10(counter = (0x53485546 'SHUF', j, epoch lo, epoch hi), key = seed), j = 0, 1."""
    words = []
    for j in (0, 1):
        x = philox4x32_10(np.uint64(0x53485546), np.uint64(j), np.uint64(epoch & 0xFFFFFFFF),
                          np.uint64((epoch >> 32) & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        words += [int(w) for w in x]
    return words[:6]


def feistel_permutation(n: int, seed: int, epoch: int) -> np.ndarray:
    """perm[i] = source row of shuffled row i (include/kucd.h:kucd_dataset_shuffle): a 6-round balanced Feistel
    network over the smallest even-width power of two >= n, round function fmix32(R ^ k_r), walked until the value
    lands below n (cycle walking: the restriction of a bijection of [0, 2^w) to [0, n) stays a bijection)."""
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    bits = 2
    while (1 << bits) < n:
        bits += 1
    bits += bits & 1
    hb = np.uint64(bits // 2)
    mask = np.uint64((1 << (bits // 2)) - 1)
    keys = [np.uint64(k) for k in feistel_keys(seed, epoch)]
    x = np.arange(n, dtype=np.uint64)
    todo = np.ones(n, dtype=bool)  # every index takes at least one pass
    while todo.any():
        v = x[todo]
        L, R = (v >> hb) & mask, v & mask
        for k in keys:
            L, R = R, L ^ (_fmix32(R ^ k) & mask)
        v = (L << hb) | R
        x[todo] = v
        todo[todo] = v >= np.uint64(n)
    return x.astype(np.int64)


def pack_bits(x) -> np.ndarray:
    """0/1 matrix -> bytes, column j = bit j % 8 of byte j // 8 (the layout of include/kucd.h's packed tensors)."""
    x = np.asarray(x)
    rows, cols = x.shape
    out = np.zeros((rows, (cols + 7) // 8), dtype=np.uint8)
    for j in range(cols):
        out[:, j // 8] |= ((x[:, j] != 0).astype(np.uint8) << np.uint8(j % 8))
    return out


def unpack_bits(packed, cols: int) -> np.ndarray:
    packed = np.asarray(packed, dtype=np.uint8)
    out = np.zeros((packed.shape[0], cols), dtype=np.float32)
    for j in range(cols):
        out[:, j] = (packed[:, j // 8] >> np.uint8(j % 8)) & np.uint8(1)
    return out


# --------------------------------------------------------------------------------------------
# elementwise maths in the working precision (float32 like K.floatx(), rbm.py:39)
# --------------------------------------------------------------------------------------------
def sigmoid(x):
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(over="ignore"):
        return (F32(1.0) / (F32(1.0) + np.exp(-x))).astype(F32)


def softplus(x):
    """log(1 + exp(x)) (rbm.py:74), evaluated without the overflow of the naive form; the two agree
    wherever the naive one is finite."""
    x = np.asarray(x, dtype=np.float64)
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


class OracleRBM:
    """The intended arithmetic of ku.ebm.RBM.

    compute:
      'f32'   float32 matmuls, the reference's own precision (numpy/BLAS accumulation order)
      'f64'   float32 parameters, float64 accumulation: order-independent stand-in for 'f32'
      'bf16'  what the engine's bf16 mode computes: W, real-valued inputs and the h_neg operand of the
              dW contraction rounded to bfloat16 (RNE) before the products, exact accumulation; the
              bias statistic dc sums the unrounded fp32 probabilities
    """

    def __init__(self, W, b, c, mode=MODE_VISIBLE_BERNOULLI, compute="f64"):
        self.W = np.array(W, dtype=np.float32)      # rbm_weight (V,H)        rbm.py:30-33
        self.b = np.array(b, dtype=np.float32)      # rbm_visible_bias (V,)   rbm.py:38-40
        self.c = np.array(c, dtype=np.float32)      # rbm_hidden_bias (H,)    rbm.py:34-37
        self.V, self.H = self.W.shape
        assert self.b.shape == (self.V,) and self.c.shape == (self.H,)
        self.mode = mode
        assert compute in ("f32", "f64", "bf16")
        self.compute = compute
        self.mW = self.mb = self.mc = None
        self.chains = None

    @staticmethod
    def init_params(V, H, seed=0):
        """Keras 'uniform' = RandomUniform(-0.05, 0.05), float32 (rbm.py:32,36,38)."""
        rng = np.random.default_rng(seed)
        return (rng.uniform(-0.05, 0.05, (V, H)).astype(F32), rng.uniform(-0.05, 0.05, V).astype(F32),
                rng.uniform(-0.05, 0.05, H).astype(F32))

    # ---- contractions ----
    def _mm(self, a, w):
        a = np.asarray(a, dtype=np.float32)
        if self.compute == "f32":
            return (a @ w).astype(F32)
        if self.compute == "bf16":
            a, w = bf16_round(a), bf16_round(w)
        return (a.astype(np.float64) @ w.astype(np.float64)).astype(F32)

    def pre_h(self, v):  # K.dot(v, W) + c      rbm.py:47
        return (self._mm(v, self.W) + self.c).astype(F32)

    def pre_v(self, h):  # K.dot(h, W^T) + b    rbm.py:53
        return (self._mm(h, self.W.T) + self.b).astype(F32)

    def prob_h(self, v):
        x = self.pre_h(v)
        return sigmoid(x)

    def prob_v(self, h):
        return sigmoid(self.pre_v(h))

    # ---- sampling nodes ----
    def sample_h(self, v, u):
        """rbm.py:46-47 (Bernoulli) / :58-59 (Gaussian: relu in place of sigmoid).  Strict <."""
        p = self.prob_h(v) if self.mode == MODE_VISIBLE_BERNOULLI else np.maximum(self.pre_h(v), F32(0))
        return (np.asarray(u, F32) < p).astype(F32), p

    def sample_v(self, h, u):
        """rbm.py:52-53 (Bernoulli).  Gaussian (rbm.py:64-66): mean + unit normal; `u` carries the normals."""
        if self.mode == MODE_VISIBLE_BERNOULLI:
            p = self.prob_v(h)
            return (np.asarray(u, F32) < p).astype(F32), p
        mean = self.pre_v(h)
        return (mean + np.asarray(u, F32)).astype(F32), mean

    def free_energy(self, v):
        """rbm.py:73-75:  F = -( v.b + sum_j log(1 + exp((vW + c)_j)) )."""
        v = np.asarray(v, dtype=np.float32)
        vb = v.astype(np.float64) @ self.b.astype(np.float64) if self.compute != "bf16" else \
            bf16_round(v).astype(np.float64) @ self.b.astype(np.float64)
        return (-(vb + softplus(self.pre_h(v)).sum(axis=-1))).astype(F32)

    # ---- CD statistics ----
    def cd_stats(self, v, u_h, u_v, k=1, persistent=False, u_hc=None, wire_shards=0, wire_sum_bf16=False, need=7):
        """rbm.py:119-126,131,134 generalised to CD-k / PCD.

        wire_shards = n > 0 models the engine's opt-in "bf16 partial sums on the wire" exchange: the minibatch rows
        are n contiguous shards (one per data-parallel rank), each shard's part of dW is rounded to bf16 (RNE) and
        the n parts are summed in rank order in float32.  wire_sum_bf16 additionally rounds the running sum to bf16
        after every addition: a bf16 all-reduce (exact model for two ranks; for more, NCCL's order is its own).

        u_h[0] draws h_pos; u_v[t] (t = 1..k) draws the t-th v_neg; u_h[t] (t = 1..k-1) the intermediate
        hidden samples; the final hidden term is the probability (rbm.py:124).  With `persistent` the
        negative chain starts from self.chains[:rows] (first hidden sample drawn with u_hc) and the final
        v_neg is stored back.

        need (bits as in `apply`: 1 = dW, 2 = dc, 4 = db): the statistics to compute.  The reference builds one graph
        per parameter (rbm.py:127-134) and each evaluates only what its update depends on: the c graph stops after
        h_neg (no outer products), the b graph after v_neg (no final hidden projection either) - 5 + 3 + 2 contractions
        per minibatch, which is what its CPU time is made of."""
        v = np.asarray(v, dtype=np.float32)
        rows = v.shape[0]
        # row_margin: the smallest |u - p| over every Bernoulli draw of a row's chain.  A row whose margin exceeds the
        # rounding difference between two sigmoid implementations has a uniquely determined chain (tests use it to
        # tell apart "differs from the oracle" and "a draw sat inside the rounding gap")
        margin = np.full(rows, np.inf)

        def note(u, p):
            np.minimum(margin, np.abs(np.asarray(u, np.float64) - p).min(axis=1), out=margin)

        h_pos, p_h_pos = self.sample_h(v, u_h[0])
        note(u_h[0], p_h_pos)
        if persistent:
            h, p_ = self.sample_h(self.chains[:rows], u_hc)
            note(u_hc, p_)
        else:
            h = h_pos
        v_neg = h_neg = None
        for t in range(1, k + 1):
            v_neg, p_ = self.sample_v(h, u_v[t])
            if self.mode == MODE_VISIBLE_BERNOULLI:
                note(u_v[t], p_)
            if t < k:
                h, p_ = self.sample_h(v_neg, u_h[t])
                note(u_h[t], p_)
            elif need & 3:
                h_neg = self.prob_h(v_neg)  # always the logistic, also in Gaussian mode (rbm.py:145)
        if persistent:
            self.chains[:rows] = v_neg
        if not need & 3:
            db = (v.astype(np.float64).sum(0) - v_neg.astype(np.float64).sum(0)).astype(F32) if self.compute != "bf16" \
                else (bf16_round(v).astype(np.float64).sum(0) - bf16_round(v_neg).astype(np.float64).sum(0)).astype(F32)
            return dict(h_pos=h_pos, p_h_pos=p_h_pos, v_neg=v_neg, h_neg=None, dW=None, dc=None, db=db, rows=rows,
                        row_margin=margin)
        if self.compute == "bf16":
            hn_mm, v0_mm, vn_mm = bf16_round(h_neg), bf16_round(v), bf16_round(v_neg)
        else:
            hn_mm, v0_mm, vn_mm = h_neg, v, v_neg
        f = np.float32 if self.compute == "f32" else np.float64
        if not need & 1:
            dW = None
        elif wire_shards and wire_shards > 0:
            if rows % wire_shards:
                raise ValueError("rows must divide by wire_shards")
            rb = rows // wire_shards
            dW = np.zeros((self.V, self.H), dtype=F32)
            for s in range(wire_shards):
                sl = slice(s * rb, (s + 1) * rb)
                part = v0_mm[sl].astype(f).T @ h_pos[sl].astype(f) - vn_mm[sl].astype(f).T @ hn_mm[sl].astype(f)
                dW = (dW + bf16_round(part.astype(F32))).astype(F32)
                if wire_sum_bf16:
                    dW = bf16_round(dW)
        else:
            dW = (v0_mm.astype(f).T @ h_pos.astype(f) - vn_mm.astype(f).T @ hn_mm.astype(f)).astype(F32)  # :125-126
        dc = (h_pos.astype(f).sum(0) - h_neg.astype(f).sum(0)).astype(F32)  # :131 (the fp32 probabilities)
        db = (v0_mm.astype(f).sum(0) - vn_mm.astype(f).sum(0)).astype(F32)                             # :134
        return dict(h_pos=h_pos, p_h_pos=p_h_pos, v_neg=v_neg, h_neg=h_neg, dW=dW, dc=dc, db=db, rows=rows,
                    row_margin=margin)

    def apply(self, st, lr, mask=7, momentum=0.0, weight_decay=0.0, scale=1.0):
        """rbm.py:127-134: parameter += lr * batch SUM (scale = 1).  Extensions: momentum, weight decay
        (on W only), scale = 1/rows for mean normalisation.  mask bits: 1 = W, 2 = c, 4 = b."""
        lr, momentum, weight_decay, scale = F32(lr), F32(momentum), F32(weight_decay), F32(scale)

        def step(x, d, m, wd):
            s = (lr * (scale * d - wd * x)).astype(F32)
            if momentum != 0:
                m = s if m is None else (momentum * m + s).astype(F32)
                s = m
            return (x + s).astype(F32), m

        if mask & 1:
            self.W, self.mW = step(self.W, st["dW"], self.mW, weight_decay)
        if mask & 2:
            self.c, self.mc = step(self.c, st["dc"], self.mc, F32(0))
        if mask & 4:
            self.b, self.mb = step(self.b, st["db"], self.mb, F32(0))

    def delta_rule(self, forward, x, target, lr, scale=1.0):
        """include/kucd.h: kucd_rbm_delta_rule - one delta-rule step of the directed sigmoid layer that shares this
        RBM's parameters (the reference has no fine-tuning; Hinton, Osindero & Teh 2006, appendix B):
          forward:  p = sigmoid(x.W + c),   W += lr x^T (t - p),  c += lr sum_rows (t - p)
          backward: p = sigmoid(x.W^T + b), W += lr (t - p)^T x,  b += lr sum_rows (t - p)
        Computed as the engine does, as x^T t - x^T p with the operands rounded to bf16 in bf16 mode (p like the
        h_neg operand of rbm.py:126) and the bias statistic from the unrounded probabilities.  Returns p."""
        x, t = np.asarray(x, F32), np.asarray(target, F32)
        p = self.prob_h(x) if forward else self.prob_v(x)
        if self.compute == "bf16":
            x_mm, t_mm, p_mm = bf16_round(x), bf16_round(t), bf16_round(p)
        else:
            x_mm, t_mm, p_mm = x, t, p
        f = np.float32 if self.compute == "f32" else np.float64
        if forward:
            dW = (x_mm.astype(f).T @ t_mm.astype(f) - x_mm.astype(f).T @ p_mm.astype(f)).astype(F32)
        else:
            dW = (t_mm.astype(f).T @ x_mm.astype(f) - p_mm.astype(f).T @ x_mm.astype(f)).astype(F32)
        dbias = (t_mm.astype(f).sum(0) - p.astype(f).sum(0)).astype(F32)
        st = {"dW": dW, "dc": dbias, "db": dbias}
        self.apply(st, lr, 1 | (2 if forward else 4), scale=scale)
        return p

    def fused_step(self, v, u_h, u_v, lr, k=1, **kw):
        """One chain, all three parameters from the same statistics (the engine's timed schedule)."""
        persistent = kw.pop("persistent", False)
        u_hc = kw.pop("u_hc", None)
        wire_shards = kw.pop("wire_shards", 0)
        wire_sum_bf16 = kw.pop("wire_sum_bf16", False)
        st = self.cd_stats(v, u_h, u_v, k=k, persistent=persistent, u_hc=u_hc, wire_shards=wire_shards,
                           wire_sum_bf16=wire_sum_bf16)
        self.apply(st, lr, 7, **kw)
        return st

    def reference_step(self, v, draws, lr, k=1, scale=1.0):
        """rbm.py:214-233 as written: run A updates W, run B (fresh draws, new W) updates c, run C
        (fresh draws, new W and c) updates b, then the score with a fourth chain (run D): 5 + 3 + 2 contractions,
        two free energies and the score chain's two = 14 per minibatch.
        draws = [(u_h, u_v)] * 4.  With k > 1 (the reference only has CD-1) u_h / u_v of runs A-C are the lists
        `cd_stats` takes and every run walks the whole CD-k chain: (2k+3) + (2k+1) + 2k + 4 = 6k + 8 contractions."""
        (ah, av), (bh, bv), (ch, cv), (dh, dv) = draws
        if k == 1 and not isinstance(ah, (list, tuple)):
            ah, av, bh, bv, ch, cv = [ah], [None, av], [bh], [None, bv], [ch], [None, cv]
        self.apply(self.cd_stats(v, ah, av, k=k, need=1), lr, 1, scale=scale)   # rbm_weight_update_func   :214/:221
        self.apply(self.cd_stats(v, bh, bv, k=k, need=2), lr, 2, scale=scale)   # hidden_bias_update_func  :215/:222
        self.apply(self.cd_stats(v, ch, cv, k=k, need=4), lr, 4, scale=scale)   # visible_bias_update_func :216/:223
        return self.score(v, dh, dv)

    def score(self, v, u_h, u_v):
        """rbm.py:227-233: mean |F(v) - F(v_neg)| with a fresh chain."""
        fe = self.free_energy(v)
        h, _ = self.sample_h(v, u_h)
        v_neg, _ = self.sample_v(h, u_v)
        fe_p = self.free_energy(v_neg)
        return float(np.mean(np.abs(fe.astype(np.float64) - fe_p.astype(np.float64))))


def batches(n_rows: int, batch_size: int):
    """rbm.py:110-111,163,211,218: sequential slices, remainder last."""
    num_step = n_rows // batch_size if n_rows % batch_size == 0 else n_rows // batch_size + 1
    for i in range(num_step):
        yield i * batch_size, min((i + 1) * batch_size, n_rows)


def philox_fit(rbm: OracleRBM, V, batch_size, epochs, lr, seed, k=1, persistent=False, step0=0, row0=0,
               normalize=False, **kw):
    """The engine's fit loop with its own Philox draws (fused schedule): what fit_epoch must reproduce.
    normalize: divide the batch sums by the rows of the current minibatch (the remainder one is shorter)."""
    step = step0
    for _ in range(epochs):
        for lo, hi in batches(V.shape[0], batch_size):
            rows = hi - lo
            if rbm.mode == MODE_VISIBLE_BERNOULLI:
                gen_v = philox_uniform
            else:
                gen_v = philox_normal
            u_h = [philox_uniform(seed, draw_id("train", step, 0 if t == 0 else 2 * t + 1), row0, rows, rbm.H)
                   for t in range(k)]
            u_v = [None] + [gen_v(seed, draw_id("train", step, 2 * t), row0, rows, rbm.V) for t in range(1, k + 1)]
            u_hc = philox_uniform(seed, draw_id("train", step, 1), row0, rows, rbm.H) if persistent else None
            if normalize:
                kw["scale"] = 1.0 / rows
            rbm.fused_step(V[lo:hi], u_h, u_v, lr, k=k, persistent=persistent, u_hc=u_hc, **kw)
            step += 1
    return step


def philox_fit_reference(rbm: OracleRBM, V, batch_size, epochs, lr, seed, step0=0, score0=0):
    """The engine's compat='reference' loop with its own Philox draws: per minibatch three single-parameter CD-1 runs
    (training draw ids of three consecutive steps), then the score chain (score draw ids).  Returns the scores."""
    step, nscore, scores = step0, score0, []
    for _ in range(epochs):
        for lo, hi in batches(V.shape[0], batch_size):
            rows = hi - lo
            v = V[lo:hi]
            for mask in (1, 2, 4):
                u_h = [philox_uniform(seed, draw_id("train", step, 0), 0, rows, rbm.H)]
                u_v = [None, philox_uniform(seed, draw_id("train", step, 2), 0, rows, rbm.V)]
                rbm.apply(rbm.cd_stats(v, u_h, u_v), lr, mask)
                step += 1
            scores.append(rbm.score(v, philox_uniform(seed, draw_id("score", nscore, 0), 0, rows, rbm.H),
                                    philox_uniform(seed, draw_id("score", nscore, 1), 0, rows, rbm.V)))
            nscore += 1
    return scores


def condition_margin(rbm: OracleRBM, v, u_h, u_v, k=1, margin=1e-4, rng=None, persistent=False, u_hc=None):
    """Re-draw every injected uniform that falls within `margin` of the probability it is compared with
    (walking the chain in order), so that an implementation whose probabilities differ from the oracle's
    by less than `margin` must produce bit-identical samples.  Returns the adjusted (u_h, u_v, u_hc)."""
    rng = rng or np.random.default_rng(99)
    assert rbm.mode == MODE_VISIBLE_BERNOULLI

    def fix(u, p):
        u = np.array(u, dtype=np.float32)
        while True:
            bad = np.abs(u.astype(np.float64) - p.astype(np.float64)) <= margin
            if not bad.any():
                return u
            u[bad] = lattice_uniform(rng, int(bad.sum()))

    u_h = [None if u is None else np.array(u, F32) for u in u_h]
    u_v = [None if u is None else np.array(u, F32) for u in u_v]
    v = np.asarray(v, F32)
    u_h[0] = fix(u_h[0], rbm.prob_h(v))
    h, _ = rbm.sample_h(v, u_h[0])
    if persistent:
        ch = rbm.chains[:v.shape[0]]
        u_hc = fix(u_hc, rbm.prob_h(ch))
        h, _ = rbm.sample_h(ch, u_hc)
    for t in range(1, k + 1):
        u_v[t] = fix(u_v[t], rbm.prob_v(h))
        vn, _ = rbm.sample_v(h, u_v[t])
        if t < k:
            u_h[t] = fix(u_h[t], rbm.prob_h(vn))
            h, _ = rbm.sample_h(vn, u_h[t])
    return u_h, u_v, u_hc


class OracleDBN:
    """ku/ebm/dbn.py with the fixes of SURVEY.md 2.3 (D7): add_stack validates prev.H == next.V
    (dbn.py:24-30), fit trains each layer on the previous layer's sampled hidden states (dbn.py:51-55),
    inv_transform walks the stack in reverse (dbn.py:92-94)."""

    def __init__(self):
        self.layers = []

    def add_stack(self, rbm: OracleRBM):
        if self.layers and self.layers[-1].H != rbm.V:
            raise ValueError("A previous RBM layer's output dimension must be equal to a next one's input dimension.")
        self.layers.append(rbm)

    def transform(self, V, draws):
        if not self.layers:
            raise ValueError("Any rbm layer doesn't exist.")
        x = np.array(V, F32)
        for rbm, u in zip(self.layers, draws):
            x, _ = rbm.sample_h(x, u)
        return x

    def inv_transform(self, Hs, draws):
        if not self.layers:
            raise ValueError("Any rbm layer doesn't exist.")
        x = np.array(Hs, F32)
        for rbm, u in zip(reversed(self.layers), draws):
            x, _ = rbm.sample_v(x, u)
        return x

    # ---- up-down fine-tuning (an extension: the reference stops at greedy pretraining, dbn.py:34-55) ----
    def untie(self):
        """Give every layer below the top one its own generative parameters (a copy of W, b), as the up-down
        algorithm does before fine-tuning: `layers[l]` keeps the recognition direction (W, c), `gen[l]` the
        generative one (W, b).  The top RBM stays undirected."""
        if not hasattr(self, "gen") or len(self.gen) != len(self.layers) - 1:
            self.gen = [OracleRBM(r.W.copy(), r.b.copy(), r.c.copy(), mode=r.mode, compute=r.compute)
                        for r in self.layers[:-1]]
        return self.gen

    def up_down_step(self, v, draws, lr, k=1, scale=1.0):
        """One minibatch of the up-down algorithm (Hinton, Osindero & Teh 2006, appendix B), what DBN.fine_tune runs:
          wake    s_0 = v;  s_l = 1[u < sigmoid(s_{l-1} W_l + c_l)]                     l = 1 .. L-1   (draws["up"][l-1])
          top     CD-k of the top RBM on s_{L-1} (draws["top"] = (u_h, u_v) as for cd_stats); t_{L-1} = its last v_neg
          sleep   t_{l-1} = 1[u < sigmoid(t_l G_l^T + b_l)]                                l = L-1 .. 1   (draws["down"][l-1])
          generative weights, from the wake states:   gen[l-1].delta_rule(backward, s_l -> s_{l-1})
          recognition weights, from the sleep states: layers[l-1].delta_rule(forward, t_{l-1} -> t_l)
        Returns the wake and sleep states."""
        L = len(self.layers)
        if L < 2:
            raise ValueError("up-down fine-tuning needs at least two layers")
        gen = self.untie()
        s = [np.asarray(v, F32)]
        for l in range(1, L):
            h, _ = self.layers[l - 1].sample_h(s[-1], draws["up"][l - 1])
            s.append(h)
        u_h, u_v = draws["top"]
        st = self.layers[-1].fused_step(s[-1], u_h, u_v, lr, k=k, scale=scale)
        t = [None] * L
        t[L - 1] = st["v_neg"]
        for l in range(L - 1, 0, -1):
            t[l - 1], _ = gen[l - 1].sample_v(t[l], draws["down"][l - 1])
        for l in range(1, L):
            gen[l - 1].delta_rule(False, s[l], s[l - 1], lr, scale=scale)
            self.layers[l - 1].delta_rule(True, t[l - 1], t[l], lr, scale=scale)
        return s, t
