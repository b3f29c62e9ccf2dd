"""ku.ebm.RBM, same surface, B200 engine underneath.

Mirrors /root/reference/ku/ebm/rbm.py: constructor `RBM(hps, output_dim, name=None,
mode=MODE_VISIBLE_GAUSSIAN, **kwargs)` (rbm.py:22), `build` (:29), `call` (:80), `transform` (:88),
`inv_transform` (:91), `compute_output_shape` (:94), `cal_free_energy` (:97), `fit` (:100),
`get_config` (:236), attributes `rbm_weight`, `hidden_bias`, `visible_bias` (:30,34,38).  The defects
that keep the reference from running are resolved the way the surrounding
code intends: draws take the row count of their input, the hidden draw is (rows, H), the remainder
minibatch is the remaining rows, `transform`/`inv_transform` stay methods, `get_config` carries `mode`.

`hps` keeps the reference's keys `batch_size`, `epochs`, `lr` (rbm.py:46,110,113,128).  Optional keys whose
defaults are the reference's arithmetic: `k` (1), `persistent` (False), `momentum` (0), `weight_decay` (0),
`normalize` ('sum' | 'mean'), `shuffle` (False; True: a fresh keyed row permutation per epoch, `shuffle_seed`),
`dtype` ('float32' = fp32-grade three-term contractions | 'bf16'), `seed` (42).
One default is NOT the reference's: `compat` = 'fused' (one chain per minibatch updates W, b and c from the same
statistics; the score is printed once per epoch), where rbm.py:214-216 runs three sequential single-parameter graphs
with fresh draws and rbm.py:233-234 prints the score after every minibatch - `compat='reference'` reproduces that
schedule run for run (single rank).
`stream` (True: a one-epoch fit of a host array is streamed, minibatch copies overlapped with the chains; False: the
array is uploaded first and trained by CUDA-graph replay).

Every array method calls libkucd.so; nothing is computed in Python and there is no fallback.
"""
from __future__ import annotations

import math
import os

import numpy as np

from .. import _lib as L
from ..data import PackedBits
from ..engine import Context, Dataset, Machine
from ..parallel import shard_rows

# Constants (rbm.py:14-16)
MODE_VISIBLE_BERNOULLI = 0
MODE_VISIBLE_GAUSSIAN = 1
MODE_COMPLEX = 2  # TODO in the reference as well


def _keras_layer_base():
    """The reference's RBM is a Keras Layer (rbm.py:19: `class RBM(Layer)`), so that it can sit inside a
    `tf.keras.Model` (examples/rbm/rbm_softmax_mnist.py:48-64).  When TensorFlow is importable the class below
    subclasses `tf.keras.layers.Layer` too; otherwise it is a plain object with the same surface.
    KUCD_KERAS=0 forces the plain class."""
    if os.environ.get("KUCD_KERAS", "1") == "0":
        return object
    try:
        import tensorflow as tf

        return tf.keras.layers.Layer
    except Exception:  # not installed, or broken: the engine itself never needs TensorFlow
        return object


_Base = _keras_layer_base()
IS_KERAS_LAYER = _Base is not object


def _unwrap(x):
    """K.function takes and returns lists of one array (rbm.py:89,211,230); accept both."""
    if isinstance(x, (list, tuple)) and len(x) == 1:
        return x[0]
    return x


class RBM(_Base):
    """Restricted Boltzmann machine with the reference's API."""

    def __init__(self, hps, output_dim, name=None, mode=MODE_VISIBLE_GAUSSIAN, **kwargs):
        if mode not in (MODE_VISIBLE_BERNOULLI, MODE_VISIBLE_GAUSSIAN):
            raise ValueError("mode must be MODE_VISIBLE_BERNOULLI or MODE_VISIBLE_GAUSSIAN "
                             "(MODE_COMPLEX is a TODO in the reference, rbm.py:16,68-70)")
        context = kwargs.pop("context", None)
        return_list = kwargs.pop("return_list", True)
        input_shape = kwargs.pop("input_shape", None)
        input_dim = kwargs.pop("input_dim", None)
        if IS_KERAS_LAYER:
            # rbm.py:25-27 with defect D6 resolved: the name goes through the base class (a Layer's `name`,
            # `input_shape`, `output_shape` are read-only properties and are never assigned here)
            super().__init__(name=name, **kwargs)
        else:
            self.name = name
            self.built = False
            self.input_shape = None
            self.output_shape = None
        self.hps = hps
        self.output_dim = output_dim
        self.mode = mode
        self._context = context
        self.return_list = return_list
        self._kwargs = kwargs
        self._machine = None
        self._n_visible = None
        self._keras_vars = None  # (rbm_weight, rbm_hidden_bias, rbm_visible_bias) Keras variables, Layer mode only
        self.history = []
        self._epochs_done = 0  # epochs trained so far: numbers the shuffling permutations
        if input_shape is not None:
            self.build((None,) + tuple(input_shape))
        elif input_dim is not None:
            self.build((None, int(input_dim)))

    # ---- Keras-layer surface -------------------------------------------------------------------
    def build(self, input_shape):
        """rbm.py:29-40: W (V,H), c (H,), b (V,) ~ Keras 'uniform' = U(-0.05, 0.05), float32."""
        n_visible = int(input_shape[1])
        ctx = self._context or Context.default()
        dtype = str(self.hps.get("dtype", "float32")).lower()
        if dtype in ("float32", "fp32", "f32", "f32x3"):
            compute = L.COMPUTE_F32X3
        elif dtype in ("bf16", "bfloat16"):
            compute = L.COMPUTE_BF16
        else:
            raise ValueError("hps['dtype'] must be 'float32' or 'bf16'")
        seed = self.seed
        self._machine = Machine(ctx, n_visible, int(self.output_dim), int(self.mode), compute, seed=seed)
        if IS_KERAS_LAYER:
            # the reference's variables, under its names (rbm.py:30-40); the engine trains its own copies on the GPU and
            # writes them back after every fit (_to_keras), so Model.save / save_weights see the trained values
            kW = self.add_weight(name="rbm_weight", shape=(n_visible, int(self.output_dim)), initializer="uniform",
                                 trainable=True)
            kc = self.add_weight(name="rbm_hidden_bias", shape=(int(self.output_dim),), initializer="uniform",
                                 trainable=True)
            kb = self.add_weight(name="rbm_visible_bias", shape=(n_visible,), initializer="uniform", trainable=False)
            self._keras_vars = (kW, kc, kb)
            W, c, b = (np.asarray(v.numpy(), dtype=np.float32) for v in self._keras_vars)
        else:
            rng = np.random.default_rng(seed)
            W = rng.uniform(-0.05, 0.05, (n_visible, int(self.output_dim))).astype(np.float32)
            b = rng.uniform(-0.05, 0.05, n_visible).astype(np.float32)
            c = rng.uniform(-0.05, 0.05, int(self.output_dim)).astype(np.float32)
            self.input_shape = (None, n_visible)
            self.output_shape = (None, int(self.output_dim))
        self._machine.set_params(W, b, c)
        self._n_visible = n_visible
        self.built = True

    def _to_keras(self):
        """engine -> Keras variables (Layer mode): after training, before a Keras save."""
        if self._keras_vars is not None:
            W, b, c = self._machine.get_params()
            for var, val in zip(self._keras_vars, (W, c, b)):
                var.assign(val)

    def sync_from_keras(self):
        """Keras variables -> engine (Layer mode): after `Model.load_weights`, which writes the variables directly."""
        if self._keras_vars is not None:
            W, c, b = (np.asarray(v.numpy(), dtype=np.float32) for v in self._keras_vars)
            self._machine.set_params(W, b, c)
        return self

    @property
    def seed(self):
        """Philox key / initialiser seed: hps['seed'] (42), offset per position in a DBN stack so that stacked
        layers do not share a random stream."""
        return int(self.hps.get("seed", 42)) + 1000003 * int(getattr(self, "_stack_index", 0))

    def _set_stack_index(self, i):
        self._stack_index = int(i)
        if self.built:
            self._machine.set_seed(self.seed, self._machine.counters()["step_count"])

    def _ensure_built(self, x):
        if not self.built:
            shape = x.shape if not isinstance(x, Dataset) else x.shape
            self.build((None, int(shape[1])))

    if not IS_KERAS_LAYER:
        def __call__(self, x):
            return self.call(x)

    def call(self, x):
        """rbm.py:80-86: the layer's forward pass is the sampled hidden state.  Inside a Keras model the input is a
        TensorFlow tensor: the engine is called through tf.numpy_function (no gradient flows through the sampling,
        as in the reference, whose K.less / K.cast have none)."""
        if IS_KERAS_LAYER and type(x).__module__.split(".")[0] in ("tensorflow", "keras", "tf_keras"):
            import tensorflow as tf

            self._ensure_built(x)
            out = tf.numpy_function(lambda a: np.asarray(_unwrap(self.transform(a)), dtype=np.float32), [x], tf.float32)
            if hasattr(out, "set_shape"):
                out.set_shape((None, int(self.output_dim)))
            return out
        return _unwrap(self.transform(x))

    def compute_output_shape(self, input_shape):
        return (input_shape[0], self.output_dim)  # rbm.py:94-95

    def get_config(self):
        """rbm.py:236-242, plus `mode` (the reference drops it and reloads as Gaussian)."""
        return {"hps": self.hps, "output_dim": self.output_dim, "name": self.name, "mode": self.mode}

    @classmethod
    def from_config(cls, config, **kwargs):
        """Inverse of get_config (the Keras convention the reference inherits from Layer)."""
        config = dict(config)
        return cls(config.pop("hps"), config.pop("output_dim"), name=config.pop("name", None),
                   mode=config.pop("mode", MODE_VISIBLE_GAUSSIAN), **kwargs)

    # ---- parameters: same attribute names as the reference -------------------------------------
    @property
    def rbm_weight(self):
        return self._machine.get_params()[0]

    @rbm_weight.setter
    def rbm_weight(self, W):
        self._machine.set_params(W=W)
        self._to_keras()

    @property
    def visible_bias(self):
        return self._machine.get_params()[1]

    @visible_bias.setter
    def visible_bias(self, b):
        self._machine.set_params(b=b)
        self._to_keras()

    @property
    def hidden_bias(self):
        return self._machine.get_params()[2]

    @hidden_bias.setter
    def hidden_bias(self, c):
        self._machine.set_params(c=c)
        self._to_keras()

    def get_weights(self):
        """Keras order of creation: rbm_weight, rbm_hidden_bias, rbm_visible_bias (rbm.py:30,34,38)."""
        W, b, c = self._machine.get_params()
        return [W, c, b]

    def set_weights(self, weights):
        W, c, b = weights
        self._machine.set_params(W, b, c)
        self._to_keras()

    # ---- checkpoint / resume (the reference relies on Keras HDF5 and saves no trainer state) ----
    @staticmethod
    def _rank_path(path, rank, world):
        """Data-parallel ranks hold identical parameters but different persistent chains: rank r > 0 writes its own
        file next to rank 0's (`model.npz`, `model.rank1.npz`, ...)."""
        path = str(path)
        if world <= 1 or rank == 0:
            return path
        stem = path[:-4] if path.endswith(".npz") else path
        return "%s.rank%d.npz" % (stem, rank)

    def save(self, path):
        """Parameters under the reference's variable names (rbm.py:30,34,40) plus what resuming a fit needs for the
        resumed run to equal an uninterrupted one: the positions of the Philox streams (training, inference, score), the
        persistent chains and the momentum buffers.  Under data parallelism every rank calls this; chains and momentum
        are the rank's own (a checkpoint with either is restored into a group of the same size)."""
        W, b, c = self._machine.get_params()
        st = self._machine.counters()
        st.update(self._machine.draw_counters())
        ctx = self._machine.ctx
        extra = {}
        if st["n_chains"] > 0:
            extra["chains"] = self._machine.get_chains(st["n_chains"])
        mom = self._machine.get_momentum()
        if mom is not None:
            extra.update(momentum_weight=mom[0], momentum_visible_bias=mom[1], momentum_hidden_bias=mom[2])
        extra.update(infer_draws=np.uint64(st["infer_draws"]), score_draws=np.uint64(st["score_draws"]))
        np.savez(self._rank_path(path, ctx.rank, ctx.world), rbm_weight=W, rbm_hidden_bias=c, rbm_visible_bias=b,
                 seed=np.uint64(st["seed"]), step_count=np.uint64(st["step_count"]), mode=np.int64(self.mode),
                 output_dim=np.int64(self.output_dim), epochs_done=np.int64(self._epochs_done),
                 rank=np.int64(ctx.rank), world=np.int64(ctx.world), **extra)

    def load(self, path):
        ctx = getattr(getattr(self, "_machine", None), "ctx", None) or self._context or Context.default()
        path = self._rank_path(path, ctx.rank, ctx.world)
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        if "world" in z.files and int(z["world"]) != ctx.world and ("chains" in z.files or "momentum_weight" in z.files):
            raise ValueError("checkpoint with persistent chains or momentum was written by %d ranks, this group has %d"
                             % (int(z["world"]), ctx.world))
        if not self.built:
            self.build((None, int(z["rbm_weight"].shape[0])))
        if z["rbm_weight"].shape != (self._machine.V, self._machine.H):
            raise ValueError("checkpoint has rbm_weight %s, this RBM is %s" % (z["rbm_weight"].shape,
                                                                               (self._machine.V, self._machine.H)))
        self._machine.set_params(z["rbm_weight"], z["rbm_visible_bias"], z["rbm_hidden_bias"])
        self._to_keras()
        self._machine.set_seed(int(z["seed"]), int(z["step_count"]))
        if "infer_draws" in z.files:
            self._machine.set_draw_counters(int(z["infer_draws"]), int(z["score_draws"]))
        if "momentum_weight" in z.files:
            self._machine.set_momentum(z["momentum_weight"], z["momentum_visible_bias"], z["momentum_hidden_bias"])
        if "epochs_done" in z.files:
            self._epochs_done = int(z["epochs_done"])
        if "chains" in z.files:
            self._machine.set_chains(z["chains"])
            self._chains_set = int(z["chains"].shape[0])
        return self

    # ---- inference -----------------------------------------------------------------------------
    def _wrap(self, out):
        return [out] if self.return_list else out

    def transform(self, v, u=None, out_dtype=None):
        """rbm.py:88-89 -> transform_func (:45-48): h = 1[u < sigmoid(v.W + c)], sampled, float32.
        out_dtype='bits' returns the states as data.PackedBits (one bit per unit)."""
        v = _unwrap(v)
        self._ensure_built(v)
        if isinstance(v, Dataset):
            return self._machine.transform_dataset(v)
        return self._wrap(self._machine.transform(v, u=u, out_dtype=out_dtype))

    def inv_transform(self, h, u=None, out_dtype=None):
        """rbm.py:91-92 -> inv_transform_func (:51-54 / :64-67)."""
        h = _unwrap(h)
        if not self.built:
            raise ValueError("inv_transform needs a built RBM (the visible dimension is unknown)")
        if isinstance(h, Dataset):
            return self._machine.inv_transform_dataset(h)
        return self._wrap(self._machine.inv_transform(h, u=u, out_dtype=out_dtype))

    def cal_free_energy(self, v):
        """rbm.py:97-98 -> free_energy_func (:73-76)."""
        v = _unwrap(v)
        self._ensure_built(v)
        return self._wrap(self._machine.free_energy(v))

    def transform_proba(self, v):
        """sigmoid(v.W + c) itself (not part of the reference surface; used by parity tests)."""
        v = _unwrap(v)
        self._ensure_built(v)
        return self._machine.transform(v, want_p=True)[1]

    # ---- training ------------------------------------------------------------------------------
    def _hparams(self, update_mask=L.UPDATE_ALL, want_stats=False):
        hps = self.hps
        norm = hps.get("normalize", "sum")
        return Machine.hparams(lr=hps["lr"], k=hps.get("k", 1), persistent=hps.get("persistent", False),
                               momentum=hps.get("momentum", 0.0), weight_decay=hps.get("weight_decay", 0.0),
                               normalize=(norm in ("mean", True, 1)), update_mask=update_mask,
                               want_stats=want_stats)

    def _shard(self, V, batch):
        """Data-parallel layout: rank r keeps rows [r*b, (r+1)*b) of every global minibatch."""
        ctx = self._machine.ctx
        return shard_rows(V, batch, ctx.rank, ctx.world)

    def fit(self, V, verbose=1):
        """Train RBM with the data V (rbm.py:100-234); see _fit.  In Layer mode the trained parameters are then
        written back into the Keras variables."""
        self._fit(V, verbose)
        self._to_keras()
        return self

    def _fit(self, V, verbose=1):
        """Train RBM with the data V (rbm.py:100-234).

        V : 2d array (rows x input_dim), or an engine Dataset already resident on the GPU.
        Sequential minibatches of hps['batch_size'] rows, remainder last, hps['epochs'] passes, no
        shuffling (rbm.py:110-113,163,211,218).
        """
        V = _unwrap(V)
        self._ensure_built(V)
        hps = self.hps
        batch = int(hps["batch_size"])
        epochs = int(hps["epochs"])
        compat = hps.get("compat", "fused")
        m = self._machine
        if hps.get("persistent", False) and m.ctx is not None:
            self._ensure_chains(batch)
        if compat == "reference":
            if m.ctx.world > 1:
                # the reference's schedule walks the whole minibatch through three sequential single-parameter runs on
                # one device; sharding it would need its own row bookkeeping (every rank would otherwise add the same
                # gradient world times)
                raise ValueError("hps['compat'] = 'reference' runs on a single rank (world size is %d); "
                                 "use the default compat='fused' for data-parallel training" % m.ctx.world)
            dense = V.to_dense() if isinstance(V, PackedBits) else np.asarray(V, dtype=np.float32)
            return self._fit_reference(dense, batch, epochs, verbose)
        if compat != "fused":
            raise ValueError("hps['compat'] must be 'fused' or 'reference'")
        shuffle = bool(hps.get("shuffle", False))  # extension: the reference walks the rows in order (rbm.py:218)

        if isinstance(V, Dataset):
            ds, local_batch, row0, owns = V, batch, 0, False
        else:
            local, local_batch, row0 = self._shard(V, batch)
            raw = local.data if isinstance(local, PackedBits) else local
            on_host = not (L._is_torch(raw) and raw.is_cuda)
            if epochs == 1 and on_host and hps.get("stream", True) and not shuffle:
                # a single pass: stream the minibatches from host memory, copies overlapped with the chains
                if verbose == 1:
                    print(1, "/", epochs, " epochs", end="\r")
                st = m.fit_host(local, local_batch, self._hparams(), global_row0=row0, want_recon=bool(verbose))
                self._epochs_done += 1                 # numbers the shuffle permutations, saved in checkpoints
                st["epoch"] = 1
                st.setdefault("last_score", 0.0)       # same entry shape as the resident path below
                if verbose:
                    st["last_score"] = m.score(local[-(local.shape[0] % local_batch or local_batch):])
                    n_step = int(st["steps"])
                    print("\n{0:d}/{1:d}, score: {2:f}".format(n_step, n_step, st["last_score"]))
                self.history.append(st)
                m.ctx.sync()
                return self
            ds, owns = Dataset.from_array(m.ctx, local, m.compute), True
        n_rows = ds.shape[0]
        num_step = int(math.ceil(n_rows / local_batch)) if n_rows else 0
        hp = self._hparams()
        # hps['shuffle']: every epoch trains on the rows in a fresh keyed pseudo-random order (a device-side gather into
        # one reused buffer, so the captured step graph survives).  Under data parallelism each rank permutes its own
        # shard: global minibatch i is then made of every rank's i-th shuffled slice.
        order = None
        try:
            for k in range(epochs):
                if verbose == 1:
                    print(k + 1, "/", epochs, " epochs", end="\r")  # rbm.py:115
                want = bool(verbose)
                hp.want_stats = int(want)
                cur = ds
                if shuffle:
                    order = ds.shuffled(int(hps.get("shuffle_seed", self.seed)), self._epochs_done, into=order)
                    cur = order
                st = m.fit_epoch(cur, local_batch, hp, global_row0=row0, want_stats=True)
                self._epochs_done += 1
                st["epoch"] = k + 1
                self.history.append(st)
                if want:
                    # the reference prints this after every step (rbm.py:234); once per epoch here
                    print("\n{0:d}/{1:d}, score: {2:f}".format(num_step, num_step, st["last_score"]))
        finally:
            m.ctx.sync()
            if order is not None:
                order.close()
            if owns:
                ds.close()
        return self

    def _ensure_chains(self, batch):
        """Persistent chains for a GLOBAL minibatch of `batch` rows; this rank stores its batch / world of them.
        `_chains_set` counts the rows stored on THIS rank (what `save` writes and `load` restores)."""
        m = self._machine
        ctx = m.ctx
        rows = batch // ctx.world if ctx.world > 1 else batch
        if getattr(self, "_chains_set", 0) >= rows:
            return
        rng = np.random.default_rng(int(self.hps.get("chain_seed", 99)))
        full = (rng.random((batch, m.V)) < 0.5).astype(np.float32)
        m.set_chains(full[ctx.rank * rows:(ctx.rank + 1) * rows] if ctx.world > 1 else full)
        self._chains_set = rows

    def _fit_reference(self, V, batch, epochs, verbose):
        """The reference schedule, run for run (rbm.py:214-234): three single-parameter updates with
        fresh draws each, then the score from a fourth chain, printed per step."""
        m = self._machine
        n_rows = V.shape[0]
        num_step = n_rows // batch if n_rows % batch == 0 else n_rows // batch + 1  # rbm.py:110-111
        for k in range(epochs):
            if verbose == 1:
                print(k + 1, "/", epochs, " epochs", end="\r")
            for i in range(num_step):
                V_batch = V[i * batch:min((i + 1) * batch, n_rows)]       # rbm.py:211,218
                m.cd_step(V_batch, self._hparams(L.UPDATE_W))              # rbm_weight_update_func
                m.cd_step(V_batch, self._hparams(L.UPDATE_C))              # hidden_bias_update_func
                m.cd_step(V_batch, self._hparams(L.UPDATE_B))              # visible_bias_update_func
                score = m.score(V_batch)                                    # rbm.py:227-233
                self.history.append({"epoch": k + 1, "step": i + 1, "last_score": score})
                if verbose:
                    print("\n{0:d}/{1:d}, score: {2:f}".format(i + 1, num_step, score))  # rbm.py:234
        m.ctx.sync()
        return self
