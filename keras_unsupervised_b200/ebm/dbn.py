"""ku.ebm.DBN, same surface: a stack of RBMs trained greedily, layer by layer.

Mirrors /root/reference/ku/ebm/dbn.py (`add_stack` :14, `fit` :34, `transform` :57, `inv_transform`
:77) with the defects of SURVEY.md 2.3 (D7) resolved as the code intends: the loop variable is used
instead of the undefined `self.rbm_layer`, the dimension check looks at `_rbm_layers`, and
`inv_transform` really walks the stack in reverse.  Between layers the data never leaves the GPU: the
sampled hidden states of layer l (dbn.py:55) are produced as an engine Dataset and consumed as one.
"""
from __future__ import annotations

import json
import os

import numpy as np

from ..engine import Dataset

# Constants (dbn.py:6-8)
MODE_VISIBLE_BERNOULLI = 0
MODE_VISIBLE_GAUSSIAN = 1
MODE_COMPLEX = 2  # TODO


def _visible_dim(layer):
    """A layer's input dimension if it is known yet (a Keras Layer's `input_shape` raises until the layer is called)."""
    n = getattr(layer, "_n_visible", None)
    if n is not None:
        return int(n)
    shape = getattr(layer, "input_shape", None)
    return int(shape[1]) if shape else None


class DBN(object):
    """Deep belief network."""

    def add_stack(self, rbm_layer):
        """Add a rbm layer to the dbn stack (dbn.py:14-32)."""
        if hasattr(self, "_rbm_layers"):
            prev = self._rbm_layers[-1]
            prev_out = prev.output_dim
            next_in = _visible_dim(rbm_layer)
            if next_in is not None and int(prev_out) != int(next_in):
                raise ValueError("A previous RBM layer's output dimension must"
                                 + "be equal to a next one's input dimension.")  # dbn.py:29-30
            self._rbm_layers.append(rbm_layer)
        else:
            self._rbm_layers = [rbm_layer]
        if hasattr(rbm_layer, "_set_stack_index"):
            rbm_layer._set_stack_index(len(self._rbm_layers) - 1)

    def _check(self):
        if hasattr(self, "_rbm_layers") != True:  # noqa: E712  (dbn.py:47-48)
            raise ValueError("Any rbm layer doesn't exist.")

    def fit(self, V, verbose=1):
        """Train DBN with the data V (dbn.py:34-55): for each layer, fit, then feed its sampled hidden
        states of the whole data set to the next layer."""
        self._check()
        cur = V[0] if isinstance(V, (list, tuple)) and len(V) == 1 else V
        cur_owned = False
        for i, rbm_layer in enumerate(self._rbm_layers):
            print("Train {0:s}.".format(str(rbm_layer.name)))  # dbn.py:53
            if not rbm_layer.built:
                rbm_layer.build((None, int(cur.shape[1])))
            m = rbm_layer._machine
            if not isinstance(cur, Dataset) and m.ctx.world == 1:
                cur, cur_owned = Dataset.from_array(m.ctx, cur, m.compute), True
            rbm_layer.fit(cur, verbose=verbose)                      # dbn.py:54
            nxt = None
            if i + 1 < len(self._rbm_layers):
                nxt = rbm_layer.transform(cur)                       # dbn.py:55: Dataset in -> Dataset out
                nxt = nxt[0] if isinstance(nxt, list) else nxt
            if cur_owned:
                cur.close()
            cur, cur_owned = nxt, isinstance(nxt, Dataset)
        return self

    def transform(self, V):
        """Transform the visible unit (dbn.py:57-75)."""
        self._check()
        V_p = V[0] if isinstance(V, (list, tuple)) and len(V) == 1 else V
        V_p = V_p.copy() if isinstance(V_p, np.ndarray) else V_p
        for rbm_layer in self._rbm_layers:
            out = rbm_layer.transform(V_p)
            V_p = out[0] if isinstance(out, list) else out
        return V_p

    def inv_transform(self, H):
        """Transform the hidden unit (dbn.py:77-96), top layer first.  After fine_tune the layers below the top one
        go down through their own generative weights."""
        self._check()
        H_p = H[0] if isinstance(H, (list, tuple)) and len(H) == 1 else H
        H_p = H_p.copy() if isinstance(H_p, np.ndarray) else H_p
        gen = getattr(self, "_gen", None)
        for i in reversed(range(len(self._rbm_layers))):
            rbm_layer = self._rbm_layers[i]
            if gen is not None and i < len(gen):
                out = gen[i].inv_transform_dataset(H_p) if isinstance(H_p, Dataset) else gen[i].inv_transform(H_p)
            else:
                out = rbm_layer.inv_transform(H_p)
            H_p = out[0] if isinstance(out, list) else out
        return H_p

    # ---- fine-tuning after the greedy pass (the reference stops at dbn.py:34-55) ----
    def untie(self):
        """Give every layer below the top one its own generative parameters: a second engine model initialised with the
        layer's W, b, c.  The layer itself keeps the recognition direction (v.W + c), the copy the generative one
        (h.W^T + b); the top RBM stays undirected.  Idempotent."""
        self._check()
        gen = getattr(self, "_gen", None)
        if gen is None or len(gen) != len(self._rbm_layers) - 1:
            from ..engine import Machine

            gen = []
            for layer in self._rbm_layers[:-1]:
                if not layer.built:
                    raise ValueError("untie needs a trained (built) stack")
                m = layer._machine
                g = Machine(m.ctx, m.V, m.H, m.mode, m.compute, seed=layer.seed + 500009)
                g.set_params(*m.get_params())
                gen.append(g)
            self._gen = gen
        return self._gen

    def fine_tune(self, V, epochs=1, batch_size=None, lr=None, k=1, normalize=None, verbose=0):
        """Up-down (contrastive wake-sleep) fine-tuning of a greedily trained stack (Hinton, Osindero & Teh 2006,
        appendix B; the CPU restatement the tests compare with is OracleDBN.up_down_step).  Per minibatch:
          wake    s_l = sampled hidden states of the recognition layers, bottom-up (RBM.transform)
          top     CD-k of the top RBM on s_{L-1}; t_{L-1} = the visible state its chain ends in
          sleep   t_{l-1} = sampled through the generative weights, top-down
          update  generative weights predict s_{l-1} from s_l, recognition weights predict t_l from t_{l-1}
                  (kucd_rbm_delta_rule: the CD path's projection, dW contraction and update kernels)
        Bernoulli layers only.  V: a host array (rows, input_dim).  Returns self."""
        from ..engine import Machine

        self._check()
        layers = self._rbm_layers
        if len(layers) < 2:
            raise ValueError("fine_tune needs at least two stacked layers")
        if any(not l.built for l in layers):
            raise ValueError("fine_tune needs a trained (built) stack: call fit first")
        if any(int(l.mode) != MODE_VISIBLE_BERNOULLI for l in layers):
            raise ValueError("fine_tune is implemented for Bernoulli (sigmoid) layers")
        V = V[0] if isinstance(V, (list, tuple)) and len(V) == 1 else V
        V = V.numpy() if isinstance(V, Dataset) else np.asarray(V)
        top = layers[-1]
        batch = int(batch_size or top.hps["batch_size"])
        lr = float(lr if lr is not None else top.hps["lr"])
        if normalize is None:
            normalize = top.hps.get("normalize", "sum") in ("mean", True, 1)
        gen = self.untie()
        hp = Machine.hparams(lr=lr, k=int(k), normalize=bool(normalize))
        n = V.shape[0]
        for epoch in range(int(epochs)):
            for lo in range(0, n, batch):                          # sequential slices, remainder last (rbm.py:211,218)
                s = [V[lo:lo + batch]]
                rows = s[0].shape[0]
                for layer in layers[:-1]:
                    s.append(layer._machine.transform(s[-1]))
                top._machine.cd_step(s[-1], hp)
                t = [None] * len(layers)
                t[-1] = top._machine.last_stats(rows, states=True, grads=False)["v_neg"]
                for i in range(len(layers) - 2, -1, -1):
                    t[i] = gen[i].inv_transform(t[i + 1])
                for i in range(len(layers) - 1):
                    gen[i].delta_rule(False, s[i + 1], s[i], lr, normalize)
                    layers[i]._machine.delta_rule(True, t[i], t[i + 1], lr, normalize)
            if verbose:
                print("fine-tune epoch {0:d} done.".format(epoch + 1))
        for layer in layers:
            if hasattr(layer, "_to_keras"):
                layer._to_keras()
        return self

    def generate(self, n_samples, gibbs_steps=100, seed=0, h_init=None):
        """Draw visible samples from the trained stack (what the reference's top-down pass, dbn.py:77-96, is for; SURVEY.md
        8f rank 4): alternating Gibbs sampling in the top RBM for `gibbs_steps` sweeps (h -> v -> h, the associative
        memory of a DBN), starting from `h_init` or Bernoulli(0.5) hidden states, then one top-down pass through
        the lower layers (inv_transform, top layer first).  Everything between the first upload and the final read-back
        stays on the GPU as engine Datasets."""
        self._check()
        top = self._rbm_layers[-1]
        if not top.built:
            raise ValueError("generate needs a trained (built) stack")
        m = top._machine
        if h_init is None:
            rng = np.random.default_rng(seed)
            h_init = (rng.random((int(n_samples), int(top.output_dim))) < 0.5).astype(np.float32)
        h = Dataset.from_array(m.ctx, np.asarray(h_init, dtype=np.float32), m.compute)
        for _ in range(int(gibbs_steps)):
            v = m.inv_transform_dataset(h)
            h.close()
            h = m.transform_dataset(v)
            v.close()
        cur = h
        gen = getattr(self, "_gen", None)
        for i in reversed(range(len(self._rbm_layers))):
            down = gen[i] if gen is not None and i < len(gen) else self._rbm_layers[i]._machine
            nxt = down.inv_transform_dataset(cur)
            cur.close()
            cur = nxt
        out = cur.numpy()
        cur.close()
        return out

    # ---- checkpoint: one JSON file with the stack's configs + one .npz per layer (cf. ku/utility.py:7-33, which
    # stores a JSON architecture next to an HDF5 weight file) ----
    def save(self, directory):
        self._check()
        os.makedirs(directory, exist_ok=True)
        configs = []
        for i, layer in enumerate(self._rbm_layers):
            cfg = layer.get_config()
            cfg["input_dim"] = _visible_dim(layer)
            configs.append(cfg)
            if layer.built:
                layer.save(os.path.join(directory, "layer%d.npz" % i))
        # the untied generative parameters fine_tune trains (one engine model per layer below the top one): without them
        # a reloaded stack would generate through the recognition weights again
        gen = getattr(self, "_gen", None) or []
        for i, g in enumerate(gen):
            W, b, c = g.get_params()
            np.savez(os.path.join(directory, "gen%d.npz" % i), rbm_weight=W, rbm_visible_bias=b, rbm_hidden_bias=c)
        with open(os.path.join(directory, "dbn.json"), "w") as f:
            json.dump({"layers": configs, "untied": len(gen)}, f, indent=1)

    @classmethod
    def load(cls, directory, layer_class=None, **kwargs):
        if layer_class is None:
            from .rbm import RBM as layer_class
        with open(os.path.join(directory, "dbn.json")) as f:
            meta = json.load(f)
        configs = meta["layers"]
        dbn = cls()
        for i, cfg in enumerate(configs):
            input_dim = cfg.pop("input_dim", None)
            layer = layer_class.from_config(cfg, **kwargs)
            dbn.add_stack(layer)
            path = os.path.join(directory, "layer%d.npz" % i)
            if os.path.exists(path):
                if input_dim is not None and not layer.built:
                    layer.build((None, input_dim))
                layer.load(path)
        if meta.get("untied", 0):
            gen = dbn.untie()
            for i, g in enumerate(gen):
                z = np.load(os.path.join(directory, "gen%d.npz" % i))
                g.set_params(z["rbm_weight"], z["rbm_visible_bias"], z["rbm_hidden_bias"])
        return dbn
