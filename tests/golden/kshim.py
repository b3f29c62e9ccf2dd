"""A numpy stand-in for the few Keras-backend calls /root/reference/ku/ebm/rbm.py makes, so that the
UNMODIFIED reference file can be imported and executed in a container without TensorFlow.

Only used by make_reference_fixtures.py (in the build container, where /root/reference exists) to
freeze golden input/output vectors of the reference's own op sequence.  It is deliberately dumb: a lazy
expression graph evaluated with float32 numpy, one evaluation per K.function call, with every random
node drawing afresh per call (as a TF graph execution does) from a recorded stream.

Semantics implemented (tensorflow.python.keras.backend of TF 2.3, the release matching setup.py:70-74):
  placeholder, variable, dot, transpose, sigmoid, relu, exp, log, sum, squeeze, expand_dims, less, cast,
  random_uniform (float32 in [0,1)), update_add, function, floatx; Layer.add_weight / build / get_config;
  initializers.get('uniform') = RandomUniform(-0.05, 0.05); and, for the Gaussian-visible mode,
  ku.backend_ext.tensorflow_backend.multivariate_normal_diag(loc, scale_diag).sample().
One leniency: K.cast(x) without dtype (rbm.py:46,52,58,82,85) is taken as float32 - the dtype rbm.py:123
spells out - where real Keras would raise TypeError (SURVEY.md defect D2).
"""
from __future__ import annotations

import sys
import types

import numpy as np

F32 = np.float32


class Recorder:
    """Source of all randomness + log of every K.function call."""

    def __init__(self, seed=7, param_seed=0):
        self.rng = np.random.default_rng(seed)
        self.param_rng = np.random.default_rng(param_seed)
        self.calls = []      # dicts: fn id, inputs, outputs, draws [(node id, kind, array)]
        self._cur = None
        self.n_nodes = 0

    def uniform(self, shape):
        return (self.rng.integers(0, 1 << 23, size=shape, dtype=np.uint32).astype(F32) * F32(2.0 ** -23)).astype(F32)

    def normal(self, shape):
        return self.rng.standard_normal(size=shape).astype(F32)


REC = Recorder()


class Node:
    def __init__(self, op, inputs=(), **attrs):
        self.op, self.inputs, self.attrs = op, tuple(inputs), attrs
        self.id = REC.n_nodes
        REC.n_nodes += 1

    # arithmetic used by rbm.py: a + b, a - b, scalar * a, -1 * (...), 1 + a
    def __add__(self, o): return Node("add", (self, wrap(o)))
    def __radd__(self, o): return Node("add", (wrap(o), self))
    def __sub__(self, o): return Node("sub", (self, wrap(o)))
    def __rsub__(self, o): return Node("sub", (wrap(o), self))
    def __mul__(self, o): return Node("mul", (self, wrap(o)))
    def __rmul__(self, o): return Node("mul", (wrap(o), self))
    def __neg__(self): return Node("mul", (wrap(-1.0), self))

    @property
    def shape(self):
        return self.attrs.get("shape")


class Variable(Node):
    def __init__(self, value, name=None):
        super().__init__("variable", (), name=name)
        self.value = np.array(value, dtype=F32)

    @property
    def shape(self):
        return self.value.shape


def wrap(x):
    return x if isinstance(x, Node) else Node("const", (), value=np.asarray(x, dtype=F32))


def evaluate(node, feed, memo):
    if node.id in memo:
        return memo[node.id]
    op = node.op
    if op == "const":
        r = node.attrs["value"]
    elif op == "variable":
        r = node.value
    elif op == "placeholder":
        r = feed[node.id]
    elif op == "random_uniform":
        r = REC.uniform(node.attrs["shape"])
        REC._cur["draws"].append((node.id, "uniform", r))
    elif op == "normal_sample":
        loc = evaluate(node.inputs[0], feed, memo)
        scale = node.attrs["scale"]
        n = REC.normal(loc.shape)
        REC._cur["draws"].append((node.id, "normal", n))
        r = (loc + F32(1.0) * (np.asarray(scale, F32) * n)).astype(F32)
    else:
        a = [evaluate(i, feed, memo) for i in node.inputs]
        if op == "add": r = (a[0] + a[1]).astype(F32)
        elif op == "sub": r = (a[0] - a[1]).astype(F32)
        elif op == "mul": r = (a[0] * a[1]).astype(F32)
        elif op == "dot": r = np.dot(a[0], a[1]).astype(F32)
        elif op == "transpose": r = a[0].T
        elif op == "sigmoid": r = (F32(1) / (F32(1) + np.exp(-a[0]))).astype(F32)
        elif op == "relu": r = np.maximum(a[0], F32(0))
        elif op == "exp": r = np.exp(a[0]).astype(F32)
        elif op == "log": r = np.log(a[0]).astype(F32)
        elif op == "sum": r = a[0].sum(axis=node.attrs["axis"]).astype(F32)
        elif op == "squeeze": r = np.squeeze(a[0], axis=node.attrs["axis"])
        elif op == "expand_dims": r = np.expand_dims(a[0], axis=node.attrs["axis"])
        elif op == "less": r = a[0] < a[1]
        elif op == "cast": r = a[0].astype(node.attrs["dtype"])
        elif op == "update_add": r = (node.inputs[0].value + a[1]).astype(F32)
        else:
            raise NotImplementedError(op)
    memo[node.id] = r
    return r


class Function:
    _count = 0

    def __init__(self, inputs, outputs):
        self.inputs, self.outputs = list(inputs), list(outputs)
        self.id = Function._count
        Function._count += 1

    def __call__(self, values):
        if not isinstance(values, (list, tuple)):
            values = [values]
        feed = {p.id: np.asarray(v, dtype=F32) for p, v in zip(self.inputs, values)}
        REC._cur = {"fn": self.id, "inputs": [np.array(v, F32) for v in values], "draws": []}
        memo = {}
        outs = [np.array(evaluate(o, feed, memo)) for o in self.outputs]
        for o in self.outputs:  # apply assignments after everything was evaluated
            if o.op == "update_add":
                o.inputs[0].value = memo[o.id]
        REC._cur["outputs"] = outs
        REC.calls.append(REC._cur)
        REC._cur = None
        return outs


def _backend():
    K = types.ModuleType("tensorflow.python.keras.backend")
    K.floatx = lambda: "float32"
    K.placeholder = lambda shape=None, name=None, **kw: Node("placeholder", (), shape=shape, name=name)
    K.variable = lambda value, dtype=None, name=None: Variable(value, name)
    K.dot = lambda a, b: Node("dot", (wrap(a), wrap(b)))
    K.transpose = lambda a: Node("transpose", (wrap(a),))
    K.sigmoid = lambda a: Node("sigmoid", (wrap(a),))
    K.relu = lambda a: Node("relu", (wrap(a),))
    K.exp = lambda a: Node("exp", (wrap(a),))
    K.log = lambda a: Node("log", (wrap(a),))
    K.sum = lambda a, axis=None, **kw: Node("sum", (wrap(a),), axis=axis)
    K.squeeze = lambda a, axis: Node("squeeze", (wrap(a),), axis=axis)
    K.expand_dims = lambda a, axis=-1: Node("expand_dims", (wrap(a),), axis=axis)
    K.less = lambda a, b: Node("less", (wrap(a), wrap(b)))
    K.cast = lambda a, dtype="float32": Node("cast", (wrap(a),), dtype=np.dtype(dtype))
    K.random_uniform = lambda shape, **kw: Node("random_uniform", (), shape=tuple(int(s) for s in shape))
    K.update_add = lambda var, delta: Node("update_add", (var, wrap(delta)))
    K.function = lambda inputs, outputs, **kw: Function(inputs, outputs)
    return K


class Layer(object):
    def __init__(self, **kwargs):
        self.built = False

    def add_weight(self, name=None, shape=None, initializer=None, trainable=True, **kw):
        return Variable(_initializers().get(initializer)(shape), name)

    def build(self, input_shape):
        self.built = True

    def get_config(self):
        return {}


def _initializers():
    m = types.ModuleType("tensorflow.python.keras.initializers")

    def get(name):
        assert name == "uniform"
        return lambda shape: REC.param_rng.uniform(-0.05, 0.05, size=shape).astype(F32)

    m.get = get
    return m


class _MVN:
    def __init__(self, loc, scale_diag):
        self.loc, self.scale = loc, scale_diag

    def sample(self):
        return Node("normal_sample", (wrap(self.loc),), scale=self.scale)


def install(reference_root="/root/reference"):
    """Register the stand-in modules and expose the reference's ku.ebm without running ku/__init__.py
    (which imports cupy, cv2, ... for unrelated sub-packages)."""
    import os

    def pkg(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
        return m

    tf = pkg("tensorflow")
    tfp = pkg("tensorflow.python")
    tfk = pkg("tensorflow.python.keras")
    K = _backend()
    sys.modules["tensorflow.python.keras.backend"] = K
    layers = types.ModuleType("tensorflow.python.keras.layers")
    layers.Layer = Layer
    sys.modules["tensorflow.python.keras.layers"] = layers
    init = _initializers()
    sys.modules["tensorflow.python.keras.initializers"] = init
    tf.python, tfp.keras = tfp, tfk
    tfk.backend, tfk.layers, tfk.initializers = K, layers, init

    pkg("ku", os.path.join(reference_root, "ku"))
    pkg("ku.ebm", os.path.join(reference_root, "ku", "ebm"))
    be = pkg("ku.backend_ext")
    Ke = types.ModuleType("ku.backend_ext.tensorflow_backend")
    Ke.multivariate_normal_diag = lambda loc, scale_diag: _MVN(loc, scale_diag)
    sys.modules["ku.backend_ext.tensorflow_backend"] = Ke
    be.tensorflow_backend = Ke
    return K
