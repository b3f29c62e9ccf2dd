"""A minimal stand-in for the part of TensorFlow that keras_unsupervised_b200.ebm.RBM touches when it subclasses
`tf.keras.layers.Layer` (TensorFlow is not installable here).  TEST INFRASTRUCTURE: it reproduces the
behaviours of tf.keras that matter for the wiring - `name`, `input_shape` and `output_shape` are read-only properties
(defect D6 of SURVEY.md 2.3), `input_shape` raises until the layer has been called, unknown constructor keywords are
refused, `__call__` builds on first use, weights are variables with `.numpy()` / `.assign()`.  tests/test_keras_layer.py
puts this directory on sys.path in a subprocess."""
import numpy as np

float32 = np.float32


class Variable:
    def __init__(self, value, name, trainable=True):
        self._v = np.array(value, dtype=np.float32)
        self.name = name + ":0"
        self.trainable = trainable
        self.shape = self._v.shape

    def numpy(self):
        return self._v.copy()

    def assign(self, value):
        value = np.asarray(value, dtype=np.float32)
        if value.shape != self._v.shape:
            raise ValueError("shape mismatch in assign")
        self._v = value.copy()
        return self


class Tensor:
    """An eager tensor: wraps an ndarray, has a shape and set_shape."""

    def __init__(self, value):
        self._v = np.asarray(value)
        self.shape = self._v.shape

    def numpy(self):
        return self._v

    def set_shape(self, shape):
        self.shape = tuple(shape)

    def __array__(self, dtype=None, copy=None):
        return self._v if dtype is None else self._v.astype(dtype)


def convert_to_tensor(x):
    return x if isinstance(x, Tensor) else Tensor(x)


def numpy_function(func, inp, Tout):
    return Tensor(np.asarray(func(*[np.asarray(i) for i in inp]), dtype=Tout))


class _Layer:
    _counter = {}

    def __init__(self, trainable=True, name=None, dtype=None, **kwargs):
        if kwargs:
            raise TypeError("Keyword argument not understood: %s" % sorted(kwargs))
        if name is None:
            base = type(self).__name__.lower()
            n = _Layer._counter.get(base, 0)
            _Layer._counter[base] = n + 1
            name = base if n == 0 else "%s_%d" % (base, n)
        self.__dict__["_name"] = name
        self.trainable = trainable
        self.built = False
        self._weights = []
        self._called_with = None

    name = property(lambda self: self._name)

    @property
    def input_shape(self):
        if self._called_with is None:
            raise AttributeError("The layer has never been called and thus has no defined input shape.")
        return self._called_with

    @property
    def output_shape(self):
        if self._called_with is None:
            raise AttributeError("The layer has never been called and thus has no defined output shape.")
        return self.compute_output_shape(self._called_with)

    def add_weight(self, name=None, shape=None, initializer=None, trainable=True, **kw):
        if initializer != "uniform":
            raise ValueError("only the 'uniform' initializer (RandomUniform(-0.05, 0.05)) is modelled")
        rng = np.random.default_rng(len(self._weights) + 1)
        v = Variable(rng.uniform(-0.05, 0.05, shape), name, trainable)
        self._weights.append(v)
        return v

    @property
    def weights(self):
        return list(self._weights)

    def get_weights(self):
        return [w.numpy() for w in self._weights]

    def set_weights(self, values):
        for w, v in zip(self._weights, values):
            w.assign(v)

    def build(self, input_shape):
        self.built = True

    def compute_output_shape(self, input_shape):
        return input_shape

    def __call__(self, x):
        x = convert_to_tensor(x)
        if not self.built:
            self.build((None,) + tuple(x.shape[1:]))
            self.built = True
        self._called_with = (None,) + tuple(x.shape[1:])
        return self.call(x)


class _Layers:
    Layer = _Layer


class _Keras:
    layers = _Layers


keras = _Keras
