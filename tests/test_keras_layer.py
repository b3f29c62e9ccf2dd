"""`RBM` as a Keras Layer (rbm.py:19).  TensorFlow cannot be installed here, so the class wiring is
exercised against tests/fake_tf (a stand-in with tf.keras' read-only `name` / `input_shape` properties, strict
constructor keywords, build-on-first-call and assignable variables), in a subprocess so that the fake module never
leaks into the other tests.  The CPU part needs no engine; the GPU part trains through the Layer surface."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAKE = os.path.join(ROOT, "tests", "fake_tf")


def _run(code):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([FAKE, ROOT, os.environ.get("PYTHONPATH", "")]))
    res = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


def test_plain_class_without_tensorflow():
    from keras_unsupervised_b200.ebm import rbm

    assert rbm.IS_KERAS_LAYER is False and rbm.RBM.__mro__[1] is object      # TensorFlow is absent in this image


def test_rbm_is_a_keras_layer_when_tensorflow_imports():
    out = _run("""
        import tensorflow as tf
        from keras_unsupervised_b200.ebm import RBM, DBN, MODE_VISIBLE_BERNOULLI, rbm as R
        assert R.IS_KERAS_LAYER and issubclass(RBM, tf.keras.layers.Layer)
        hps = {'batch_size': 128, 'epochs': 1, 'lr': 1e-3}
        a = RBM(hps, 500, name='rbm', mode=MODE_VISIBLE_BERNOULLI)            # rbm.py:22; D6: name via the base class
        assert a.name == 'rbm' and not a.built and a.output_dim == 500
        assert a.get_config() == {'hps': hps, 'output_dim': 500, 'name': 'rbm', 'mode': MODE_VISIBLE_BERNOULLI}
        assert RBM(hps, 10).name == 'rbm'                                      # Keras-generated default name
        try:
            RBM(hps, 10, bogus=1)
            raise SystemExit('unknown keyword accepted')
        except TypeError:
            pass
        assert a.compute_output_shape((None, 784)) == (None, 500)
        twin = RBM.from_config(a.get_config())
        assert twin.name == 'rbm' and twin.mode == MODE_VISIBLE_BERNOULLI
        d = DBN(); d.add_stack(a); d.add_stack(RBM(hps, 64, name='top'))       # unbuilt Layers: dimensions unknown yet
        assert len(d._rbm_layers) == 2
        print('ok')
    """)
    assert out.strip().endswith("ok")


@pytest.mark.gpu
def test_layer_surface_trains_and_mirrors_the_keras_variables():
    out = _run("""
        import numpy as np, tensorflow as tf
        from keras_unsupervised_b200.ebm import RBM, MODE_VISIBLE_BERNOULLI
        rng = np.random.default_rng(0)
        X = (rng.random((512, 200)) < 0.2).astype(np.float32)
        rbm = RBM({'batch_size': 128, 'epochs': 2, 'lr': 1e-3, 'seed': 7}, 64, name='rbm', mode=MODE_VISIBLE_BERNOULLI)
        y = rbm(X)                                                             # Layer.__call__: build, then call
        assert rbm.built and tuple(y.shape) == (None, 64) and rbm.input_shape == (None, 200)
        h = np.asarray(y)
        assert h.shape == (512, 64) and set(np.unique(h)) <= {0.0, 1.0}
        names = [w.name for w in rbm.weights]
        assert names == ['rbm_weight:0', 'rbm_hidden_bias:0', 'rbm_visible_bias:0']     # rbm.py:30,34,40
        W0 = rbm.weights[0].numpy()
        assert np.array_equal(W0, rbm.rbm_weight)                              # the engine starts from the Keras init
        rbm.fit(X, verbose=0)
        W1 = rbm.rbm_weight
        assert not np.array_equal(W0, W1) and np.array_equal(rbm.weights[0].numpy(), W1)   # written back after fit
        assert np.array_equal(rbm.weights[1].numpy(), rbm.hidden_bias)
        # Model.load_weights writes the variables directly: sync_from_keras pushes them to the engine
        rbm.weights[0].assign(W0)
        rbm.sync_from_keras()
        assert np.array_equal(rbm.rbm_weight, W0)
        rbm.set_weights([W1, rbm.hidden_bias, rbm.visible_bias])
        assert np.array_equal(rbm.weights[0].numpy(), W1) and np.array_equal(rbm.get_weights()[0], W1)
        print('ok')
    """)
    assert out.strip().endswith("ok")
