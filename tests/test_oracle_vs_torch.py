"""The oracle against a second, independent float32 library.

The reference's arithmetic lives in TensorFlow's CPU kernels (K.dot, K.sigmoid, K.less, K.cast, K.update_add), which
cannot be installed here: the fixtures of tests/golden pin the reference's OP SEQUENCE (its own rbm.py executed on a numpy
stand-in), not TensorFlow's rounding.  This file narrows that gap from the other side: the same op sequence written with
torch's CPU float32 kernels (oneDNN / MKL matmul, its own logistic) - another implementation of the same IEEE float32
operations TensorFlow would run - agrees with the oracle's float32 mode to float32 rounding, and with its float64 mode
(the yardstick of the GPU parity tests) to 1e-6.  So the tolerance the GPU tests use (1e-5) is not an artefact of which
float32 library evaluates the graph."""
import numpy as np
import pytest
import torch

from oracle import cd_oracle as O


def _torch_cd1(W, b, c, v, u_h, u_v, lr):
    """rbm.py:119-134 op for op (fused: one chain feeds the three updates), torch CPU float32."""
    W, b, c, v = (torch.from_numpy(np.array(x, np.float32)) for x in (W, b, c, v))
    u_h, u_v = torch.from_numpy(u_h), torch.from_numpy(u_v)
    p_h = torch.sigmoid(v @ W + c)                                    # :120  K.sigmoid(K.dot(v, W) + c)
    h_pos = (u_h < p_h).to(torch.float32)                             #       K.cast(K.less(u, p))
    p_v = torch.sigmoid(h_pos @ W.t() + b)                            # :121-122
    v_neg = (u_v < p_v).to(torch.float32)                             # :123
    h_neg = torch.sigmoid(v_neg @ W + c)                              # :124  probability, not a sample
    dW = v.t() @ h_pos - v_neg.t() @ h_neg                            # :125-126
    W2 = W + lr * dW                                                  # :127-128  K.update_add
    c2 = c + lr * (h_pos.sum(0) - h_neg.sum(0))                       # :129-131
    b2 = b + lr * (v.sum(0) - v_neg.sum(0))                           # :132-134
    fe = -(v @ b + torch.log(1 + torch.exp(v @ W + c)).sum(1))        # :73-75  free energy, the naive softplus
    return {k: t.numpy() for k, t in dict(p_h=p_h, h_pos=h_pos, p_v=p_v, v_neg=v_neg, h_neg=h_neg, dW=dW, W=W2, b=b2,
                                          c=c2, fe=fe).items()}


@pytest.mark.parametrize("V,H,B", [(784, 500, 128), (333, 130, 72)])
def test_oracle_float32_mode_equals_torch_float32(V, H, B):
    rng = np.random.default_rng(17)
    W, b, c = O.OracleRBM.init_params(V, H, seed=3)
    v = (rng.random((B, V)) < 0.1307).astype(np.float32)
    u_h, u_v = O.lattice_uniform(rng, (B, H)), O.lattice_uniform(rng, (B, V))
    ref64 = O.OracleRBM(W, b, c, compute="f64")
    (u_h,), (_, u_v), _ = O.condition_margin(ref64, v, [u_h], [None, u_v], k=1, margin=1e-4)   # no draw within 1e-4 of p
    lr = 1e-3
    t = _torch_cd1(W, b, c, v, u_h, u_v, lr)
    for mode, tol in (("f32", 2e-6), ("f64", 2e-6)):
        orc = O.OracleRBM(W, b, c, compute=mode)
        fe = orc.free_energy(v)
        st = orc.fused_step(v, [u_h], [None, u_v], lr)
        assert np.array_equal(st["h_pos"], t["h_pos"]) and np.array_equal(st["v_neg"], t["v_neg"])   # sampled states
        np.testing.assert_allclose(st["p_h_pos"], t["p_h"], rtol=tol, atol=1e-7)
        np.testing.assert_allclose(st["h_neg"], t["h_neg"], rtol=tol, atol=1e-7)
        np.testing.assert_allclose(st["dW"], t["dW"], rtol=1e-5, atol=2e-5)       # sums of up to B terms of size <= 1
        np.testing.assert_allclose(orc.W, t["W"], rtol=tol, atol=1e-7)
        np.testing.assert_allclose(orc.b, t["b"], rtol=tol, atol=1e-7)
        np.testing.assert_allclose(orc.c, t["c"], rtol=tol, atol=1e-7)
        np.testing.assert_allclose(fe, t["fe"], rtol=2e-6, atol=1e-4)


def test_float32_libraries_differ_by_rounding_only():
    """How far two float32 evaluations of the same pre-activation are apart (numpy/OpenBLAS vs torch/oneDNN): a few
    ulps of the contraction - the size of the effect the 1e-5 bar of the GPU tests has to absorb."""
    rng = np.random.default_rng(5)
    V, H, B = 4096, 256, 64
    W = rng.uniform(-0.05, 0.05, (V, H)).astype(np.float32)
    v = (rng.random((B, V)) < 0.5).astype(np.float32)
    x_np = v @ W
    x_t = (torch.from_numpy(v) @ torch.from_numpy(W)).numpy()
    x_64 = v.astype(np.float64) @ W.astype(np.float64)
    scale = np.abs(x_64).max()
    assert np.abs(x_np - x_64).max() < 2e-6 * scale * 4 and np.abs(x_t - x_64).max() < 2e-6 * scale * 4
    assert np.abs(x_np - x_t).max() < 1e-5 * scale
