"""Pin the oracle to the reference itself: golden vectors produced by executing the UNMODIFIED
/root/reference/ku/ebm/rbm.py (on tests/golden/kshim.py, a numpy stand-in for the keras backend) are
replayed through oracle/cd_oracle.py with the same recorded random draws.

Two sizes (64 -> 64 with minibatches of 16; 128 -> 128 with minibatches of 32, other seeds and learning rate).
The fixtures cover what the reference can execute at HEAD (see make_reference_fixtures.py): build,
transform_func / inv_transform_func / free_energy_func, and the first N/B - 1 minibatches of fit() -
runs A, B, C, the two free energies, run D and the printed score (rbm.py:214-234) - in both modes.
"""
import json
import os

import numpy as np
import pytest

from oracle import cd_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


@pytest.mark.parametrize("size", ["", "_128"])
@pytest.mark.parametrize("mode_name,mode", [("bernoulli", O.MODE_VISIBLE_BERNOULLI), ("gaussian", O.MODE_VISIBLE_GAUSSIAN)])
def test_inference_functions(mode_name, mode, size):
    g = _load("ref_rbm_%s%s.npz" % (mode_name, size))
    orc = O.OracleRBM(g["W0"], g["b0"], g["c0"], mode=mode, compute="f32")
    h, _ = orc.sample_h(g["infer_x"], g["infer_u_h"])                  # rbm.py:45-48 / :57-60
    assert np.array_equal(h, g["infer_h"])
    v, _ = orc.sample_v(g["infer_h_in"], g["infer_u_v"])               # rbm.py:51-54 / :63-67
    if mode == O.MODE_VISIBLE_BERNOULLI:
        assert np.array_equal(v, g["infer_v"])
    else:
        np.testing.assert_allclose(v, g["infer_v"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(orc.free_energy(g["infer_x"]), g["infer_fe"], rtol=1e-6)   # rbm.py:73-76


@pytest.mark.parametrize("size", ["", "_128"])
@pytest.mark.parametrize("mode_name,mode", [("bernoulli", O.MODE_VISIBLE_BERNOULLI), ("gaussian", O.MODE_VISIBLE_GAUSSIAN)])
def test_fit_schedule_replayed(mode_name, mode, size):
    g = _load("ref_rbm_%s%s.npz" % (mode_name, size))
    orc = O.OracleRBM(g["W0"], g["b0"], g["c0"], mode=mode, compute="f32")
    X, B, lr = g["X"], int(g["batch"]), float(g["lr"])
    steps = int(g["steps"])
    assert steps == X.shape[0] // B - 1      # the reference dies on the last minibatch (rbm.py:169)
    assert "int()" in str(g["fit_error"])
    lohi = list(O.batches(X.shape[0], B))
    for s in range(steps):
        lo, hi = lohi[s]
        v = X[lo:hi]
        draws = [(g["s%d_%s_uh" % (s, t)], g["s%d_%s_uv" % (s, t)]) for t in "ABCD"]
        (ah, av), (bh, bv), (ch, cv), (dh, dv) = draws
        orc.apply(orc.cd_stats(v, [ah], [None, av]), lr, 1)
        np.testing.assert_allclose(orc.W, g["s%d_W_after_A" % s], rtol=1e-6, atol=1e-7)
        orc.apply(orc.cd_stats(v, [bh], [None, bv]), lr, 2)
        np.testing.assert_allclose(orc.c, g["s%d_c_after_B" % s], rtol=1e-6, atol=1e-7)
        orc.apply(orc.cd_stats(v, [ch], [None, cv]), lr, 4)
        np.testing.assert_allclose(orc.b, g["s%d_b_after_C" % s], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(orc.free_energy(v), g["s%d_fe" % s], rtol=1e-6)
        h, _ = orc.sample_h(v, dh)
        v_neg, _ = orc.sample_v(h, dv)
        np.testing.assert_allclose(v_neg, g["s%d_v_neg_D" % s], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(orc.free_energy(v_neg), g["s%d_fe_p" % s], rtol=1e-6)
        assert abs(orc.score(v, dh, dv) - g["scores"][s]) <= 1e-5 * max(1.0, abs(g["scores"][s]))
    np.testing.assert_allclose(orc.W, g["W_final"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(orc.b, g["b_final"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(orc.c, g["c_final"], rtol=1e-6, atol=1e-7)


def test_reference_step_is_the_same_schedule():
    """OracleRBM.reference_step == the explicit A, B, C, D sequence used above."""
    g = _load("ref_rbm_bernoulli.npz")
    a = O.OracleRBM(g["W0"], g["b0"], g["c0"], compute="f32")
    v = g["X"][:int(g["batch"])]
    draws = [(g["s0_%s_uh" % t], g["s0_%s_uv" % t]) for t in "ABCD"]
    score = a.reference_step(v, draws, float(g["lr"]))
    np.testing.assert_allclose(a.W, g["s0_W_after_A"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(a.c, g["s0_c_after_B"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(a.b, g["s0_b_after_C"], rtol=1e-6, atol=1e-7)
    assert abs(score - g["scores"][0]) < 1e-5 * g["scores"][0]


def test_reference_config_omits_mode():
    """rbm.py:236-242 (defect D8): the reference's get_config has hps, output_dim, name and no mode."""
    cfg = json.loads(str(_load("ref_rbm_bernoulli.npz")["config"]))
    assert set(cfg) == {"hps", "output_dim", "name"}


def test_reference_dbn_defects_and_errors():
    """dbn.py at HEAD, executed with mock layers: which behaviours are kept (the ValueErrors) and which are
    defects the new DBN resolves (D7)."""
    with open(os.path.join(GOLD, "ref_dbn.json")) as f:
        ref = json.load(f)
    assert ref["fit_empty"] == "ValueError: Any rbm layer doesn't exist."
    assert ref["transform_empty"] == "ValueError: Any rbm layer doesn't exist."
    assert ref["second_add_stack"].startswith("AttributeError") and ref["fit"].startswith("AttributeError")
    assert ref["inv_transform_is_identity"] is True
    assert ref["constants"] == [0, 1, 2]
    dbn = O.OracleDBN()
    with pytest.raises(ValueError, match="Any rbm layer doesn't exist."):
        dbn.transform(np.zeros((2, 4)), [])


@pytest.mark.skipif(not os.path.exists("/root/reference/ku/ebm/rbm.py"), reason="the reference tree is only mounted in the build container")
def test_committed_fixtures_are_what_the_reference_produces(tmp_path, capsys):
    """Where /root/reference is mounted: re-run the generator (the unmodified ku/ebm/rbm.py on the numpy backend
    stand-in) and compare with the committed golden vectors, array by array."""
    import importlib.util
    import sys

    sys.path.insert(0, GOLD)
    spec = importlib.util.spec_from_file_location("make_reference_fixtures", os.path.join(GOLD, "make_reference_fixtures.py"))
    gen = importlib.util.module_from_spec(spec)
    saved = {k: v for k, v in sys.modules.items() if k == "ku" or k.startswith(("ku.", "tensorflow"))}
    try:
        spec.loader.exec_module(gen)
        for size in gen.SIZES:
            for mode in ("bernoulli", "gaussian"):
                gen.run_rbm(mode, out_dir=str(tmp_path), size=size)
                name = "ref_rbm_%s%s.npz" % (mode, size)
                new, old = np.load(tmp_path / name), _load(name)
                assert sorted(new.files) == sorted(old.files)
                for key in old.files:
                    assert np.array_equal(new[key], old[key]), (name, key)
        gen.run_dbn(out_dir=str(tmp_path))
        assert json.load(open(tmp_path / "ref_dbn.json")) == json.load(open(os.path.join(GOLD, "ref_dbn.json")))
    finally:
        for k in [k for k in sys.modules if k == "ku" or k.startswith(("ku.", "tensorflow"))]:
            del sys.modules[k]
        sys.modules.update(saved)
        capsys.readouterr()
