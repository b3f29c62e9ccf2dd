"""Host logic of RBM.fit without a GPU: which engine entry points a fit drives, in which order and with which
arguments (minibatch slicing of rbm.py:110-111,163,211,218; the three single-parameter runs of rbm.py:214-216; the
extensions `shuffle`, `stream`, `persistent`, packed input).  The engine layer (Machine / Dataset) is replaced by
recording fakes; nothing here computes."""
import numpy as np
import pytest

from keras_unsupervised_b200 import _lib as L
from keras_unsupervised_b200.data import PackedBits
from keras_unsupervised_b200.ebm import rbm as R


class FakeCtx:
    rank, world = 0, 1

    def sync(self):
        pass


class FakeDataset:
    made = []

    def __init__(self, ctx, rows, dim, tag):
        self.ctx, self._shape, self.tag, self.closed = ctx, (rows, dim), tag, False
        self.shuffles = []

    @classmethod
    def from_array(cls, ctx, data, compute):
        ds = cls(ctx, data.shape[0], data.shape[1], "uploaded")
        ds.source = data
        cls.made.append(ds)
        return ds

    @property
    def shape(self):
        return self._shape

    def shuffled(self, seed, epoch, into=None):
        self.shuffles.append((seed, epoch, into))
        if into is not None:
            return into
        out = FakeDataset(self.ctx, *self._shape, tag="order")
        FakeDataset.made.append(out)
        return out

    def close(self):
        self.closed = True


class FakeMachine:
    def __init__(self, ctx, V, H, mode, compute, seed=None):
        self.ctx, self.V, self.H, self.mode, self.compute, self.seed = ctx, V, H, mode, compute, seed
        self.calls = []
        self._steps = 0

    hparams = staticmethod(R.Machine.hparams)

    def set_params(self, W=None, b=None, c=None):
        self.calls.append(("set_params",))

    def get_params(self):
        return (np.zeros((self.V, self.H), np.float32), np.zeros(self.V, np.float32), np.zeros(self.H, np.float32))

    def set_seed(self, seed, step_count=0):
        self.calls.append(("set_seed", seed, step_count))

    def counters(self):
        return {"seed": self.seed, "step_count": self._steps, "n_chains": 0}

    def set_chains(self, v):
        self.calls.append(("set_chains", v.shape, float(np.mean(v))))

    def fit_host(self, V, batch, hp, global_row0=0, want_recon=True):
        self.calls.append(("fit_host", V, batch, hp, global_row0, want_recon))
        steps = -(-V.shape[0] // batch)
        return {"steps": steps, "rows": V.shape[0], "device_ms": 0.0, "last_score": 0.0, "last_recon_err": 0.0,
                "step_recon_err": np.zeros(steps, np.float32)}

    def fit_epoch(self, ds, batch, hp, global_row0=0, want_stats=True):
        self.calls.append(("fit_epoch", ds, batch, hp.want_stats, global_row0))
        return {"steps": -(-ds.shape[0] // batch), "rows": ds.shape[0], "device_ms": 0.0, "last_score": 0.5,
                "last_recon_err": 0.1}

    def cd_step(self, v, hp, **kw):
        self.calls.append(("cd_step", v.shape[0], hp.update_mask, hp.k))

    def score(self, v, u_h=None, u_v=None):
        self.calls.append(("score", v.shape[0]))
        return 0.25


@pytest.fixture
def fakes(monkeypatch):
    FakeDataset.made = []
    monkeypatch.setattr(R, "Machine", FakeMachine)
    monkeypatch.setattr(R, "Dataset", FakeDataset)
    return FakeCtx()


def _rbm(ctx, **hps):
    base = {"batch_size": 128, "epochs": 1, "lr": 1e-3}
    base.update(hps)
    r = R.RBM(base, 64, name="r", mode=R.MODE_VISIBLE_BERNOULLI, context=ctx)
    return r


def _names(m):
    return [c[0] for c in m.calls]


def test_one_epoch_of_a_host_array_is_streamed(fakes):
    X = np.zeros((1000, 200), np.float32)
    r = _rbm(fakes, k=3, momentum=0.5, weight_decay=1e-4, normalize="mean")
    r.fit(X, verbose=0)
    m = r._machine
    assert (m.V, m.H, m.compute) == (200, 64, L.COMPUTE_F32X3)                # built from the data, float32 by default
    assert _names(m) == ["set_params", "fit_host"] and not FakeDataset.made
    _, V, batch, hp, row0, want_recon = m.calls[1]
    assert V is X and batch == 128 and row0 == 0 and not want_recon
    assert (hp.k, hp.persistent, hp.normalize, hp.update_mask) == (3, 0, 1, L.UPDATE_ALL)
    assert hp.lr == pytest.approx(1e-3) and hp.momentum == pytest.approx(0.5) and hp.weight_decay == pytest.approx(1e-4)
    assert r.history[-1]["steps"] == 8                                         # ceil(1000 / 128), rbm.py:110-111


def test_several_epochs_train_a_resident_data_set(fakes, capsys):
    X = np.zeros((300, 50), np.float32)
    r = _rbm(fakes, epochs=3, dtype="bf16")
    r.fit(X, verbose=1)
    m = r._machine
    assert m.compute == L.COMPUTE_BF16
    assert _names(m) == ["set_params", "fit_epoch", "fit_epoch", "fit_epoch"]
    (ds,) = FakeDataset.made
    assert ds.source is X and ds.closed and all(c[1] is ds and c[2] == 128 for c in m.calls[1:])
    assert [h["epoch"] for h in r.history] == [1, 2, 3]
    out = capsys.readouterr().out
    assert out.count("3/3, score: 0.500000") == 3 and "1 / 3  epochs" in out     # rbm.py:115,234 print formats
    # stream=False sends a single epoch the same way
    FakeDataset.made = []
    r2 = _rbm(fakes, stream=False)
    r2.fit(X, verbose=0)
    assert _names(r2._machine) == ["set_params", "fit_epoch"] and FakeDataset.made[0].closed


def test_shuffle_numbers_its_permutations_across_fits(fakes):
    X = np.zeros((300, 50), np.float32)
    r = _rbm(fakes, epochs=2, shuffle=True, seed=11)
    r.fit(X, verbose=0)
    r.fit(X, verbose=0)
    first, order1, second, order2 = FakeDataset.made
    assert [(s, e) for s, e, _ in first.shuffles] == [(11, 0), (11, 1)]
    assert [(s, e) for s, e, _ in second.shuffles] == [(11, 2), (11, 3)]        # the count goes on: no order is reused
    assert first.shuffles[0][2] is None and first.shuffles[1][2] is order1      # one buffer per fit, overwritten
    fits = [c for c in r._machine.calls if c[0] == "fit_epoch"]
    assert [c[1] for c in fits] == [order1, order1, order2, order2]             # training reads the shuffled copy
    assert all(d.closed for d in FakeDataset.made)
    r3 = _rbm(fakes, epochs=1, shuffle=True, shuffle_seed=5)
    r3.fit(X, verbose=0)                                                         # one epoch + shuffle: resident, not streamed
    assert "fit_host" not in _names(r3._machine) and FakeDataset.made[-2].shuffles[0][:2] == (5, 0)


def test_reference_schedule_runs_three_single_parameter_steps_per_minibatch(fakes, capsys):
    X = np.zeros((300, 50), np.float32)
    r = _rbm(fakes, compat="reference", epochs=2)
    r.fit(X, verbose=1)
    calls = [c for c in r._machine.calls if c[0] in ("cd_step", "score")]
    per_batch = [("cd_step", L.UPDATE_W), ("cd_step", L.UPDATE_C), ("cd_step", L.UPDATE_B), ("score", None)]
    rows = [128, 128, 44] * 2                                                    # remainder last (rbm.py:211,218), 2 epochs
    assert len(calls) == 4 * len(rows)
    for i, n in enumerate(rows):
        got = calls[4 * i:4 * i + 4]
        assert [(c[0], c[2] if c[0] == "cd_step" else None) for c in got] == per_batch      # rbm.py:214-216, 227-233
        assert all(c[1] == n for c in got)
    assert capsys.readouterr().out.count("score: 0.250000") == 6                  # printed per step (rbm.py:234)
    with pytest.raises(ValueError):
        _rbm(fakes, compat="bogus").fit(X, verbose=0)


def test_packed_input_goes_through_unchanged(fakes):
    x = (np.random.default_rng(0).random((300, 50)) < 0.5).astype(np.float32)
    p = PackedBits.from_dense(x)
    r = _rbm(fakes)
    r.fit(p, verbose=0)
    assert r._machine.V == 50 and r._machine.calls[1][0] == "fit_host" and r._machine.calls[1][1] is p
    ref = _rbm(fakes, compat="reference")
    ref.fit(p, verbose=0)                                                        # the step-by-step schedule slices dense rows
    assert [c[1] for c in ref._machine.calls if c[0] == "score"] == [128, 128, 44]


def test_persistent_chains_are_initialised_once(fakes):
    X = np.zeros((256, 40), np.float32)
    r = _rbm(fakes, persistent=True, epochs=2)
    r.fit(X, verbose=0)
    r.fit(X, verbose=0)
    chains = [c for c in r._machine.calls if c[0] == "set_chains"]
    assert len(chains) == 1 and chains[0][1] == (128, 40) and 0.4 < chains[0][2] < 0.6     # Bernoulli(0.5), seed 99
    assert all(c[3] in (0, 1) for c in r._machine.calls if c[0] == "fit_epoch")


def test_bad_dtype_and_list_of_one_input(fakes):
    with pytest.raises(ValueError):
        _rbm(fakes, dtype="fp8").fit(np.zeros((4, 4), np.float32), verbose=0)
    r = _rbm(fakes)
    r.fit([np.zeros((10, 4), np.float32)], verbose=0)                            # K.function style list of one array
    assert r._machine.V == 4


def test_dbn_fit_feeds_each_layer_with_the_sampled_states_of_the_one_below(fakes, monkeypatch, capsys):
    """dbn.py:51-55: fit layer l on V_p, then V_p <- transform(V_p) over the whole data set - here as device-resident
    data sets handed from layer to layer, each closed once consumed."""
    from keras_unsupervised_b200.ebm import dbn as D

    monkeypatch.setattr(D, "Dataset", FakeDataset)

    def transform_dataset(self, ds):
        self.calls.append(("transform_dataset", ds))
        out = FakeDataset(self.ctx, ds.shape[0], self.H, tag="hidden of %d" % self.H)
        FakeDataset.made.append(out)
        return out

    monkeypatch.setattr(FakeMachine, "transform_dataset", transform_dataset, raising=False)
    X = np.zeros((300, 50), np.float32)
    dbn = D.DBN()
    layers = [R.RBM({"batch_size": 128, "epochs": 2, "lr": 1e-3, "seed": 3}, d, name="l%d" % i,
                    mode=R.MODE_VISIBLE_BERNOULLI, context=fakes) for i, d in enumerate((40, 30, 20))]
    for layer in layers:
        dbn.add_stack(layer)
    assert [layer.seed for layer in layers] == [3, 3 + 1000003, 3 + 2 * 1000003]      # one Philox stream per layer
    dbn.fit(X, verbose=0)
    assert capsys.readouterr().out.split("\n")[:3] == ["Train l0.", "Train l1.", "Train l2."]      # dbn.py:53
    assert [(l._machine.V, l._machine.H) for l in layers] == [(50, 40), (40, 30), (30, 20)]
    up, h0, h1 = FakeDataset.made
    assert up.source is X and (h0.shape, h1.shape) == ((300, 40), (300, 30))
    for layer, ds in zip(layers, (up, h0, h1)):
        fits = [c for c in layer._machine.calls if c[0] == "fit_epoch"]
        assert len(fits) == 2 and all(c[1] is ds for c in fits)
    assert [c[0] for c in layers[2]._machine.calls].count("transform_dataset") == 0   # nothing above the top layer
    assert all(d.closed for d in FakeDataset.made)
