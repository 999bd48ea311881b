"""Host utilities the reference package exports beside the emulators (gp_emulator/__init__.py:3-4): ``lhd`` against
frozen reference designs, ``EmulatorStorage`` round trips.  No GPU needed: a stored emulator only uploads at predict."""
import shelve

import numpy as np
import pytest
import scipy.stats as ss

from gp_emulator_b200 import EmulatorStorage, GaussianProcess, MultivariateEmulator, lhd
from tests.conftest import golden

D0, D1, D2, D3 = ss.uniform(loc=-1, scale=2), ss.norm(loc=0, scale=1), ss.beta(2, 5), ss.expon(scale=1 / 1.5)


def test_lhd_reproduces_reference_designs(capsys):
    g = golden("U")
    np.random.seed(5); assert np.array_equal(lhd(dist=D0, size=5), g["single"])
    np.random.seed(6); assert np.array_equal(lhd(dist=D1, size=7, dims=5), g["dims"])
    np.random.seed(7); assert np.array_equal(lhd(dist=(D1, D2, D3), size=6), g["multi"])
    np.random.seed(8); assert np.array_equal(lhd(dist=(D0, D1, D2, D3), size=100), g["big"])
    np.random.seed(9)
    assert np.array_equal(lhd(dist=(D0, D1, D2, D3), size=12, form="spacefilling", iterations=7), g["space"])
    assert "Optimized Distance" in capsys.readouterr().out


def test_lhd_is_a_latin_hypercube_and_edge_cases(capsys):
    np.random.seed(1)
    x = lhd(dist=D0, size=40, dims=3)
    assert x.shape == (40, 3)
    strata = np.floor((x + 1.0) / 2.0 * 40).astype(int)
    for k in range(3):                                    # exactly one sample per stratum and column
        assert sorted(strata[:, k]) == list(range(40))
    assert lhd(dist=None, size=5) is None and lhd(dist=D0, size=None) is None
    with pytest.raises(NotImplementedError):
        lhd(dist=D0, size=4, form="orthogonal")
    with pytest.raises(ValueError):
        lhd(dist=D0, size=4, form="nonsense")
    lhd(dist=(D0, D1), size=8, showcorrelations=True)
    assert "Variance Inflation Factor" in capsys.readouterr().out


def test_emulator_storage_round_trip(tmp_path, capsys):
    rs = np.random.RandomState(0)
    x = rs.random_sample((12, 2)); t = np.sin(x.sum(axis=1))
    gp = GaussianProcess(x, t)
    gp._set_params(np.array([0.1, -0.3, 0.2, -4.0]))
    y = rs.random_sample((10, 2))
    X = np.sin(np.linspace(0, 1, 20)[None, :] * y[:, :1] * 3.0) + y[:, 1:2]
    s = np.linalg.svd(X, compute_uv=False)
    P = int(np.sum(s.cumsum() / s.sum() <= 0.999))
    hyper = np.tile(np.array([0.1, 0.2, 0.0, -5.0])[:, None], (1, P))
    mv = MultivariateEmulator(X=X, y=y, hyperparams=hyper, thresh=0.999)
    assert mv.n_pcs == P >= 2
    store = EmulatorStorage(str(tmp_path / "emus"))
    with pytest.raises(IOError):
        store.get_keys()
    store.dump_emulator(gp, (30, 0, 40))
    store.dump_emulator(mv, [30, 0, 41])                 # a list tag: unreadable in the reference (save_emulators.py:43 vs :78)
    store.dump_emulator(gp, "plain")
    assert sorted(store.get_keys()) == sorted(["(30, 0, 40)", "(30, 0, 41)", "plain"])
    g2 = store.get_emulator((30, 0, 40))
    assert np.array_equal(g2.inputs, x) and np.array_equal(g2.theta, gp.theta) and np.array_equal(g2.invQt, gp.invQt)
    m2 = store.get_emulator((30, 0, 41))
    assert m2.n_pcs == P and np.array_equal(m2.basis_functions, mv.basis_functions)
    assert np.array_equal(m2.hyperparams, mv.hyperparams)
    assert np.array_equal(m2.emulators[1].invQt, mv.emulators[1].invQt)
    assert np.array_equal(store.get_emulator("plain").targets, t)
    with pytest.raises(TypeError):
        store.dump_emulator(object(), "x")
    # a record as the REFERENCE writes it (only the "input" spelling, list tag stored as repr(list))
    with shelve.open(str(tmp_path / "emus")) as db:
        db[repr([1, 2])] = {"input": x, "targets": t, "theta": gp.theta}
    g3 = store.get_emulator([1, 2])
    assert np.array_equal(g3.invQ, gp.invQ)
