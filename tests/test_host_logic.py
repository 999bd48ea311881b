"""Host-side logic of the drop-in classes (no GPU): training-state precompute, likelihood gradient,
npz wire format, PCA selection, shard arithmetic."""
import os

import numpy as np
import pytest

import gp_emulator_b200 as gpe
from gp_emulator_b200 import GaussianProcess, MultivariateEmulator, k_fold_cross_validation
from gp_emulator_b200.sharding import shard_range
from oracle import gp_oracle as orc
from tests.conftest import golden


def test_set_params_matches_reference_state():
    g = golden("T")
    gp = GaussianProcess(g["inputs"], g["targets"])
    gp._set_params(g["theta"])
    assert orc.ref_err(gp.invQ, g["invQ"]) < 1e-7
    assert orc.ref_err(gp.invQt, g["invQt"]) < 1e-7
    invQ, invQt = orc.prepare_likelihood(g["inputs"], g["targets"], g["theta"])
    assert orc.ref_err(gp.invQ, invQ) < 1e-7 and orc.ref_err(gp.invQt, invQt) < 1e-7
    assert (gp.n, gp.D) == g["inputs"].shape


def test_partial_devs_is_the_gradient_of_loglikelihood():
    rs = np.random.RandomState(4)
    x = rs.random_sample((30, 3)); t = np.sin(x.sum(axis=1))
    gp = GaussianProcess(x, t)
    th = np.array([0.3, -0.2, 0.1, 0.2, -3.0])
    gp.loglikelihood(th)
    g = gp.partial_devs(th)
    for i in range(th.size):
        d = np.zeros_like(th); d[i] = 1e-6
        fd = (gp.loglikelihood(th + d) - gp.loglikelihood(th - d)) / 2e-6
        assert abs(fd - g[i]) < 1e-4 * max(1.0, abs(g[i]))


def test_learn_hyperparameters_reduces_cost(capsys):
    rs = np.random.RandomState(0)
    x = rs.random_sample((25, 2)); t = np.sin(3 * x[:, 0]) + x[:, 1] + 0.05 * rs.standard_normal(25)
    gp = GaussianProcess(x, t)
    np.random.seed(0)
    cost, theta = gp.learn_hyperparameters(n_tries=3)
    assert np.isfinite(cost) and theta.shape == (4,)
    assert cost <= gp.loglikelihood(np.zeros(4)) + 1e-9
    assert "minimum cost" in capsys.readouterr().out


def test_k_fold():
    folds = list(k_fold_cross_validation(range(10), 5))
    assert len(folds) == 5 and all(len(v) == 2 and len(t) == 8 for t, v in folds)


def test_multivariate_npz_roundtrip_and_pca(tmp_path):
    rs = np.random.RandomState(2)
    y = rs.random_sample((20, 3))
    wl = np.linspace(0, 1, 50)
    X = np.sin(np.outer(y[:, 0], wl) * 3) + np.outer(y[:, 1], wl ** 2) + 0.1 * np.outer(y[:, 2], np.ones(50)) + 1.0
    # choose thresh so that >= 2 PCs are kept, then fix hyper-parameters (no training, no GPU)
    s = np.linalg.svd(X, compute_uv=False)
    frac = s.cumsum() / s.sum()
    thresh = float(frac[1] + 1e-9)
    npcs = int(np.sum(frac <= thresh))
    hyp = np.tile(np.array([0., 0., 0., 0., -6.])[:, None], (1, npcs))
    mv = MultivariateEmulator(X=X, y=y, hyperparams=hyp, thresh=thresh)
    assert mv.n_pcs == npcs and mv.basis_functions.shape == (npcs, 50)
    assert np.allclose(mv.compress(X), orc.mv_compress(X, mv.basis_functions))
    f = os.path.join(tmp_path, "emu.npz")
    mv.dump_emulator(f)
    with np.load(f) as z:
        assert sorted(z.files) == sorted(["X", "y", "hyperparams", "thresh", "basis_functions", "n_pcs"])
    mv2 = MultivariateEmulator(dump=f)
    assert mv2.n_pcs == npcs and np.array_equal(mv2.basis_functions, mv.basis_functions)
    assert np.allclose(mv2.emulators[0].invQt, mv.emulators[0].invQt)
    with pytest.raises(ValueError):
        MultivariateEmulator(dump=f, X=X, y=y)
    with pytest.raises(ValueError):
        MultivariateEmulator(X=X)


def test_prosail_fixture_loads_as_reference_model():
    g = golden("P")
    gp = GaussianProcess(g["y"], g["train_data"][3])
    gp._set_params(g["hyperparams"][:, 3])
    assert orc.ref_err(gp.invQt, g["invQt"][3]) < 1e-5


@pytest.mark.parametrize("N,world", [(10, 1), (10, 3), (7, 8), (100000000, 8), (0, 4)])
def test_shard_range_partitions(N, world):
    ranges = [shard_range(N, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == N
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


def test_bind_host_to_gpu_is_a_noop_without_nvml_device():
    """No GPU here: the NUMA helper must report None and leave the affinity mask alone."""
    import os
    from gp_emulator_b200 import sharding
    before = os.sched_getaffinity(0)
    assert sharding.bind_host_to_gpu(0) is None or os.sched_getaffinity(0) <= before
    if sharding.bind_host_to_gpu(0) is None:
        assert os.sched_getaffinity(0) == before


def test_argument_checks_that_need_no_device():
    from gp_emulator_b200 import DeviceModel
    from gp_emulator_b200.training import DeviceTrainer, minimise_batched
    x = np.zeros((4, 2)); th = np.zeros(4); a = np.zeros(4); q = np.eye(4)
    with pytest.raises(ValueError):
        DeviceModel(x, th, a, q, symmetric_variance="sometimes")
    with pytest.raises(ValueError):
        DeviceModel(x, th[:2], a, q)                       # theta shorter than D + 1
    with pytest.raises(ValueError):
        DeviceModel(x, th, a[:3], q)
    with pytest.raises(ValueError):
        DeviceTrainer(x, np.zeros((2, 5)))                  # targets must be (T, M)
    assert minimise_batched(lambda t, i: None, [])[0] == []
    gp = GaussianProcess(x, a)
    with pytest.raises(ValueError):
        gp.predict([[0.0, 0.0]])                           # the reference indexes testing.shape: arrays only
    with pytest.raises(AssertionError):
        gp.predict(np.zeros((3, 5)))                       # wrong number of columns (reference :229 asserts)


def test_get_gpu_block_matches_reference_chunker():
    """GaussianProcess.get_gpu_block against block boundaries frozen from the reference (GaussianProcess.py:253-270)."""
    g = golden("B")
    gp = gpe.GaussianProcess(np.zeros((3, 2)), [])
    for i, (size, block) in enumerate(g["cases"]):
        a, b = gp.get_gpu_block(int(size), int(block))
        assert np.array_equal(a, g["start_%d" % i]) and np.array_equal(b, g["end_%d" % i]), (size, block)


def test_model_attribute_rebinding_is_counted():
    """MultivariateEmulator keys its device bank on each emulator's assignment counter."""
    gp = gpe.GaussianProcess(np.zeros((3, 2)), [])
    v0 = gp._version
    gp.theta = np.zeros(4)
    gp.invQt = np.zeros(3)
    assert gp._version == v0 + 2
    gp.targets = [1]                     # not a predict-relevant attribute
    assert gp._version == v0 + 2


def test_resolve_devices_and_closed_handles():
    from gp_emulator_b200 import engine
    assert engine.resolve_devices(3) == [3] and engine.resolve_devices([1, 0]) == [1, 0]
    with pytest.raises(ValueError):
        engine.resolve_devices("some")
    with pytest.raises(ValueError):
        engine.resolve_devices([])
    h = engine._Handle()
    h._own(None, lambda _h: None)
    h.close()
    with pytest.raises(gpe.GpemuError):
        h._h
