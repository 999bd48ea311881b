"""Batched training objective: oracle pinned to the reference's loglikelihood / partial_devs, the
lockstep L-BFGS-B driver (host logic, no GPU), and -- marked gpu -- gpe_trainer_eval against both.

Tolerance: 1e-10 in the reference's metric max|x - ref| / max|ref| per problem, at thetas drawn the way the reference
draws its L-BFGS-B starts (5 (U - 0.5), gp_emulator/GaussianProcess.py:201; cond(Q) <= ~1e5 there).
"""
import ctypes as C

import numpy as np
import pytest

from gp_emulator_b200 import GaussianProcess, GpemuError, MultivariateEmulator
from gp_emulator_b200.training import DeviceTrainer, minimise_batched
from oracle import gp_oracle as orc
from tests.conftest import golden

TOL = 1e-10


def _golden_cases():
    g = golden("L")
    for tag in "abc":
        M, D, T, B, seed = (int(v) for v in g["shape_" + tag])
        yield tag, orc.make_training_problem(M, D, T, B, seed), g["ll_" + tag], g["grad_" + tag]


def test_oracle_training_objective_matches_reference():
    for tag, (inputs, targets, thetas, tidx), ll_ref, grad_ref in _golden_cases():
        for n in range(thetas.shape[0]):
            ll, grad = orc.loglikelihood_and_grad(inputs, targets[tidx[n]], thetas[n])
            assert abs(ll - ll_ref[n]) <= 1e-13 * abs(ll_ref[n]), tag
            assert orc.ref_err(grad, grad_ref[n]) < 1e-13, tag


def test_dropin_class_training_objective_matches_reference():
    for tag, (inputs, targets, thetas, tidx), ll_ref, grad_ref in _golden_cases():
        if inputs.shape[0] > 100:
            continue
        for n in range(thetas.shape[0]):
            gp = GaussianProcess(inputs, targets[tidx[n]])
            assert abs(gp.loglikelihood(thetas[n]) - ll_ref[n]) <= 1e-12 * abs(ll_ref[n])
            assert orc.ref_err(gp.partial_devs(thetas[n]), grad_ref[n]) < 1e-11


def _host_evaluate(inputs, targets):
    def evaluate(thetas, tidx):
        ll = np.empty(len(thetas)); grad = np.empty((len(thetas), inputs.shape[1] + 2)); st = np.zeros(len(thetas), dtype=np.int32)
        for n, th in enumerate(thetas):
            gp = GaussianProcess(inputs, targets[tidx[n]])
            try:
                ll[n] = gp.loglikelihood(th)
                grad[n] = gp.partial_devs(th)
            except np.linalg.LinAlgError:          # what gpe_trainer_eval reports as status 1
                ll[n], grad[n], st[n] = np.nan, np.nan, 1
        return ll, grad, st
    return evaluate


def test_lockstep_descents_equal_sequential_descents():
    """The batching driver changes WHEN evaluations happen, not what each descent sees: with the same evaluator every
    descent must end exactly where the sequential reference loop (GaussianProcess._learn) ends."""
    inputs, targets, thetas, tidx = orc.make_training_problem(M=24, D=2, T=2, B=5, seed=3)
    from gp_emulator_b200 import training
    assert training._reverse_communication_ok(), "the installed scipy should pass the reverse-communication self-test"
    refs = [GaussianProcess(inputs, targets[tidx[n]])._learn(thetas[n], False) for n in range(5)]
    for driver in (None, "reverse-communication", "threads"):
        fits, stats = minimise_batched(_host_evaluate(inputs, targets), [(int(tidx[n]), thetas[n]) for n in range(5)],
                                       driver=driver)
        assert stats["evaluations"] >= 5 and stats["rounds"] <= stats["evaluations"]
        assert stats["driver"] == (driver or "reverse-communication")
        for n in range(5):
            assert np.array_equal(fits[n][0], refs[n][0]) and fits[n][1] == refs[n][1], (driver, n)
    with pytest.raises(ValueError):
        minimise_batched(_host_evaluate(inputs, targets), [(0, thetas[0])], driver="nonsense")


def test_lockstep_driver_failure_conventions():
    inputs, targets, thetas, tidx = orc.make_training_problem(M=12, D=2, T=1, B=3, seed=5)
    host = _host_evaluate(inputs, targets)

    def one_bad(th, ti):                    # every evaluation of target row 1 reports a non-positive-definite covariance
        ll, g, st = host(th, np.zeros_like(ti))
        st[np.asarray(ti) == 1] = 1
        return ll, g, st
    for driver in ("reverse-communication", "threads"):
        fits, _ = minimise_batched(one_bad, [(0, thetas[0]), (1, thetas[1]), (0, thetas[2])], driver=driver)
        assert fits[1][1] == 9999 and np.array_equal(fits[1][0], thetas[1])      # reference GaussianProcess.py:174-179
        for n in (0, 2):                        # the other descents are not disturbed by the drop-out
            ref = GaussianProcess(inputs, targets[0])._learn(thetas[n], False)
            assert np.array_equal(fits[n][0], ref[0]) and fits[n][1] == ref[1]

    def broken(th, ti):
        raise RuntimeError("device lost")
    for driver in ("reverse-communication", "threads"):
        with pytest.raises(RuntimeError):
            minimise_batched(broken, [(0, thetas[0]), (0, thetas[1])], driver=driver)

    def not_pd(th, ti):
        ll, g, st = host(th, ti)
        st[:] = 1
        return ll, g, st
    for driver in ("reverse-communication", "threads"):
        fits, _ = minimise_batched(not_pd, [(0, thetas[0])], driver=driver)
        assert fits[0][1] == 9999 and np.array_equal(fits[0][0], thetas[0])     # reference GaussianProcess.py:174-179


def test_trainer_argument_validation_and_no_fallback(lib):
    h = C.c_void_p()
    x = np.zeros((4, 2)); t = np.zeros((1, 4))
    assert lib.gpe_trainer_create(0, 0, 2, 1, x.ctypes.data, t.ctypes.data, C.byref(h)) == -1
    assert lib.gpe_trainer_create(0, 4, 2, 1, None, t.ctypes.data, C.byref(h)) == -1
    assert lib.gpe_trainer_create(0, 2000, 2, 1, x.ctypes.data, t.ctypes.data, C.byref(h)) == -4
    assert lib.gpe_trainer_create(0, 4, 40, 1, x.ctypes.data, t.ctypes.data, C.byref(h)) == -4
    assert lib.gpe_trainer_eval(None, 1, None, None, None, None, None) == -1
    if lib.gpe_device_count() == 0:
        with pytest.raises(GpemuError):
            DeviceTrainer(np.zeros((4, 2)), np.zeros(4))
        with pytest.raises(GpemuError):     # batched=True never falls back to the host loop
            GaussianProcess(np.random.rand(6, 2), np.random.rand(6)).learn_hyperparameters(n_tries=1, batched=True)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_trainer_eval_matches_reference_goldens(lib):
    assert lib.gpe_device_count() > 0
    n0 = lib.gpe_launch_count()
    for tag, (inputs, targets, thetas, tidx), ll_ref, grad_ref in _golden_cases():
        tr = DeviceTrainer(inputs, targets)
        ll, grad, st = tr.evaluate(thetas, tidx)
        tr.close()
        assert not st.any()
        assert np.max(np.abs(ll - ll_ref) / np.abs(ll_ref)) < TOL, tag
        for n in range(len(ll)):
            assert orc.ref_err(grad[n], grad_ref[n]) < TOL, (tag, n)
    assert lib.gpe_launch_count() - n0 >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("M,D,T,B", [(1, 1, 1, 2), (2, 1, 1, 3), (31, 3, 2, 7), (64, 5, 4, 300), (65, 7, 1, 1),
                                     (129, 32, 2, 5), (300, 6, 3, 4), (513, 3, 1, 2)])
def test_trainer_eval_shapes_against_oracle(M, D, T, B):
    inputs, targets, thetas, tidx = orc.make_training_problem(M, D, T, B, seed=100 + M)
    tr = DeviceTrainer(inputs, targets)
    ll, grad, st = tr.evaluate(thetas, tidx)
    ll2, grad2, _ = tr.evaluate(thetas[::-1].copy(), tidx[::-1].copy())      # order / batch position must not matter
    tr.close()
    assert not st.any()
    assert np.array_equal(ll, ll2[::-1]) and np.array_equal(grad, grad2[::-1])
    for n in range(0, B, max(1, B // 8)):
        ll_o, grad_o = orc.loglikelihood_and_grad(inputs, targets[tidx[n]], thetas[n])
        assert abs(ll[n] - ll_o) <= TOL * abs(ll_o)
        assert orc.ref_err(grad[n], grad_o) < TOL


@pytest.mark.gpu
def test_trainer_default_target_and_reuse():
    inputs, targets, thetas, _ = orc.make_training_problem(40, 3, 1, 4, seed=8)
    tr = DeviceTrainer(inputs, targets[0])
    ll, grad, st = tr.evaluate(thetas)                   # target_index = None -> row 0
    ll1, grad1, _ = tr.evaluate(thetas[2])               # one theta, after a larger batch (workspace reuse)
    assert np.array_equal(ll1, ll[2:3]) and np.array_equal(grad1[0], grad[2])
    with pytest.raises(GpemuError):
        tr.evaluate(thetas, np.array([0, 0, 1, 0]))      # target row out of range
    tr.close()
    with pytest.raises(GpemuError):
        tr.evaluate(thetas)


@pytest.mark.gpu
def test_trainer_batches_larger_than_one_launch_and_shared_by_threads():
    """More problems than one launch takes (8 per SM) are walked in chunks; threads may share a trainer (serialised)."""
    import threading
    inputs, targets, thetas, tidx = orc.make_training_problem(9, 2, 3, 1500, seed=12)
    tr = DeviceTrainer(inputs, targets)
    ll, grad, st = tr.evaluate(thetas, tidx)
    assert not st.any()
    for n in (0, 1183, 1184, 1185, 1499):
        ll_o, grad_o = orc.loglikelihood_and_grad(inputs, targets[tidx[n]], thetas[n])
        assert abs(ll[n] - ll_o) <= TOL * abs(ll_o) and orc.ref_err(grad[n], grad_o) < TOL
    got = {}

    def work(k):
        got[k] = tr.evaluate(thetas[k * 100:(k + 1) * 100], tidx[k * 100:(k + 1) * 100])
    th = [threading.Thread(target=work, args=(k,)) for k in range(6)]
    [t.start() for t in th]
    [t.join() for t in th]
    tr.close()
    for k in range(6):
        assert np.array_equal(got[k][0], ll[k * 100:(k + 1) * 100]) and np.array_equal(got[k][1], grad[k * 100:(k + 1) * 100])


@pytest.mark.gpu
def test_trainer_flags_non_positive_definite_covariance():
    """Duplicate training inputs and (numerically) zero noise make Q singular: np.linalg.cholesky raises in the reference
    (GaussianProcess.py:73-75); here the problem comes back with status 1 and NaNs, its neighbours untouched."""
    inputs, targets, thetas, _ = orc.make_training_problem(30, 2, 1, 3, seed=9)
    inputs[7] = inputs[3]
    thetas[1, -1] = -800.0                                # exp(-800) == 0: no jitter on the diagonal
    tr = DeviceTrainer(inputs, targets)
    ll, grad, st = tr.evaluate(thetas)
    tr.close()
    assert list(st) == [0, 1, 0] and np.isnan(ll[1]) and np.isnan(grad[1]).all()
    huge = thetas.copy()
    huge[2, 0] = 800.0                                    # exp(800) overflows: non-finite results are failures too
    ll_h, grad_h, st_h = DeviceTrainer(inputs, targets).evaluate(huge)
    assert st_h[2] == 1 and np.isnan(ll_h[2]) and np.isnan(grad_h[2]).all() and st_h[0] == 0
    with pytest.raises(np.linalg.LinAlgError):
        orc.loglikelihood_and_grad(inputs, targets[0], thetas[1])
    for n in (0, 2):
        ll_o, grad_o = orc.loglikelihood_and_grad(inputs, targets[0], thetas[n])
        assert abs(ll[n] - ll_o) <= TOL * abs(ll_o) and orc.ref_err(grad[n], grad_o) < TOL


@pytest.mark.gpu
def test_batched_learn_hyperparameters_follows_the_host_descents(capsys):
    """Same starts, same optimiser, cost and gradient equal to ~1e-14: the batched fit must find the host fit's optimum."""
    rs = np.random.RandomState(0)
    x = rs.random_sample((40, 2)); t = np.sin(3 * x[:, 0]) + x[:, 1] + 0.05 * rs.standard_normal(40)
    np.random.seed(7)
    host = GaussianProcess(x, t)
    c_host, th_host = host.learn_hyperparameters(n_tries=4)
    np.random.seed(7)
    dev = GaussianProcess(x, t)
    c_dev, th_dev = dev.learn_hyperparameters(n_tries=4, batched=True)
    assert abs(c_dev - c_host) <= 1e-6 * max(1.0, abs(c_host))
    # the state predict reads comes from the host _set_params at the chosen theta
    fresh = GaussianProcess(x, t)
    fresh._set_params(th_dev)
    assert np.array_equal(dev.invQ, fresh.invQ) and np.array_equal(dev.invQt, fresh.invQt)
    invQ, invQt = orc.prepare_likelihood(x, t, th_dev)
    assert orc.ref_err(dev.invQ, invQ) < 1e-7 and orc.ref_err(dev.invQt, invQt) < 1e-7
    mu, var, _ = dev.predict(x[:5])
    assert np.max(np.abs(mu - t[:5])) < 0.5


@pytest.mark.gpu
def test_multivariate_emulator_batched_training(capsys):
    rs = np.random.RandomState(2)
    y = rs.random_sample((30, 3))
    wl = np.linspace(0.0, 1.0, 50)
    X = np.sin(4 * wl[None, :] * y[:, :1]) + y[:, 1:2] * wl[None, :] ** 2 + 0.2 * y[:, 2:3] + 0.02 * rs.standard_normal((30, 50))
    np.random.seed(3)
    mv = MultivariateEmulator(X=X, y=y, thresh=0.93, n_tries=3, batched_training=True)
    assert mv.training_stats["evaluations"] > mv.training_stats["rounds"] >= 1
    assert mv.hyperparams.shape == (5, mv.n_pcs) and np.isfinite(mv.hyperparams).all()
    np.random.seed(3)
    ref = MultivariateEmulator(X=X, y=y, thresh=0.93, n_tries=3)
    for a, b in zip(mv.emulators, ref.emulators):        # same starts -> same optima (up to the optimiser's tolerance)
        ca, cb = a.loglikelihood(a.theta), b.loglikelihood(b.theta)
        assert abs(ca - cb) <= 1e-5 * max(1.0, abs(cb))
    fwd, _ = mv.predict(y[4])
    assert fwd.shape == (50,) and np.max(np.abs(fwd - X[4])) < 0.5


@pytest.mark.gpu
def test_fit_bank_trains_all_bands_in_one_batch(capsys):
    from gp_emulator_b200.training import fit_bank
    rs = np.random.RandomState(4)
    x = rs.random_sample((36, 2))
    targets = np.stack([np.sin(3 * x[:, 0]) + x[:, 1], np.cos(2 * x[:, 1]) * x[:, 0], x[:, 0] ** 2 - x[:, 1]])
    targets = targets + 0.03 * rs.standard_normal(targets.shape)
    np.random.seed(11)
    gps, stats = fit_bank(x, targets, n_tries=3)
    assert len(gps) == 3 and stats["evaluations"] > stats["rounds"]
    np.random.seed(11)
    for e, gp in enumerate(gps):                       # same draws, same optimiser as E sequential host fits
        host = GaussianProcess(x, targets[e])
        c_host, _ = host.learn_hyperparameters(n_tries=3)
        assert abs(gp.fit_cost - c_host) <= 1e-5 * max(1.0, abs(c_host))
        mu, var, _ = gp.predict(x)
        assert np.max(np.abs(mu - targets[e])) < 0.3 and (var > -1e-9).all()
