"""Drop-in ``GaussianProcess`` (same constructor, attributes and method names as the reference class in
gp_emulator/GaussianProcess.py:28-366) whose prediction runs on the B200 engine.

Training (``learn_hyperparameters`` and the invQ / invQt precompute) stays on the host in numpy by default, as
the scope requires; it is restated here so the class is usable under Python 3 (the reference file is Python 2).
``learn_hyperparameters(batched=True)`` serves the optimiser's cost / gradient evaluations from the GPU
(``training.py``); the state prediction reads is always built by the host ``_set_params``.  ``predict`` and ``hessian`` always go through libgpemu: there is no CPU prediction path, and the
``is_gpu`` / ``threshold`` arguments of the reference signature are accepted and ignored (chunking lives
below the C ABI).  Model attributes (``inputs, theta, invQ, invQt``) stay plain numpy arrays and are
re-read at every call -- the reference's own benchmark overwrites them between calls
(tests/benchmark.py:11-15) -- with the device copy cached on their contents.
"""
from __future__ import annotations

import warnings

import numpy as np

import ctypes

from .engine import DeviceModel, MultiDeviceModel, resolve_devices

_memcmp = ctypes.CDLL(None).memcmp
_memcmp.restype = ctypes.c_int
_memcmp.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]

# attributes whose rebinding changes what predict computes (MultivariateEmulator watches the counter)
_MODEL_ATTRS = frozenset(("inputs", "theta", "invQ", "invQt"))


def _same_bytes(a, b):
    """Exact (bitwise) equality of two C-contiguous float64 arrays: one memcmp, ~12 us for 500 KB."""
    return a.shape == b.shape and _memcmp(a.__array_interface__["data"][0], b.__array_interface__["data"][0], a.nbytes) == 0


def k_fold_cross_validation(X, K, randomise=False):
    """K (training, validation) partitions of X (reference GaussianProcess.py:9-26)."""
    import random
    items = list(X)
    if randomise:
        random.shuffle(items)
    for k in range(K):
        yield ([x for i, x in enumerate(items) if i % K != k], [x for i, x in enumerate(items) if i % K == k])


class GaussianProcess:
    """Squared-exponential (ARD) GP emulator.  ``inputs`` (Ntrain, Ninputs), ``targets`` (Ntrain,)."""

    # How predict decides whether the device copy of the model is still current (the reference re-reads
    # inputs / theta / invQ / invQt on every call, GaussianProcess.py:228-240):
    #   "auto"    (default) theta and invQt are always compared in full; the matrices inputs / invQ are compared
    #             bitwise against the uploaded copy (one memcmp) whenever that is cheap next to the call -- always while
    #             they hold <= 1 MB (M <= 360: 12 us), and for calls of >= 1000 points at any size -- and by identity +
    #             64 sampled elements for tiny calls on larger models (where an in-place edit needs invalidate_device());
    #   "full"    always the bitwise comparison;  "sampled"  always the fingerprint.
    cache_check = "auto"
    _FULL_CHECK_BYTES = 1 << 20
    _FULL_CHECK_POINTS = 1000

    def __init__(self, inputs, targets, device=0, symmetric_variance="auto"):
        """``device``: a GPU index, a list of indices, or ``"all"`` -- with more than one device a host batch is spread
        over all of them inside ONE ``predict`` call (``MultiDeviceModel``); the reference API has no notion of ranks.
        ``symmetric_variance``: ``"auto"`` (default) lets the device fold ``invQ`` onto its upper triangle when it is
        symmetric to 1e-6 -- any ``invQ`` that ``_prepare_likelihood`` produced -- which halves the tensor-core work of
        the variance (1.4-1.5x the throughput; same value to the rounding of the summation order), and keeps the general
        dense formula for a non-symmetric ``invQ`` such as the random one of the reference's benchmark
        (tests/benchmark.py:11-15); False always uses the dense formula, True always folds (``DeviceModel``)."""
        self._version = 0
        self.inputs = inputs
        self.targets = targets
        (self.n, self.D) = self.inputs.shape
        self.device = device
        self.symmetric_variance = symmetric_variance
        self._dev_model = None
        self._dev_key = None
        self._dev_copy = None

    def __setattr__(self, name, value):
        if name in _MODEL_ATTRS:
            object.__setattr__(self, "_version", getattr(self, "_version", 0) + 1)
        object.__setattr__(self, name, value)

    # ------------------------------------------------------------------ host training path (numpy)
    def _prepare_likelihood(self):
        """Q, invQ, invQt, log|Q| for the current theta (reference GaussianProcess.py:52-75)."""
        e = np.exp(self.theta)
        x = np.asarray(self.inputs, dtype=np.float64)
        diff2 = (x[:, None, :] - x[None, :, :]) ** 2
        self.Z = e[self.D] * np.exp(-0.5 * np.tensordot(diff2, e[: self.D], axes=([2], [0])))
        self.Q = self.Z + e[self.D + 1] * np.eye(self.n)
        self.invQ = np.linalg.inv(self.Q)
        self.invQt = np.dot(self.invQ, self.targets)
        self.logdetQ = 2.0 * np.sum(np.log(np.diag(np.linalg.cholesky(self.Q))))

    def loglikelihood(self, theta):
        """Negative log marginal likelihood at ``theta`` (reference GaussianProcess.py:78-95)."""
        self._set_params(theta)
        ll = 0.5 * self.logdetQ + 0.5 * np.dot(self.targets, self.invQt) + 0.5 * self.n * np.log(2.0 * np.pi)
        self.current_theta = theta
        self.current_loglikelihood = ll
        return ll

    def partial_devs(self, theta):
        """Gradient of ``loglikelihood`` w.r.t. theta (reference GaussianProcess.py:97-125)."""
        out = np.zeros(self.D + 2)
        x = np.asarray(self.inputs, dtype=np.float64)
        a = self.invQt
        for d in range(self.D):
            V = (x[:, d][:, None] - x[:, d][None, :]) ** 2 * self.Z
            out[d] = np.exp(self.theta[d]) * (np.dot(a, np.dot(V, a)) - np.sum(self.invQ * V)) / 4.0
        out[self.D] = 0.5 * np.sum(self.invQ * self.Z) - 0.5 * np.dot(a, np.dot(self.Z, a))
        noise = np.exp(self.theta[self.D + 1])
        out[self.D + 1] = 0.5 * np.trace(self.invQ) * noise - 0.5 * np.dot(a, a) * noise
        return out

    def _set_params(self, theta):
        """Fix the hyper-parameters and precompute what predict needs (reference GaussianProcess.py:127-139)."""
        self.theta = theta
        self._prepare_likelihood()

    def _learn(self, theta0, verbose):
        """One L-BFGS-B descent from ``theta0`` (reference GaussianProcess.py:141-181)."""
        from scipy.optimize import fmin_l_bfgs_b
        self._set_params(theta0)
        try:
            # (scipy >= 1.15 dropped the iprint argument the reference passes; verbose reports the result instead)
            res = fmin_l_bfgs_b(self.loglikelihood, theta0, fprime=self.partial_devs, factr=0.1, pgtol=1e-20)
            if verbose:
                print("L-BFGS-B: cost %e after %d evaluations (%s)" % (res[1], res[2]["funcalls"], res[2]["task"]))
        except np.linalg.LinAlgError:
            warnings.warn("Optimisation resulted in linear algebra error. Returning last loglikelihood "
                          "calculated, but this is fishy", RuntimeWarning)
            res = [self.current_theta, 9999]
        return res

    def learn_hyperparameters(self, n_tries=15, verbose=False, batched=False):
        """Multi-start fit; returns (min cost, theta) (reference GaussianProcess.py:183-209).

        ``batched=True`` runs the same ``n_tries`` L-BFGS-B descents from the same random starts, but in lockstep with
        every round of cost + gradient evaluations served by one GPU launch (``training.minimise_batched``); the
        default is the reference's sequential host path.
        """
        starts = 5.0 * (np.random.rand(n_tries, self.D + 2) - 0.5)
        if batched:
            from .training import DeviceTrainer, minimise_batched
            trainer = DeviceTrainer(self.inputs, self.targets, device=resolve_devices(self.device)[0])
            try:
                fits, _ = minimise_batched(trainer.evaluate, [(0, th) for th in starts], verbose=verbose)
            finally:
                trainer.close()
        else:
            fits = [self._learn(theta, verbose) for theta in starts]
        costs = np.array([T[1] for T in fits])
        params = [T[0] for T in fits]
        idx = int(np.argsort(costs)[0])
        print("After %d, the minimum cost was %e" % (n_tries, costs[idx]))
        self._set_params(params[idx])
        return costs[idx], params[idx]

    # ------------------------------------------------------------------ device prediction path
    def _device_model(self, n_points=None):
        """Device copy of the current numpy state, re-uploaded only when the arrays change.

        The reference's benchmark REBINDS the attributes between calls (tests/benchmark.py:11-15) and the reference
        itself re-reads them on every predict, so every call checks the state against what was uploaded; see
        ``cache_check`` for how.  ``n_points`` is the size of the call the model is fetched for.
        """
        def fingerprint(a):      # ~1 us: identity, shape and the bytes of ~64 evenly spaced elements
            q = np.asarray(a)
            return (id(a), q.shape, q.dtype.str, q.reshape(-1)[::max(1, q.size // 64)].tobytes())

        invQ = getattr(self, "invQ", None)
        mode = self.cache_check
        mats = [self.inputs] + ([invQ] if invQ is not None else [])
        key = (np.asarray(self.theta).tobytes(), np.asarray(self.invQt).tobytes(), str(self.device),
               self.symmetric_variance, invQ is not None)
        ok = self._dev_model is not None and key == self._dev_key
        if ok:
            for a, (copy, mark) in zip(mats, self._dev_copy):
                q = np.asarray(a)
                if mode == "full" or (mode == "auto" and (q.nbytes <= self._FULL_CHECK_BYTES or (
                        n_points is not None and n_points >= self._FULL_CHECK_POINTS))):
                    ok = _same_bytes(np.ascontiguousarray(q, dtype=np.float64), copy)
                else:
                    ok = fingerprint(a) == mark
                if not ok:
                    break
        if not ok:
            if self._dev_model is not None:
                self._dev_model.close()
            devices = resolve_devices(self.device)
            if len(devices) > 1:
                self._dev_model = MultiDeviceModel(self.inputs, self.theta, self.invQt, invQ, devices=devices,
                                                   symmetric_variance=self.symmetric_variance)
            else:
                self._dev_model = DeviceModel(self.inputs, self.theta, self.invQt, invQ, device=devices[0],
                                              symmetric_variance=self.symmetric_variance)
            # private copies of what was uploaded (0.5 MB at M = 250) for the bitwise check, and the fingerprints
            self._dev_copy = [(np.array(a, dtype=np.float64, order="C"), fingerprint(a)) for a in mats]
            self._dev_key = key
        return self._dev_model

    def invalidate_device(self):
        """Drop the cached device copy (next predict re-uploads the model)."""
        if self._dev_model is not None:
            self._dev_model.close()
        self._dev_model, self._dev_key, self._dev_copy = None, None, None

    def predict(self, testing, do_unc=True, do_deriv=True, is_gpu=True, precision=np.float64, threshold=2e5,
                out=None, pinned=None):
        """Mean, variance and input gradient at ``testing`` (N, D)  (reference GaussianProcess.py:327-341).

        Returns ``(mu, var, deriv)``; ``(mu, deriv)`` if ``do_unc`` is False (as the reference's CPU branch,
        :248-251); ``(mu, var)`` / ``mu`` when ``do_deriv`` is False (the upstream-style signature).
        ``deriv`` is (N, D).  Computation is FP64 on the GPU unless ``precision=np.float32`` (or a float32 CUDA
        tensor) is given and M <= 1024: then the single-precision tensor-core path runs (for larger M the FP64
        results are cast).
        ``testing`` may also be a float64 torch CUDA tensor, in which case torch tensors are returned.
        ``out`` (dict with any of "mu", "var", "deriv") supplies preallocated result buffers; ``pinned``
        chooses the memory of freshly allocated host results (None: page-locked from the second call of the same
        size on; see ``DeviceModel.predict``).
        """
        if getattr(testing, "ndim", None) != 2 and not hasattr(testing, "dim"):
            raise ValueError("testing must always be a 2-D array (N, D)")
        if testing.shape[1] != self.D:
            raise AssertionError("testing has %d columns, model has D = %d" % (testing.shape[1], self.D))
        dm = self._device_model(testing.shape[0])
        is_t = hasattr(testing, "dim")
        f32_in = (is_t and str(testing.dtype) == "torch.float32") or (not is_t and precision is np.float32)
        if f32_in and dm.M <= 1024 and dm.D <= 32 and out is None:
            # the reference's FP32 GPU build (precision=np.float32, GaussianProcess.py:289-316): single precision
            # end to end, variance contraction on the tensor cores (tcgen05, TF32 inputs)
            o = dm.predict_f32(testing, want_var=do_unc, want_deriv=do_deriv)
            res = [o["mu"]] + ([o["var"]] if do_unc else []) + ([o["deriv"]] if do_deriv else [])
            return tuple(res) if len(res) > 1 else res[0]
        out = dm.predict(testing, want_var=do_unc, want_deriv=do_deriv, out=out, pinned=pinned)
        res = [out["mu"]]
        if do_unc:
            res.append(out["var"])
        if do_deriv:
            res.append(out["deriv"])
        if isinstance(res[0], np.ndarray) and precision is not np.float64:
            res = [precision(r) for r in res]
        return tuple(res) if len(res) > 1 else res[0]

    # the reference exposes these names as well; all of them are the device path here
    def gpu_predict(self, testing, precision=np.float64, threshold=2e5):
        """``(result, error, deriv)`` as the reference GPU branch returns them (GaussianProcess.py:273-323)."""
        return self.predict(testing, do_unc=True, do_deriv=True, precision=precision, threshold=threshold)

    def cpu_predict(self, testing, do_unc=True):
        """The reference's numpy branch by name (GaussianProcess.py:211-251): ``(mu, var, deriv)`` or ``(mu, deriv)``.
        There is no CPU prediction path in this package: it runs on the device like ``predict``."""
        return self.predict(testing, do_unc=do_unc, do_deriv=True)

    def get_gpu_block(self, size, block_size):
        """Start / end indices of the blocks the reference's host chunker walks (GaussianProcess.py:253-270): blocks of
        ``block_size`` points, the last two sharing their points equally so the tail is not tiny.  Kept for callers
        that size their own buffers with it; the library's chunking lives below the C ABI and does not use it."""
        size, block_size = int(size), int(block_size)
        ind_start = np.arange(0, size, block_size, dtype=np.int64)
        ind_end = np.append(ind_start[1:], size).astype(np.int64)
        if ind_start.size > 1:
            ind_end[-2] = ind_start[-2] + (ind_end[-1] - ind_start[-2]) // 2
            ind_start[-1] = ind_end[-2]
        assert np.all(ind_end - ind_start <= block_size)
        return ind_start, ind_end

    def hessian(self, testing):
        """(N, D, D) Hessian of the predictive mean (reference GaussianProcess.py:345-366)."""
        if testing.shape[1] != self.D:
            raise AssertionError("testing has %d columns, model has D = %d" % (testing.shape[1], self.D))
        out = self._device_model(testing.shape[0]).predict(testing, want_mu=False, want_var=False, want_deriv=False,
                                                           want_hess=True)
        return out["hess"]
