"""Batched hyper-parameter fitting on the device (the caller on the input side of predict).

The reference fits each GP with ``n_tries`` independent L-BFGS-B descents (gp_emulator/GaussianProcess.py:141-209), and a
``MultivariateEmulator`` does that once per principal component on the SAME training inputs
(gp_emulator/multivariate_gp.py:176-186): n_pcs x n_tries descents, each evaluation a fresh M x M inverse in numpy
(~0.1 s at M = 250, D = 10).  Here the descents stay what they are -- scipy's ``fmin_l_bfgs_b`` with the reference's
settings (``factr=0.1, pgtol=1e-20``), one per start, so every descent follows the reference's trajectory to rounding --
but they advance in lockstep, and every round of function + gradient requests is served by ONE call of
``gpe_trainer_eval`` (one CTA per (theta, target) problem, csrc/train.cu).

The final state of the chosen theta (``invQ``, ``invQt``) is still produced by the host ``_set_params``
(GaussianProcess.py:127-139), so prediction inputs are bit-identical to the reference's.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib


class DeviceTrainer:
    """Training inputs (M, D) and T target vectors (T, M) resident on one GPU; evaluates batches of thetas."""

    def __init__(self, inputs, targets, device=0):
        self.inputs = _lib.f64c(inputs)
        self.targets = _lib.f64c(np.atleast_2d(targets))
        self.M, self.D = self.inputs.shape
        self.T = self.targets.shape[0]
        if self.targets.shape[1] != self.M:
            raise ValueError("targets must be (T, %d), got %r" % (self.M, self.targets.shape))
        self._h = C.c_void_p()
        _lib.check(_lib.load().gpe_trainer_create(int(device), self.M, self.D, self.T, self.inputs.ctypes.data,
                                                  self.targets.ctypes.data, C.byref(self._h)))

    def evaluate(self, thetas, target_index=None):
        """``(loglik (B,), grad (B, D + 2), status (B,))`` for thetas (B, D + 2); status 1 = Q not positive definite."""
        thetas = _lib.f64c(np.atleast_2d(thetas))
        B = thetas.shape[0]
        if thetas.shape[1] != self.D + 2:
            raise ValueError("thetas must be (B, %d)" % (self.D + 2))
        tidx = None
        if target_index is not None:
            tidx = np.ascontiguousarray(target_index, dtype=np.int32)
            if tidx.shape != (B,):
                raise ValueError("target_index must be (B,)")
        ll = np.empty(B)
        grad = np.empty((B, self.D + 2))
        status = np.empty(B, dtype=np.int32)
        if self._h is None:
            raise _lib.GpemuError("trainer is closed")
        _lib.check(_lib.load().gpe_trainer_eval(self._h, B, _lib.addr(tidx), thetas.ctypes.data, ll.ctypes.data,
                                                grad.ctypes.data, status.ctypes.data))
        return ll, grad, status

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().gpe_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Rendezvous:
    """Collects one request per live descent, then serves them with one batched evaluation.

    A descent thread posts (target, theta) and blocks; the thread whose request completes the round (every live
    descent has one pending) runs the batch and wakes the others.  A descent that ends drops out of the head count.
    """

    def __init__(self, evaluate, n_workers):
        self._evaluate = evaluate
        self._cv = threading.Condition()
        self._active = n_workers
        self._pending = []
        self._results = {}
        self._error = None
        self.rounds = 0
        self.evaluations = 0

    def _flush(self):
        batch, self._pending = self._pending, []
        try:
            ll, grad, status = self._evaluate(np.stack([b[2] for b in batch]), np.array([b[1] for b in batch], dtype=np.int32))
            for n, b in enumerate(batch):
                self._results[b[0]] = (float(ll[n]), np.array(grad[n]), int(status[n]))
        except BaseException as exc:          # wake everybody, each descent re-raises
            self._error = exc
            for b in batch:
                self._results[b[0]] = None
        self.rounds += 1
        self.evaluations += len(batch)
        self._cv.notify_all()

    def call(self, wid, tidx, theta):
        with self._cv:
            self._pending.append((wid, int(tidx), np.array(theta, dtype=np.float64)))
            if len(self._pending) >= self._active:
                self._flush()
            else:
                while wid not in self._results:
                    self._cv.wait()
            res = self._results.pop(wid)
            if res is None:
                raise self._error
            return res

    def leave(self):
        with self._cv:
            self._active -= 1
            if self._pending and len(self._pending) >= self._active:
                self._flush()


# ---- single-thread lockstep driver ---------------------------------------------------------------------------------
# scipy's L-BFGS-B is a reverse-communication routine: ``setulb`` returns to its caller whenever it needs the cost and
# gradient at a new point.  ``fmin_l_bfgs_b`` hides that behind a callback, which is why the portable driver below needs
# one thread per descent; driving ``setulb`` directly lets ONE thread advance every descent to its next request, evaluate
# the batch, and hand the values back -- no thread hand-offs (they cost ~4x the optimiser's own work per evaluation).
# ``setulb`` is private to scipy, so this driver mirrors the loop of ``scipy.optimize._lbfgsb_py._minimize_lbfgsb`` and is
# only used after a self-test has reproduced ``fmin_l_bfgs_b`` bit for bit on the installed scipy (else: threads).
_RC_STATE = {"checked": False, "ok": False}


class _Descent:
    """Workspace of one L-BFGS-B minimisation in reverse-communication form (m = 10, maxls = 20, no bounds,
    maxfun = maxiter = 15000: the defaults ``fmin_l_bfgs_b`` runs with in the reference call)."""
    M_CORR, MAXLS, MAXFUN, MAXITER = 10, 20, 15000, 15000

    def __init__(self, setulb, int_dtype, x0, factr, pgtol):
        n = x0.size
        m = self.M_CORR
        self._setulb, self.factr, self.pgtol = setulb, factr, pgtol
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.nbd = np.zeros(n, dtype=int_dtype)
        self.low = np.zeros(n, dtype=np.float64)
        self.up = np.zeros(n, dtype=np.float64)
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=int_dtype)
        self.task = np.zeros(2, dtype=int_dtype)
        self.ln_task = np.zeros(2, dtype=int_dtype)
        self.lsave = np.zeros(4, dtype=int_dtype)
        self.isave = np.zeros(44, dtype=int_dtype)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.nit = 0
        self.nfev = 0

    def advance(self):
        """Run the optimiser until it asks for cost + gradient at ``self.x`` (True) or stops (False)."""
        while True:
            self._setulb(self.M_CORR, self.x, self.low, self.up, self.nbd, self.f, self.g, self.factr, self.pgtol,
                         self.wa, self.iwa, self.task, self.lsave, self.isave, self.dsave, self.MAXLS, self.ln_task)
            if self.task[0] == 3:
                return True
            if self.task[0] == 1:
                self.nit += 1
                if self.nit >= self.MAXITER:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > self.MAXFUN:
                    self.task[0], self.task[1] = 5, 502
            else:
                return False

    def supply(self, f, g):
        self.f = float(f)
        self.g = np.array(g, dtype=np.float64)
        self.nfev += 1


def _setulb_handle():
    from scipy.optimize import _lbfgsb_py as mod
    return mod._lbfgsb.setulb, (np.int64 if mod.HAS_ILP64 else np.int32)


def _minimise_reverse_communication(evaluate, starts, verbose=False):
    setulb, int_dtype = _setulb_handle()
    runs = [_Descent(setulb, int_dtype, np.asarray(s[1], dtype=np.float64).ravel(), 0.1, 1e-20) for s in starts]
    tidx = [int(s[0]) for s in starts]
    last_good = [np.array(s[1], dtype=np.float64) for s in starts]
    failed = [False] * len(runs)
    pending = [k for k, r in enumerate(runs) if r.advance()]
    rounds = evaluations = 0
    while pending:
        ll, grad, status = evaluate(np.stack([runs[k].x for k in pending]),
                                    np.array([tidx[k] for k in pending], dtype=np.int32))
        rounds += 1
        evaluations += len(pending)
        nxt = []
        for n, k in enumerate(pending):
            if status[n] != 0:                       # the reference's LinAlgError: (last theta that evaluated, 9999)
                failed[k] = True
                continue
            last_good[k] = runs[k].x.copy()
            runs[k].supply(ll[n], grad[n])
            if runs[k].advance():
                nxt.append(k)
        pending = nxt
    results = []
    for k, r in enumerate(runs):
        results.append((last_good[k], 9999) if failed[k] else (r.x, float(r.f)))
        if verbose:
            print("L-BFGS-B %d: cost %e after %d evaluations" % (k, results[-1][1], r.nfev))
    return results, {"rounds": rounds, "evaluations": evaluations, "driver": "reverse-communication"}


def _reverse_communication_ok():
    """True if driving scipy's private ``setulb`` reproduces the public ``fmin_l_bfgs_b`` exactly (checked once)."""
    if _RC_STATE["checked"]:
        return _RC_STATE["ok"]
    _RC_STATE["checked"] = True
    try:
        from scipy.optimize import fmin_l_bfgs_b
        scale = np.array([1.0, 7.0, 0.3, 40.0])

        def fg(x):
            return float(0.5 * np.sum(scale * x * x) + np.sum(np.cos(x)) + 0.1 * (x[0] - x[1] ** 2) ** 2), \
                scale * x - np.sin(x) + 0.2 * (x[0] - x[1] ** 2) * np.array([1.0, -2.0 * x[1], 0.0, 0.0])

        def evaluate(thetas, tidx):
            vals = [fg(t) for t in thetas]
            return (np.array([v[0] for v in vals]), np.array([v[1] for v in vals]), np.zeros(len(vals), dtype=np.int32))
        x0s = [np.array([1.5, -0.7, 2.0, 0.1]), np.array([-2.0, 2.2, 0.4, -1.0])]
        mine, _ = _minimise_reverse_communication(evaluate, [(0, x) for x in x0s])
        ok = True
        for x0, got in zip(x0s, mine):
            ref = fmin_l_bfgs_b(fg, x0, factr=0.1, pgtol=1e-20)
            ok = ok and np.array_equal(ref[0], got[0]) and ref[1] == got[1]
        _RC_STATE["ok"] = bool(ok)
    except Exception:
        _RC_STATE["ok"] = False
    return _RC_STATE["ok"]


def minimise_batched(evaluate, starts, verbose=False, driver=None):
    """Run one L-BFGS-B descent per ``(target_index, theta0)`` in ``starts``, evaluations batched across descents.

    ``driver``: ``"reverse-communication"`` (one thread drives scipy's ``setulb`` for every descent), ``"threads"`` (one
    ``fmin_l_bfgs_b`` per thread, public API only) or None: reverse communication when the installed scipy passes the
    self-test, threads otherwise.  Both give every descent exactly the iterates scipy would compute from the values
    ``evaluate`` returns.  See ``_minimise_threads`` for the return value.
    """
    if driver is None:
        driver = "reverse-communication" if _reverse_communication_ok() else "threads"
    if driver == "reverse-communication":
        return _minimise_reverse_communication(evaluate, starts, verbose)
    if driver != "threads":
        raise ValueError("driver must be 'reverse-communication', 'threads' or None")
    return _minimise_threads(evaluate, starts, verbose)


def _minimise_threads(evaluate, starts, verbose=False):
    """Run one L-BFGS-B descent per ``(target_index, theta0)`` in ``starts``, evaluations batched across descents.

    ``evaluate(thetas (B, G), target_index (B,)) -> (loglik, grad, status)``, e.g. ``DeviceTrainer.evaluate``.
    Returns a list of ``(theta, cost)`` in the order of ``starts`` -- what ``GaussianProcess._learn`` returns per start
    (reference GaussianProcess.py:141-181), including its failure convention: a descent whose covariance stops being
    positive definite ends with ``(last theta, 9999)`` (:174-179).  The second return value carries round / evaluation
    counts.
    """
    from scipy.optimize import fmin_l_bfgs_b
    n = len(starts)
    rv = _Rendezvous(evaluate, n)
    results = [None] * n
    errors = []

    def descent(wid, tidx, theta0):
        last = [np.array(theta0, dtype=np.float64)]

        def fg(theta):
            ll, g, st = rv.call(wid, tidx, theta)
            if st != 0:
                raise np.linalg.LinAlgError("covariance matrix not positive definite")
            last[0] = np.array(theta)         # the reference's current_theta: the last theta that evaluated (:93)
            return ll, g

        try:
            res = fmin_l_bfgs_b(fg, np.array(theta0, dtype=np.float64), factr=0.1, pgtol=1e-20)
            results[wid] = (res[0], float(res[1]))
            if verbose:
                print("L-BFGS-B %d: cost %e after %d evaluations (%s)" % (wid, res[1], res[2]["funcalls"], res[2]["task"]))
        except np.linalg.LinAlgError:
            results[wid] = (last[0], 9999)
        except BaseException as exc:
            errors.append(exc)
        finally:
            rv.leave()

    threads = [threading.Thread(target=descent, args=(i, s[0], s[1]), daemon=True) for i, s in enumerate(starts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results, {"rounds": rv.rounds, "evaluations": rv.evaluations, "driver": "threads"}


def fit_bank(inputs, targets, n_tries=5, device=0, verbose=False):
    """Fit E GPs that share ``inputs`` (M, D) -- one per row of ``targets`` (E, M) -- in ONE batch of E x n_tries descents.

    The per-band pattern of the reference (tests/test_perband_emulator.py:22-37: a loop of ``GaussianProcess(y_train,
    band).learn_hyperparameters(n_tries=...)``) with every round of cost + gradient evaluations served by one GPU launch.
    Random starts are drawn as E successive ``learn_hyperparameters`` calls would draw them.  Returns the list of fitted
    ``GaussianProcess`` objects (state built by the host ``_set_params`` at the best theta of each) and the driver stats.
    """
    from .gaussian_process import GaussianProcess
    inputs = np.asarray(inputs)
    targets = np.atleast_2d(np.asarray(targets, dtype=np.float64))
    E, D = targets.shape[0], inputs.shape[1]
    starts = [(e, th) for e in range(E) for th in 5.0 * (np.random.rand(n_tries, D + 2) - 0.5)]
    trainer = DeviceTrainer(inputs, targets, device=device)
    try:
        fits, stats = minimise_batched(trainer.evaluate, starts, verbose=verbose)
    finally:
        trainer.close()
    gps = []
    for e in range(E):
        mine = fits[e * n_tries:(e + 1) * n_tries]
        best = int(np.argsort(np.array([f[1] for f in mine]))[0])
        gp = GaussianProcess(inputs, targets[e], device=device)
        gp._set_params(np.array(mine[best][0]))
        gp.fit_cost = mine[best][1]
        gps.append(gp)
    return gps, stats
