"""Batched hyper-parameter fitting on the device (the caller on the input side of predict).

The reference fits each GP with ``n_tries`` independent L-BFGS-B descents (gp_emulator/GaussianProcess.py:141-209), and a
``MultivariateEmulator`` does that once per principal component on the SAME training inputs
(gp_emulator/multivariate_gp.py:176-186): n_pcs x n_tries descents, each evaluation a fresh M x M inverse in numpy
(~0.1 s at M = 250, D = 10).  Here the descents stay what they are -- scipy's ``fmin_l_bfgs_b`` with the reference's
settings (``factr=0.1, pgtol=1e-20``), one per start, so every descent follows the reference's trajectory to rounding --
but they run in lockstep threads, and every round of function + gradient requests is served by ONE call of
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


def minimise_batched(evaluate, starts, verbose=False):
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
    return results, {"rounds": rv.rounds, "evaluations": rv.evaluations}
