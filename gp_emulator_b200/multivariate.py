"""Drop-in ``MultivariateEmulator`` (reference gp_emulator/multivariate_gp.py:38-222): PCA-compressed
multivariate output, one GP per principal component, prediction on the B200 engine.

Construction (SVD, npz load/save, per-PC training) is host-side numpy, same npz wire format as the
reference (``X, y, hyperparams, thresh, basis_functions, n_pcs``).  ``predict`` keeps the reference's
single-point call and return shapes, and additionally accepts N > 1 points (the reference cannot:
multivariate_gp.py:216 fails to broadcast), evaluated as one bank launch plus one back-projection.
"""
from __future__ import annotations

import os
import shutil

import numpy as np

from .engine import DeviceBank, resolve_devices
from .gaussian_process import GaussianProcess


class MultivariateEmulator(object):
    def __init__(self, dump=None, X=None, y=None, hyperparams=None, thresh=0.98, n_tries=5, device=0,
                 batched_training=False, basis_functions=None, n_pcs=None):
        """See reference multivariate_gp.py:40-121.  ``X`` (N_train, N_full) model outputs, ``y``
        (N_train, N_params) the parameters that produced them, ``hyperparams`` (N_params + 2, n_pcs).
        ``device``: GPU index, list of indices or ``"all"`` (host batches are then spread over all of them per call)."""
        if basis_functions is not None and n_pcs is None:      # (EmulatorStorage hands a stored basis back in)
            n_pcs = basis_functions.shape[0]
        if dump is not None:
            if X is not None or y is not None:
                raise ValueError("You specified both a dump file and X and y")
            with np.load(dump) as f:
                X = f["X"]
                y = f["y"]
                hyperparams = f["hyperparams"]
                if "thresh" in f:
                    thresh = float(f["thresh"])
                if "basis_functions" in f:
                    basis_functions = f["basis_functions"]
                    n_pcs = int(f["n_pcs"])
            if basis_functions is None:
                # older dumps carry no basis: decompose once and rewrite the file (reference :79-90)
                self.calculate_decomposition(X, thresh)
                basis_functions, n_pcs = self.basis_functions, self.n_pcs
                tmp = os.path.join("/tmp", os.path.basename(dump))
                np.savez_compressed(tmp, X=X, y=y, hyperparams=hyperparams, thresh=thresh,
                                    basis_functions=basis_functions, n_pcs=n_pcs)
                shutil.move(tmp if tmp.endswith(".npz") else tmp + ".npz", dump)
        else:
            if X is None or y is None:
                raise ValueError("Need to specify both X and y")
            assert X.shape[0] == y.shape[0]
            assert X.ndim == 2
            assert y.ndim == 2

        self.X_train = X
        self.y_train = y
        self.thresh = thresh
        self.device = device
        if basis_functions is None:
            self.calculate_decomposition(X, thresh)
            basis_functions, n_pcs = self.basis_functions, self.n_pcs
        self.n_pcs = int(n_pcs)
        self.basis_functions = basis_functions
        if hyperparams is not None:
            assert (y.shape[1] + 2 == hyperparams.shape[0]) and (self.n_pcs == hyperparams.shape[1])
        self._bank, self._bank_key = None, None
        self.train_emulators(X, y, hyperparams=hyperparams, n_tries=n_tries, batched=batched_training)

    def dump_emulator(self, fname):
        """Save for reuse, reference npz format (multivariate_gp.py:124-137)."""
        np.savez_compressed(fname, X=self.X_train, y=self.y_train, hyperparams=self.hyperparams,
                            thresh=self.thresh, basis_functions=self.basis_functions, n_pcs=self.n_pcs)

    def calculate_decomposition(self, X, thresh):
        """PCA by SVD; keep the components whose cumulative singular-value share is <= thresh
        (reference multivariate_gp.py:140-160)."""
        # (the reference asks LAPACK for the full (N_full, N_full) right factor and uses its first min(N_train, N_full)
        # rows; the economy factorisation returns just those rows -- same singular values, vectors equal to the last
        # bit or two -- and is 14x cheaper at 250 x 2101)
        _, s, V = np.linalg.svd(X, full_matrices=False)
        keep = (s.cumsum() / s.sum()) <= thresh
        self.basis_functions = V[: keep.size][keep]
        self.n_pcs = int(np.sum(keep))

    def train_emulators(self, X, y, hyperparams, n_tries=2, batched=False):
        """One GP per PC on the compressed outputs (reference multivariate_gp.py:162-188).

        ``batched=True`` (only meaningful when ``hyperparams`` is None): all ``n_pcs * n_tries`` L-BFGS-B descents run
        in lockstep, their cost + gradient evaluations served by one GPU launch per round (``training.py``) -- the PCs
        share the training inputs ``y``, so they are one batch of (theta, target vector) problems.
        """
        self.emulators = []
        train_data = self.compress(X)
        self.hyperparams = np.zeros((2 + y.shape[1], self.n_pcs))
        gps = [GaussianProcess(np.atleast_2d(y), train_data[i], device=self.device) for i in range(self.n_pcs)]
        if hyperparams is None and batched:
            from .training import DeviceTrainer, minimise_batched
            D = y.shape[1]
            # the same random draws, in the same order, as n_pcs sequential learn_hyperparameters calls
            starts = [(i, th) for i in range(self.n_pcs) for th in 5.0 * (np.random.rand(n_tries, D + 2) - 0.5)]
            trainer = DeviceTrainer(np.atleast_2d(y), train_data, device=resolve_devices(self.device)[0])
            try:
                fits, self.training_stats = minimise_batched(trainer.evaluate, starts)
            finally:
                trainer.close()
            for i, gp in enumerate(gps):
                mine = fits[i * n_tries:(i + 1) * n_tries]
                best = int(np.argsort(np.array([f[1] for f in mine]))[0])
                print("After %d, the minimum cost was %e" % (n_tries, mine[best][1]))
                self.hyperparams[:, i] = mine[best][0]
                gp._set_params(self.hyperparams[:, i])
        else:
            for i, gp in enumerate(gps):
                if hyperparams is None:
                    self.hyperparams[:, i] = gp.learn_hyperparameters(n_tries=n_tries)[1]
                else:
                    self.hyperparams[:, i] = hyperparams[:, i]
                    gp._set_params(hyperparams[:, i])
        self.emulators = gps
        self.invalidate_device()

    def compress(self, X):
        """Project full-rank vectors onto the PC basis (reference multivariate_gp.py:191-193)."""
        return X.dot(self.basis_functions.T).T

    def _device_bank(self):
        """Device copy of the per-PC GPs and the basis, rebuilt when any emulator's state is rebound (each
        ``GaussianProcess`` counts assignments to inputs / theta / invQ / invQt, so ``emulators[i]._set_params(...)``
        is seen) or the basis is replaced.  In-place edits of an emulator's arrays need ``invalidate_device()``."""
        key = (tuple((id(g), g._version) for g in self.emulators), id(self.basis_functions), self.n_pcs, str(self.device))
        if self._bank is None or key != self._bank_key:
            if self._bank is not None:
                self._bank.close()
            gps = self.emulators
            self._bank = DeviceBank(np.atleast_2d(self.y_train), np.stack([g.theta for g in gps]),
                                    np.stack([g.invQt for g in gps]), np.stack([g.invQ for g in gps]),
                                    basis=self.basis_functions[: self.n_pcs], device=self.device)
            self._bank_key = key
        return self._bank

    def invalidate_device(self):
        """Drop the device copy of the emulator bank (the next predict re-uploads it)."""
        if self._bank is not None:
            self._bank.close()
        self._bank, self._bank_key = None, None

    def predict(self, y, do_deriv=True, is_gpu=True):
        """Reconstruct the full output (and its Jacobian) at parameter vector(s) ``y``.

        One point, as in the reference (multivariate_gp.py:195-222): returns ``fwd (N_full,)`` and
        ``deriv (N_params, N_full)``.  N > 1 points (new): ``fwd (N, N_full)``, ``deriv (N, N_params, N_full)``.
        """
        bank = self._device_bank()
        if isinstance(y, np.ndarray) or not hasattr(y, "is_cuda"):
            # host caller: one library call (latency matters: the reference is used one point per call)
            y2 = np.atleast_2d(y)
            res = bank.forward(y2, want_deriv=do_deriv)
            fwd, dfull = res if do_deriv else (res, None)
        else:                                # torch CUDA tensor: everything stays on the device
            y2 = y if y.dim() == 2 else y.reshape(1, -1)
            out = bank.predict(y2, want_var=False, want_deriv=False, project=True, project_deriv=do_deriv)
            fwd, dfull = out["fwd"], out.get("deriv_full")
        if y2.shape[0] == 1:
            return (fwd[0], dfull[0]) if do_deriv else fwd[0]
        return (fwd, dfull) if do_deriv else fwd

    def predict_pcs(self, y, do_unc=True, do_deriv=True):
        """PC-space outputs for N points: dict with mu (N, P), var (N, P), deriv (N, P, D)."""
        return self._device_bank().predict(np.atleast_2d(y), want_var=do_unc, want_deriv=do_deriv)
