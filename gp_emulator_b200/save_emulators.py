"""``EmulatorStorage``: a shelve file holding many emulators under tags (host utility).

Same class, methods and on-disk records as the reference (gp_emulator/save_emulators.py:18-105): a record is a plain
dict of numpy arrays -- ``X, y, basis_functions, n_pcs, thresh, hyperparams`` for a ``MultivariateEmulator``,
``input(s), targets, theta`` for a ``GaussianProcess`` -- keyed by ``repr`` of the tag.  Two defects of the reference are
not reproduced: it writes the key ``"input"`` but reads ``"inputs"`` (:57 vs :101), and it tags with ``repr(tag)`` on
write but ``repr(tuple(tag))`` on read (:43 vs :78), so list tags could never be read back.  Here records carry both
spellings and tags are normalised the same way on both sides; files written by the reference load as they are.
"""
from __future__ import annotations

import os
import shelve

from .gaussian_process import GaussianProcess
from .multivariate import MultivariateEmulator


class EmulatorStorage(object):
    def __init__(self, fname):
        self.fname = fname

    @staticmethod
    def _key(tag):
        if isinstance(tag, str):
            return tag
        try:
            return repr(tuple(tag))
        except TypeError:
            return repr(tag)

    def _exists(self):
        return any(os.path.exists(self.fname + ext) for ext in ("", ".db", ".dat", ".dir"))

    def dump_emulator(self, emulator, tag):
        """Store ``emulator`` (scalar or multivariate) under ``tag`` (a string, or any sequence / value)."""
        if not self._exists():
            print("File doesn't exist, creating it")
        if isinstance(emulator, MultivariateEmulator):
            record = {"X": emulator.X_train, "y": emulator.y_train, "basis_functions": emulator.basis_functions,
                      "n_pcs": emulator.n_pcs, "thresh": emulator.thresh, "hyperparams": emulator.hyperparams}
        elif isinstance(emulator, GaussianProcess):
            record = {"input": emulator.inputs, "inputs": emulator.inputs, "targets": emulator.targets,
                      "theta": emulator.theta}
        else:
            raise TypeError("emulator must be a GaussianProcess or a MultivariateEmulator")
        with shelve.open(self.fname) as db:
            db[self._key(tag)] = record

    def get_keys(self):
        if not self._exists():
            raise IOError("File %s doesn't exist!" % self.fname)
        with shelve.open(self.fname) as db:
            return list(db.keys())

    def get_emulator(self, tag, device=0):
        """Rebuild the emulator stored under ``tag``; its device copy is uploaded at the first predict."""
        if not self._exists():
            raise IOError("File %s doesn't exist!" % self.fname)
        with shelve.open(self.fname) as db:
            key = self._key(tag)
            if key not in db and not isinstance(tag, str) and repr(tag) in db:
                key = repr(tag)          # a list tag written by the reference
            record = db[key]
        if "basis_functions" in record:
            return MultivariateEmulator(X=record["X"], y=record["y"], hyperparams=record["hyperparams"],
                                        thresh=record.get("thresh", 0.98), basis_functions=record["basis_functions"],
                                        n_pcs=record.get("n_pcs"), device=device)
        gp = GaussianProcess(record["inputs"] if "inputs" in record else record["input"], record["targets"], device=device)
        gp._set_params(record["theta"])
        return gp
