"""gp_emulator_b200 -- B200-native prediction engine behind the gp_emulator Python API.

Same public names as the reference package (reference gp_emulator/__init__.py:1-4): ``GaussianProcess``,
``k_fold_cross_validation``, ``MultivariateEmulator``, ``lhd``, ``EmulatorStorage``; plus the device handles
``DeviceModel`` / ``DeviceBank`` for callers that keep data on the GPU.
"""
from .engine import DeviceBank, DeviceModel, MultiDeviceModel
from .gaussian_process import GaussianProcess, k_fold_cross_validation
from .multivariate import MultivariateEmulator
from .lhd import lhd
from .save_emulators import EmulatorStorage
from ._lib import GpemuError, measure_fp64_peaks

__all__ = ["GaussianProcess", "k_fold_cross_validation", "MultivariateEmulator", "lhd", "EmulatorStorage", "DeviceModel", "DeviceBank", "MultiDeviceModel",
           "GpemuError", "measure_fp64_peaks"]
