"""Where the time of a one-point MultivariateEmulator.predict goes (12 PCs x 2101 wavelengths, Jacobian included)."""
import os, sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from gp_emulator_b200 import _lib
from gp_emulator_b200._lib import addr
from tests.conftest import golden
from tests.test_gpu_parity import _write_prosail_dump
gq = golden("P")
mv = g.MultivariateEmulator(dump=_write_prosail_dump(gq))
y = gq["points"][0]
def bench(fn, n=300):
    for _ in range(30): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6
bank = mv._device_bank()
y2 = np.atleast_2d(y)
print("mv.predict(y)                         %.1f us" % bench(lambda: mv.predict(y)))
print("mv.predict(y, do_deriv=False)         %.1f us" % bench(lambda: mv.predict(y, do_deriv=False)))
print("mv._device_bank() (state check)       %.1f us" % bench(lambda: mv._device_bank()))
print("bank.forward(y2)                      %.1f us" % bench(lambda: bank.forward(y2)))
print("bank.forward(y2, want_deriv=False)    %.1f us" % bench(lambda: bank.forward(y2, want_deriv=False)))
lib = _lib.load()
fwd = np.empty((1, bank.W)); dfull = np.empty((1, bank.D, bank.W))
FL = 0x10 | 0x20 | 0x100
print("C call, fwd + Jacobian, numpy outs    %.1f us" % bench(lambda: lib.gpe_bank_predict_ex(bank._h, addr(y2), 1, None, None, None, None, addr(fwd), addr(dfull), FL, None)))
print("C call, fwd only                      %.1f us" % bench(lambda: lib.gpe_bank_predict_ex(bank._h, addr(y2), 1, None, None, None, None, addr(fwd), None, 0x10 | 0x100, None)))
print("np.empty((1,10,2101)) + np.empty((1,2101)) %.1f us" % bench(lambda: (np.empty((1, 10, 2101)), np.empty((1, 2101)))))
