"""Small-N latency of the drop-in calls (the reference's MultivariateEmulator.predict is one point per call)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
from tests.conftest import golden
from tests.test_gpu_parity import _write_prosail_dump

def bench(fn, n=200):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6

inputs, theta, invQ, invQt, testing = orc.make_S_model(250, 10, 1000, seed=0)
gp = g.GaussianProcess(inputs, []); gp.theta, gp.invQ, gp.invQt = theta, invQ, invQt
for n in (1, 16, 1000):
    t = testing[:n]
    print("GaussianProcess.predict N=%-5d mu+var+grad %8.1f us   mu+grad %8.1f us   hessian %8.1f us" % (
        n, bench(lambda: gp.predict(t)), bench(lambda: gp.predict(t, do_unc=False)), bench(lambda: gp.hessian(t))))
gq = golden("P")
mv = g.MultivariateEmulator(dump=_write_prosail_dump(gq))
y = gq["points"][0]
print("MultivariateEmulator.predict (12 PCs, 2101 wavelengths) 1 point: %.1f us with Jacobian, %.1f us without" % (
    bench(lambda: mv.predict(y)), bench(lambda: mv.predict(y, do_deriv=False))))
models = [(e.inputs, e.theta, e.invQ, e.invQt) for e in mv.emulators]
t0 = time.perf_counter()
for _ in range(20): orc.mv_predict_point(models, mv.basis_functions, y)
print("numpy reference semantics, same call: %.1f us" % ((time.perf_counter() - t0) / 20 * 1e6))
