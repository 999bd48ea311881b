"""Throughput of the mean-only and Hessian kernels (single GP, device-resident)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
M, D = 250, 10
inputs, theta, invQ, invQt, testing = orc.make_S_model(M, D, 500, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ)
h = m.predict(testing, want_mu=True, want_var=False, want_deriv=True, want_hess=True)
mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
print("parity mu %.1e deriv %.1e hess %.1e" % (orc.ref_err(h["mu"], mu), orc.ref_err(h["deriv"], deriv),
      orc.ref_err(h["hess"], orc.hessian(inputs, theta, invQt, testing))))
N = 4_000_000
t = torch.rand(N, D, dtype=torch.float64, device="cuda")
def ev(fn, reps=3):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
s = ev(lambda: m.predict(t, want_var=False))
print("mean+grad: %.3e pts/s  (%.1f TFLOP/s of FP64 ops at 48/pair)" % (N / s, N * M * 48 * 1 / s / 1e12))
out = {"hess": torch.empty(N, D, D, dtype=torch.float64, device="cuda")}
s = ev(lambda: m.predict(t, want_mu=False, want_var=False, want_deriv=False, want_hess=True, out=out))
print("hessian:   %.3e pts/s  (%.1f G FP64 lane-ops/s at 103/pair; peak %.0f)" % (N / s, N * M * 103 / s / 1e9, 148 * 64 * 1.96))
