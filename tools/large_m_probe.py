"""Throughput of the M > 1024 variance path (K* scratch + column passes)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
P64 = g.measure_fp64_peaks(0)["dmma_tflops"]
for M in (1000, 1100, 1500, 2048, 3000, 4096):
    D = 10
    inputs, theta, invQ, invQt, testing = orc.make_S_model(M, D, 200, seed=1)
    m = g.DeviceModel(inputs, theta, invQt, invQ)
    o = m.predict(testing)
    mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
    N = 400_000 if M <= 2048 else 100_000
    t = torch.rand(N, D, dtype=torch.float64, device="cuda")
    m.predict(t); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2): m.predict(t)
    b.record(); torch.cuda.synchronize()
    s = a.elapsed_time(b) / 2 * 1e-3
    F = 2 * M * M + M * (5 * D + 6) + D + 1
    print("M=%4d: %.3e pts/s  %.1f TFLOP/s (%.2f of DMMA peak)  invQ stream %.1f TB/s | parity mu %.1e var %.1e deriv %.1e"
          % (M, N / s, N * F / s / 1e12, N * F / s / 1e12 / P64, N / 16 * 8.0 * M * M / s / 1e12,
             orc.ref_err(o["mu"], mu), orc.ref_err(o["var"], var), orc.ref_err(o["deriv"], deriv)), flush=True)
