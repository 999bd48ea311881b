"""Variance path beyond M = 4096 (invQ no longer fits L2): points/s and parity on a few points."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
for M, N in ((4096, 40000), (6000, 40000), (8192, 20000)):
    inputs, theta, invQ, invQt, tt = orc.make_S_model(M, 10, 32, seed=1)
    t0 = time.perf_counter()
    m = g.DeviceModel(inputs, theta, invQt, invQ)
    t_up = time.perf_counter() - t0
    ref = orc.predict(inputs, theta, invQ, invQt, tt)
    o = m.predict(tt)
    t = torch.rand(N, 10, dtype=torch.float64, device="cuda")
    m.predict(t); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m.predict(t); b.record(); torch.cuda.synchronize()
    s = a.elapsed_time(b) * 1e-3
    print("M=%5d upload %.1f s, %.3e pts/s = %.1f TFLOP/s; err mu %.1e var %.1e deriv %.1e" % (
        M, t_up, N / s, N / s * (2.0 * M * M + M * 56) / 1e12, orc.ref_err(o["mu"], ref[0]), orc.ref_err(o["var"], ref[1]),
        orc.ref_err(o["deriv"], ref[2])), flush=True)
    m.close(); del t
