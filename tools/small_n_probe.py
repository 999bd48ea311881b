"""GPU time of one small launch (device-resident points, CUDA events over 200 back-to-back launches)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, tt = orc.make_S_model(250, 10, 16, seed=1)
m = g.DeviceModel(inputs, theta, invQt, invQ)
for N in (1, 16, 64, 1000):
    t = torch.rand(N, 10, dtype=torch.float64, device="cuda")
    for kw, name in ((dict(), "mu+var+grad"), (dict(want_var=False), "mu+grad")):
        for _ in range(10): m.predict(t, **kw)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200): m.predict(t, **kw)
        b.record(); torch.cuda.synchronize()
        print("N=%4d %-12s %.2f us per launch (plan: %s)" % (N, name, a.elapsed_time(b) * 1e3 / 200, m.plan(N)[:40]), flush=True)
