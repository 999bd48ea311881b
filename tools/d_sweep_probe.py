"""Throughput over the number of inputs D (M = 250): fused mean + variance + gradient and mean + gradient only,
next to the FP64 instruction bound (3D + 14 per pair at 64 lanes/clk/SM; + 2 M^2 + ... for the variance)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
M = 250
sms = torch.cuda.get_device_properties(0).multi_processor_count
for D in [int(a) for a in (sys.argv[1:] or "8 10 12 14 16 20 24 32".split())]:
    N = 2_000_000
    inputs, theta, invQ, invQt, tt = orc.make_S_model(M, D, 16, seed=1)
    m = g.DeviceModel(inputs, theta, invQt, invQ)
    t = torch.rand(N, D, dtype=torch.float64, device="cuda")
    res = []
    for kw in (dict(), dict(want_var=False)):
        m.predict(t, **kw); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m.predict(t, **kw); m.predict(t, **kw); b.record(); torch.cuda.synchronize()
        res.append(2 * N / (a.elapsed_time(b) * 1e-3))
    lanes = 64 * sms * 1.96e9
    bound_mean = lanes / (M * (3 * D + 14))
    bound_full = lanes / (M * (3 * D + 14) + 256 * 256)
    print("D=%2d full %.3e (%.2f of bound)  mean+grad %.3e (%.2f of bound)" % (D, res[0], res[0] / bound_full, res[1], res[1] / bound_mean), flush=True)
    m.close(); del t
