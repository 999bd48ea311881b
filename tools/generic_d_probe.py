"""Throughput of the generic path (D > 32) next to the compiled path at D = 32."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
for M, D, N in ((250, 32, 1_000_000), (250, 40, 1_000_000), (250, 100, 400_000)):
    inputs, theta, invQ, invQt, tt = orc.make_S_model(M, D, 16, seed=1)
    m = g.DeviceModel(inputs, theta, invQt, invQ)
    t = torch.rand(N, D, dtype=torch.float64, device="cuda")
    for kw, name in ((dict(), "mu+var+grad"), (dict(want_var=False), "mu+grad")):
        m.predict(t, **kw); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m.predict(t, **kw); b.record(); torch.cuda.synchronize()
        print("M=%d D=%d %-12s %.3e points/s" % (M, D, name, N / (a.elapsed_time(b) * 1e-3)), flush=True)
    m.close(); del t
