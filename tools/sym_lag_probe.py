import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
def rate(fn, n, reps=3):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return n * reps / (a.elapsed_time(b) * 1e-3)
for M, N in ((250, 20_000_000), (500, 4_000_000), (1000, 2_000_000)):
    inputs, theta, invQ, invQt, tt = orc.make_S_model(M, 10, 500, seed=0)
    t = torch.rand(N, 10, dtype=torch.float64, device="cuda")
    for sym in (False, True):
        m = g.DeviceModel(inputs, theta, invQt, invQ, symmetric_variance=sym)
        print("lag=%s M=%d sym=%d %.3e pts/s" % (os.environ.get("GPE_RING_LAG"), M, sym, rate(lambda: m.predict(t), N)), flush=True)
        m.close()
