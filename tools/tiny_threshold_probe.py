"""Where the cluster path (predict_tiny.cuh) stops paying: latency of DeviceModel.predict over N with the threshold forced."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, tt = orc.make_S_model(250, 10, 8000, seed=1)
m = g.DeviceModel(inputs, theta, invQt, invQ)
for N in (288, 500, 1000, 2000, 4000, 7104):
    t = torch.from_numpy(tt[:N]).cuda()
    for _ in range(20): m.predict(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        m.predict(t); torch.cuda.synchronize()
    print("TINY_MAX=%s N=%5d %.1f us per synchronous device call (%s)" % (os.environ.get("GPE_TINY_MAX"), N,
          (time.perf_counter() - t0) / 200 * 1e6, m.plan(N)[:16]), flush=True)
