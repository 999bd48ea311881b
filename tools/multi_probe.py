"""One-call multi-device fan-out: throughput with host-resident (pinned) buffers on all visible GPUs."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, _ = orc.make_S_model(250, 10, 1, seed=0)
mm = g.MultiDeviceModel(inputs, theta, invQt, invQ)
N = int(float(os.environ.get("N", 4e7)))
t = torch.rand(N, 10, dtype=torch.float64).pin_memory().numpy()
out = {"mu": torch.empty(N, dtype=torch.float64).pin_memory().numpy(), "var": torch.empty(N, dtype=torch.float64).pin_memory().numpy(),
       "deriv": torch.empty(N, 10, dtype=torch.float64).pin_memory().numpy()}
mm.predict(t, out=out)
t0 = time.perf_counter()
for _ in range(3): mm.predict(t, out=out)
s = (time.perf_counter() - t0) / 3
print("devices %s: %.3e points/s host-resident (one call)" % (mm.devices, N / s))
