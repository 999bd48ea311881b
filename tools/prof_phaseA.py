"""ncu target: the fused kernel with the variance output disabled (phase A only) and the mean kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, _ = orc.make_S_model(250, 10, 1, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ)
t = torch.rand(int(float(os.environ.get("N", 1e6))), 10, dtype=torch.float64, device="cuda")
for _ in range(3):
    m.predict(t, want_var=False)
    m.predict(t, want_var=True)
torch.cuda.synchronize()
print("ok")
