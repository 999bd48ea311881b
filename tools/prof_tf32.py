"""ncu target for the tcgen05 / TMEM single-precision kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, _ = orc.make_S_model(250, 10, 1, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ)
t = torch.rand(8_000_000, 10, dtype=torch.float32, device="cuda")
for _ in range(3):
    m.predict_f32(t)
torch.cuda.synchronize()
print("ok")
