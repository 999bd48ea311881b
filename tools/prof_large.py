"""ncu target: the M > 1024 variance path (k_predict_mean2<DP,true> + k_var_large) at M = 2048, D = 10."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, _ = orc.make_S_model(2048, 10, 1, seed=0)
m = g.DeviceModel(inputs, theta, invQt, invQ)
t = torch.rand(9472, 10, dtype=torch.float64, device="cuda")   # one sub-batch
for _ in range(3):
    m.predict(t)
torch.cuda.synchronize()
print("ok")
