"""ncu target: the cluster path for a handful of points (k_predict_tiny), 16 points at M = 250, D = 10."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, tt = orc.make_S_model(250, 10, 16, seed=1)
m = g.DeviceModel(inputs, theta, invQt, invQ)
t = torch.from_numpy(tt).cuda()
for _ in range(5): m.predict(t)
torch.cuda.synchronize()
os.environ["X"] = "1"
print("ok")
