"""Fused kernel over the number of training points M (D = 10): points/s and fraction of the DMMA peak (37.1 TFLOP/s)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
D = 10
for M in [int(a) for a in (sys.argv[1:] or "32 64 100 128 160 200 224 250 256".split())]:
    N = int(min(2e7, 4e11 / (M * M)))
    inputs, theta, invQ, invQt, tt = orc.make_S_model(M, D, 16, seed=1)
    m = g.DeviceModel(inputs, theta, invQt, invQ)
    t = torch.rand(N, D, dtype=torch.float64, device="cuda")
    m.predict(t); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m.predict(t); m.predict(t); b.record(); torch.cuda.synchronize()
    pps = 2 * N / (a.elapsed_time(b) * 1e-3)
    F = 2 * M * M + M * (5 * D + 6) + D + 1
    print("skew=%s M=%4d %.3e points/s  %.2f of 37.1 TFLOP/s" % (os.environ.get("GPE_SKEW", "default"), M, pps, pps * F / 37.1e12), flush=True)
    m.close(); del t
