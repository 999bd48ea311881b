"""Host-resident predict through the drop-in class on 1 GPU and on all visible GPUs in ONE call (gpe_multi_*), pinned buffers.
GPE_PIPE_TRACE=1 prints the measured link rates and the relay routing; GPE_MULTI_RELAY=off|auto|force selects it."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, _ = orc.make_S_model(250, 10, 1, seed=0)
G = torch.cuda.device_count()
per = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
N = per * G
t = torch.rand(N, 10, dtype=torch.float64).pin_memory().numpy()
out = {k: torch.empty(s, dtype=torch.float64).pin_memory().numpy() for k, s in (("mu", (N,)), ("var", (N,)), ("deriv", (N, 10)))}
for dev in ([0], list(range(G))):
    gp = g.GaussianProcess(inputs, [], device=dev); gp.theta, gp.invQ, gp.invQt = theta, invQ, invQt
    n = per if len(dev) == 1 else N
    o = {k: v[:n] for k, v in out.items()}
    gp.predict(t[:n], out=o)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); gp.predict(t[:n], out=o); best = min(best, time.perf_counter() - t0)
    print("devices", dev, "relay", os.environ.get("GPE_MULTI_RELAY", "auto"), "%.3e points/s" % (n / best), flush=True)
    gp.invalidate_device()
