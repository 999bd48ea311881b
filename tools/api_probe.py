import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, testing = orc.make_S_model(250, 10, 100000, seed=0)
gp = g.GaussianProcess(inputs, [], device=0); gp.theta, gp.invQ, gp.invQt = theta, invQ, invQt
for _ in range(4): r = gp.predict(testing)
for rep in range(3):
    t0 = time.perf_counter(); r = gp.predict(testing); dt = time.perf_counter() - t0
    print("predict(1e5 pageable): %.3f ms" % (dt * 1e3), flush=True)
t0 = time.perf_counter()
for _ in range(20): dm = gp._device_model(100000)
print("model check: %.1f us" % ((time.perf_counter() - t0) / 20 * 1e6))
dm = gp._device_model()
t0 = time.perf_counter()
for _ in range(5): r = dm.predict(testing)
print("DeviceModel.predict: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
out = {"mu": np.empty(100000), "var": np.empty(100000), "deriv": np.empty((100000, 10))}
t0 = time.perf_counter()
for _ in range(5): r = dm.predict(testing, out=out)
print("DeviceModel.predict(out=warm pageable): %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
