"""Latency of small host calls against the mapped-buffer threshold: GPE_ZERO_COPY_MAX=<points> python tools/zero_copy_probe.py"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc
inputs, theta, invQ, invQt, testing = orc.make_S_model(250, 10, 20000, seed=0)
gp = g.GaussianProcess(inputs, []); gp.theta, gp.invQ, gp.invQt = theta, invQ, invQt
def bench(fn, n=200):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6
print("ZC_MAX", os.environ.get("GPE_ZERO_COPY_MAX"), " ".join("N=%d: %.1f/%.1f us" % (n, bench(lambda: gp.predict(testing[:n], pinned=False)), bench(lambda: gp.predict(testing[:n], do_unc=False, pinned=False))) for n in (64, 256, 1000, 4000, 16000)))
