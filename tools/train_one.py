#!/usr/bin/env python
"""One batched training evaluation (for ncu): python tools/train_one.py [M] [D] [B]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gp_emulator_b200.training import DeviceTrainer  # noqa: E402

M, D, B = (int(v) for v in (sys.argv[1:4] + ["250", "10", "148"][len(sys.argv) - 1:]))
rs = np.random.RandomState(0)
x = rs.random_sample((M, D))
tr = DeviceTrainer(x, np.sin(x.sum(axis=1)))
th = 5.0 * (rs.random_sample((B, D + 2)) - 0.5)
for _ in range(2):
    ll, g, st = tr.evaluate(th)
print("ok", M, D, B, float(ll[0]), int(st.sum()))
