"""Fused Hessian (phase C of the fused kernel) against the oracle and against the direct Hessian kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_emulator_b200 as g
from oracle import gp_oracle as orc


def model(M, D, direct, seed=0, scale=1.0, shift=0.0):
    inputs, theta, invQ, invQt, testing = orc.make_S_model(M, D, 333, seed=seed)
    inputs = inputs * scale + shift
    testing = testing * scale + shift
    if scale != 1.0 or shift != 0.0:
        invQ, invQt = orc.prepare_likelihood(inputs, np.sin(inputs.sum(1)), theta)
    if direct:
        os.environ["GPE_HESS_DIRECT"] = "1"
    else:
        os.environ.pop("GPE_HESS_DIRECT", None)
    return g.DeviceModel(inputs, theta, invQt, invQ), (inputs, theta, invQ, invQt, testing)


for (M, D, scale, shift) in [(250, 10, 1, 0), (1000, 10, 1, 0), (300, 4, 1, 0), (60, 2, 1, 0), (500, 16, 1, 0), (120, 7, 1, 0),
                             (250, 10, 1, 1000.0), (250, 10, 30.0, 0), (250, 10, 300.0, 0), (37, 3, 1, 0)]:
    m, (inputs, theta, invQ, invQt, testing) = model(M, D, False, scale=scale, shift=shift)
    r = m.predict(testing, want_mu=True, want_var=True, want_deriv=True, want_hess=True)
    h_only = m.predict(testing, want_mu=False, want_var=False, want_deriv=False, want_hess=True)["hess"]
    md, _ = model(M, D, True, scale=scale, shift=shift)
    hd = md.predict(testing, want_mu=False, want_var=False, want_deriv=False, want_hess=True)["hess"]
    mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
    ho = orc.hessian(inputs, theta, invQt, testing)
    print("M=%4d D=%2d scale=%g shift=%g: fused(with var) %.1e  fused(hess only) %.1e  direct %.1e | mu %.1e var %.1e deriv %.1e"
          % (M, D, scale, shift, orc.ref_err(r["hess"], ho), orc.ref_err(h_only, ho), orc.ref_err(hd, ho),
             orc.ref_err(r["mu"], mu), orc.var_cond_err(r["var"], var, inputs, theta, invQ, testing) if hasattr(orc, "var_cond_err") and False else orc.ref_err(r["var"], var),
             orc.ref_err(r["deriv"], deriv)), flush=True)


def ev(fn, reps=3):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


N = 4_000_000
for (M, D) in [(250, 10), (1000, 10)]:
    t = torch.rand(N if M < 500 else N // 4, D, dtype=torch.float64, device="cuda")
    n = t.shape[0]
    out = {k: torch.empty((n,) + s, dtype=torch.float64, device="cuda") for k, s in
           [("mu", ()), ("var", ()), ("deriv", (D,)), ("hess", (D, D))]}
    for direct in (False, True):
        m, _ = model(M, D, direct)
        s_all = ev(lambda: m.predict(t, want_var=True, want_deriv=True, want_hess=True, out=out))
        s_h = ev(lambda: m.predict(t, want_mu=False, want_var=False, want_deriv=False, want_hess=True, out=out))
        s_v = ev(lambda: m.predict(t, want_var=True, want_deriv=True, out=out))
        print("M=%d D=%d %s: mu+var+deriv+hess %.3e pts/s | hess only %.3e | mu+var+deriv %.3e"
              % (M, D, "direct" if direct else "fused ", n / s_all, n / s_h, n / s_v), flush=True)
