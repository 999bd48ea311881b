"""The numpy oracle against outputs frozen from the reference itself (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from oracle import gp_oracle as orc
from tests.conftest import golden

TOL = 1e-13  # same numpy/scipy calls as the reference: agreement to rounding of the BLAS reduction order


def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("tag", ["S250", "S1000", "S37", "S1500"])
def test_oracle_matches_reference_S(tag):
    g = golden(tag)
    model = orc.make_S_model(int(g["M"]), int(g["D"]), int(g["N"]), int(g["seed"]))
    assert _sha(*model) == str(g["input_sha"]), "seeded generator no longer reproduces the golden inputs"
    inputs, theta, invQ, invQt, testing = model
    mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
    assert orc.ref_err(mu, g["mu"]) < TOL
    assert orc.ref_err(var, g["var"]) < TOL
    assert orc.ref_err(deriv, g["deriv"]) < TOL
    mu2, var2, deriv2 = orc.predict(inputs, theta, invQ, invQt, testing, do_unc=False)
    assert var2 is None and np.array_equal(mu, mu2) and np.array_equal(deriv, deriv2)
    if "hess" in g:
        nh = g["hess"].shape[0]
        assert orc.ref_err(orc.hessian(inputs, theta, invQt, testing[:nh]), g["hess"]) < TOL


def test_oracle_chunking_is_exact():
    inputs, theta, invQ, invQt, testing = orc.make_S_model(60, 5, 257, 9)
    a = orc.predict(inputs, theta, invQ, invQt, testing, chunk=10 ** 6)
    b = orc.predict(inputs, theta, invQ, invQt, testing, chunk=50)
    for x, y in zip(a, b):
        assert orc.ref_err(x, y) < 1e-14


def test_oracle_matches_reference_T():
    g = golden("T")
    invQ, invQt = orc.prepare_likelihood(g["inputs"], g["targets"], g["theta"])
    # np.linalg.inv on a cond~1e6 matrix: agreement to a few digits below cond * eps
    assert orc.ref_err(invQ, g["invQ"]) < 1e-7
    assert orc.ref_err(invQt, g["invQt"]) < 1e-7
    mu, var, deriv = orc.predict(g["inputs"], g["theta"], g["invQ"], g["invQt"], g["testing"])
    assert orc.ref_err(mu, g["mu"]) < TOL
    assert orc.ref_err(deriv, g["deriv"]) < TOL
    assert orc.var_cond_err(var, g["var"], g["inputs"], g["theta"], g["invQ"], g["testing"]) < 1e-14
    assert orc.ref_err(orc.hessian(g["inputs"], g["theta"], g["invQt"], g["testing"]), g["hess"]) < TOL


def test_oracle_matches_reference_prosail():
    g = golden("P")
    y, hyp, B = g["y"], g["hyperparams"], g["basis_functions"]
    P = int(g["n_pcs"])
    models = []
    for i in range(P):
        invQ, invQt = orc.prepare_likelihood(y, g["train_data"][i], hyp[:, i])
        assert orc.ref_err(invQt, g["invQt"][i]) < 1e-5   # cond(Q) ~ 3.5e7
        models.append((y, hyp[:, i], invQ, g["invQt"][i]))  # reference alpha, locally inverted Q
    mu, var, deriv = orc.bank_predict(models, g["testing"])
    assert orc.ref_err(mu, g["pc_mu"]) < TOL
    assert orc.ref_err(deriv, g["pc_deriv"]) < TOL
    for i in (0, 5, 11):  # variance: 8-12 digits cancel in the quadratic form -> condition-scaled metric
        assert orc.var_cond_err(var[:, i], g["pc_var"][:, i], y, hyp[:, i], models[i][2], g["testing"]) < 1e-12
    assert orc.ref_err(orc.hessian(y, hyp[:, 0], g["invQt"][0], g["testing"][:8]), g["hess0"]) < TOL
    for k in range(3):
        fwd, d = orc.mv_predict_point(models, B, g["points"][k])
        assert orc.ref_err(fwd, g["fwd"][k]) < TOL
        assert orc.ref_err(d[:, g["wsub"]], g["deriv_sub"][k]) < TOL
    fwd_b, mu_b, _, grad_b, dfull = orc.mv_predict_batch(models, B, g["points"], want_deriv_full=True)
    # batched = one GEMM instead of P rank-1 updates: different summation order over the PCs
    assert orc.ref_err(fwd_b, g["fwd"]) < 1e-11
    assert orc.ref_err(dfull[:, :, g["wsub"]], g["deriv_sub"]) < 1e-11


def test_longdouble_oracle_agrees_on_well_conditioned_inputs():
    inputs, theta, invQ, invQt, testing = orc.make_S_model(40, 4, 25, 2)
    mu, var, deriv = orc.predict(inputs, theta, invQ, invQt, testing)
    lmu, lvar, lderiv, lh = orc.predict_longdouble(inputs, theta, invQ, invQt, testing, do_hess=True)
    assert orc.ref_err(mu, lmu) < 1e-14 and orc.ref_err(var, lvar) < 1e-14 and orc.ref_err(deriv, lderiv) < 1e-14
    assert orc.ref_err(orc.hessian(inputs, theta, invQt, testing), lh) < 1e-13


def test_oracle_matches_reference_perband_bank():
    """BASELINE config 5 pattern: E GPs on shared inputs, each predicted and differentiated twice by the reference's own
    predict / hessian (tests/test_perband_emulator.py:22-47) -- golden_K."""
    g = golden("K")
    E = g["thetas"].shape[0]
    models = [(g["inputs"], g["thetas"][e], g["invQ"][e], g["invQt"][e]) for e in range(E)]
    mu, var, deriv, hess = orc.bank_predict(models, g["testing"], do_hess=True)
    assert orc.ref_err(mu, g["mu"]) < TOL and orc.ref_err(deriv, g["deriv"]) < TOL and orc.ref_err(hess, g["hess"]) < TOL
    for e in range(E):
        assert orc.var_cond_err(var[:, e], g["var"][:, e], g["inputs"], g["thetas"][e], g["invQ"][e], g["testing"]) < 1e-14


def test_oracle_matches_reference_cfg4_20pcs():
    """BASELINE config 4 shape: 2101 wavelengths compressed to 20 PCs by the reference's own constructor -- golden_M20."""
    g = golden("M20")
    y, hyp, B = g["y"], g["hyperparams"], g["basis_functions"]
    assert int(g["n_pcs"]) == 20 and B.shape == (20, 2101)
    models = []
    for i in range(20):
        invQ, invQt = orc.prepare_likelihood(y, g["train_data"][i], hyp[:, i])
        assert orc.ref_err(invQt, g["invQt"][i]) < 1e-6
        models.append((y, hyp[:, i], invQ, g["invQt"][i]))
    for k in range(4):
        fwd, d = orc.mv_predict_point(models, B, g["points"][k])
        assert orc.ref_err(fwd, g["fwd"][k]) < TOL
        assert orc.ref_err(d[:, g["wsub"]], g["deriv_sub"][k]) < TOL
    fwd_b, _, _, _, dfull = orc.mv_predict_batch(models, B, g["points"], want_deriv_full=True)
    assert orc.ref_err(fwd_b, g["fwd"]) < 1e-11 and orc.ref_err(dfull[:, :, g["wsub"]], g["deriv_sub"]) < 1e-11
