"""C-ABI boundary checks that need no GPU: the library loads, exports every declared symbol, and fails
loudly (status + message, no exit(), no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpemu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpe_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    from gp_emulator_b200 import _lib
    declared = _declared_symbols()
    assert declared == sorted(_lib.SYMBOLS), "include/gpemu.h and gp_emulator_b200/_lib.py disagree"
    for name in declared:
        assert hasattr(lib, name), name


def test_header_constants_match_python_binding():
    """Flag bits and limits are declared twice (C header, ctypes layer): they must agree."""
    from gp_emulator_b200 import _lib
    text = open(os.path.join(ROOT, "include", "gpemu.h")).read()
    defs = {k: int(v, 0) for k, v in re.findall(r"#define\s+(GPE_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)", text)}
    assert defs["GPE_MAX_TRAIN"] == _lib.MAX_TRAIN and defs["GPE_MAX_INPUTS"] == _lib.MAX_INPUTS
    for cname, pyname in [("GPE_WANT_MU", "WANT_MU"), ("GPE_WANT_VAR", "WANT_VAR"), ("GPE_WANT_DERIV", "WANT_DERIV"),
                          ("GPE_WANT_HESS", "WANT_HESS"), ("GPE_HOST_PTRS", "HOST_PTRS"),
                          ("GPE_F32_FAST_TF32", "F32_FAST_TF32"), ("GPE_F32_FORCE_3X", "F32_FORCE_3X"),
                          ("GPE_OPT_SYMMETRIC_VARIANCE", "OPT_SYMMETRIC_VARIANCE")]:
        assert defs[cname] == getattr(_lib, pyname), cname


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "gpemu.h")).read()
    for cite in ("_gpu_predict.cpp:115-159", "GaussianProcess.py:211-251", "predict.cu:168-176",
                 "multivariate_gp.py:216"):
        assert cite in text


def test_version_and_launch_counter(lib):
    assert lib.gpe_version() == 100
    assert lib.gpe_launch_count() >= 0


def test_no_device_is_an_error_not_a_fallback(lib):
    if lib.gpe_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    x = np.zeros((4, 2)); e = np.ones(3); a = np.zeros(4); q = np.zeros((4, 4))
    rc = lib.gpe_model_create(0, 4, 2, x.ctypes.data, e.ctypes.data, a.ctypes.data, q.ctypes.data, C.byref(h))
    assert rc == -3 and not h.value
    assert b"no CUDA device" in lib.gpe_last_error() and b"no CPU fallback" in lib.gpe_last_error()
    from gp_emulator_b200 import GaussianProcess, GpemuError
    gp = GaussianProcess(x, a)
    gp.theta, gp.invQ, gp.invQt = np.zeros(4), q, a
    with pytest.raises(GpemuError):
        gp.predict(np.zeros((3, 2)))


def test_argument_validation(lib):
    h = C.c_void_p()
    x = np.zeros((4, 2)); e = np.ones(3); a = np.zeros(4)
    assert lib.gpe_model_create(0, 0, 2, x.ctypes.data, e.ctypes.data, a.ctypes.data, None, C.byref(h)) == -1
    assert lib.gpe_model_create(0, 4, 33, x.ctypes.data, e.ctypes.data, a.ctypes.data, None, C.byref(h)) == -1
    assert lib.gpe_model_create(0, 4, 2, None, e.ctypes.data, a.ctypes.data, None, C.byref(h)) == -1
    assert lib.gpe_predict(None, None, 1, None, None, None, None, 1, None) == -1
    assert lib.gpe_predict_wrap(e.ctypes.data, x.ctypes.data, a.ctypes.data, x.ctypes.data, x.ctypes.data,
                                a.ctypes.data, a.ctypes.data, a.ctypes.data, 1, 4, 2, 2) == -1  # theta_size < D+1


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gp_emulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src, f"{fn} mentions the oracle"
