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
    assert defs["GPE_TRAIN_MAX_M"] == _lib.TRAIN_MAX_M and defs["GPE_TRAIN_MAX_D"] == 32
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
    assert lib.gpe_model_create(0, 4, 257, x.ctypes.data, e.ctypes.data, a.ctypes.data, None, C.byref(h)) == -1
    assert lib.gpe_model_create(0, 4, 2, None, e.ctypes.data, a.ctypes.data, None, C.byref(h)) == -1
    assert lib.gpe_predict(None, None, 1, None, None, None, None, 1, None) == -1
    assert lib.gpe_bank_cost(None, None, 1, None, 0, None, None, None, None) == -1
    assert lib.gpe_predict_wrap(e.ctypes.data, x.ctypes.data, a.ctypes.data, x.ctypes.data, x.ctypes.data,
                                a.ctypes.data, a.ctypes.data, a.ctypes.data, 1, 4, 2, 2) == -1  # theta_size < D+1


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gp_emulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src, f"{fn} mentions the oracle"


def test_header_is_plain_c_and_links(tmp_path):
    """include/gpemu.h is the drop-in boundary for any host language: it must compile as C99 (no C++, no torch types)
    and a C program must link against libgpemu.so and call through it."""
    import shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "gpemu.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    gpe_model* m = 0;\n'
                   '    double x[8] = {0}, e[3] = {1, 1, 1}, a[4] = {0};\n'
                   '    int rc = gpe_model_create(0, 0, 2, x, e, a, 0, &m);   /* M = 0: invalid on any box */\n'
                   '    printf("%d %d %s\\n", gpe_version(), rc, gpe_last_error());\n'
                   '    return rc == GPE_ERR_INVALID ? 0 : 1;\n}\n')
    libdir = os.path.join(ROOT, "gp_emulator_b200")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-L", libdir, "-lgpemu", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("100 -1 ")


def test_training_chunk_pitch_is_bank_conflict_free(tmp_path):
    """The FP64 predict kernels read 8 (or 4) consecutive training rows per quarter-warp with 16-byte loads; the row pitch
    x_pitch(DP) of gpe_math.cuh must put them on disjoint groups of four shared-memory banks for every compiled DP
    (at pitch = DP, D = 8 ran slower than D = 10: DESIGN 4.1c).  Host-only: the constexpr is evaluated by nvcc's host pass."""
    import re
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dps = [int(x) for x in re.search(r"^DPS\s*:=\s*(.*)$", open(os.path.join(root, "Makefile")).read(), re.M).group(1).split()]
    src = tmp_path / "pitch.cu"
    src.write_text('#include <cstdio>\n#include "gpe_math.cuh"\nint main() { const int dps[] = {%s};\n'
                   '  for (int dp : dps) printf("%%d %%d\\n", dp, gpe::x_pitch(dp)); return 0; }\n' % ", ".join(map(str, dps)))
    exe = tmp_path / "pitch"
    subprocess.run(["nvcc", "-std=c++17", "-I", os.path.join(root, "gp_emulator_b200", "csrc"), "-o", str(exe), str(src)],
                   check=True, capture_output=True, timeout=300)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    pitches = dict(zip(map(int, out[0::2]), map(int, out[1::2])))
    assert sorted(pitches) == sorted(dps)
    for dp, xp in pitches.items():
        assert xp >= dp and xp % 2 == 0 and xp - dp <= 2                      # 16-byte aligned rows, at most one pad pair
        groups = {(r * xp * 8 // 16) % 8 for r in range(8)}                   # 16-byte bank group of the first load of row r
        assert len(groups) == 8, (dp, xp)
