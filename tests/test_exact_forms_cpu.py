"""-m "not gpu": the cheaper evaluation forms the kernels use are bit-identical to the straightforward ones (tests/exact_forms_host.cpp
runs the product's rlpt_device.cuh functions instantiated for the host against the plain forms written out)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(OUT, exist_ok=True)
    so, src = os.path.join(OUT, "libexact_forms_host.so"), os.path.join(ROOT, "tests", "exact_forms_host.cpp")
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in (src, os.path.join(CSRC, "rlpt_device.cuh"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I", CSRC, "-o", so, src])
    L = ctypes.CDLL(so)
    for f in (L.exact_check_hemisphere, L.exact_check_tri_solve, L.exact_check_zero_contribution):
        f.restype = ctypes.c_long
    L.exact_check_zero_contribution.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_uint]
    return L


def test_square_to_hemisphere_select_form_is_the_branch_form(lib):
    assert lib.exact_check_hemisphere(3_000_000, 600, 1984) == 0


def test_tri_solve_shortcuts_change_no_result(lib):
    out = np.zeros(2, np.float64)
    bad = lib.exact_check_tri_solve(600, 3000, 1984, out.ctypes.data_as(ctypes.c_void_p))
    assert bad == 0 and out[0] == 600 * 3000 * 2 and out[1] > 0.3 * out[0]          # a good share of the rays really hit (edges and vertices included)


def test_zero_contribution_threshold_is_the_division_form(lib):
    # the constant the kernels compare against (rlpt_kernels.cu, zero_contribution) -- read from the source so the two cannot drift apart
    src = open(os.path.join(CSRC, "rlpt_kernels.cu")).read()
    m = re.search(r"zero_contribution\(float lr, float lg, float lb\) \{ return \(lr \+ lg \+ lb\) <= (0x[0-9a-fp.\-]+)f; \}", src)
    assert m, "zero_contribution not found in rlpt_kernels.cu"
    thr = np.float32(float.fromhex(m.group(1)))
    assert lib.exact_check_zero_contribution(thr, np.float32(2.9e-4), np.float32(3.1e-4), 2_000_000, 7) == 0
    # and it is the largest such float
    nxt = np.nextafter(thr, np.float32(1))
    assert np.float32(thr / np.float32(3)) < np.float32(0.0001) and not (np.float32(nxt / np.float32(3)) < np.float32(0.0001))
