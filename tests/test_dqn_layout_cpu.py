"""-m "not gpu": the packed weight layout of the Neural-Q kernels (csrc/rlpt_dqn_layout.h, compiled for the host as it is by tests/dqn_layout_host.cpp) is a gap-free
bijection made of contiguous chunks in the canonical tcgen05 operand form -- for every operand the kernels pack: the forward pass's W2 (304 x 208, N split at 160),
W3 (208 x 304), W4 (144 x 208) and the backward pass's W3^T (304 x 208, split) and W2^T (208 x 304)."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(OUT, exist_ok=True)
    so, src = os.path.join(OUT, "libdqn_layout_host.so"), os.path.join(ROOT, "tests", "dqn_layout_host.cpp")
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in (src, os.path.join(CSRC, "rlpt_dqn_layout.h"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", CSRC, "-o", so, src])
    L = ctypes.CDLL(so)
    L.dqn_layout_check.restype = ctypes.c_long
    return L


@pytest.mark.parametrize("n_pad,k_pad,split", [(304, 208, True), (208, 304, False), (144, 208, False)])
def test_packed_weight_layout_is_a_bijection_of_contiguous_chunks(lib, n_pad, k_pad, split):
    assert lib.dqn_layout_check(lib.dqn_layout_split() if split else 0, n_pad, k_pad) == 0


def test_chunks_are_large_and_fit_the_ring(lib):
    """One bulk copy holds the SM's copy engine >= ~360 cycles whatever its size (scratch/ubench/copy_bw.cu), so every chunk but a layer's last should be
    >= 24 KB, and the largest must fit a ring stage (two stages + both A operands + constants <= 227 KB of shared memory)."""
    sizes = []
    for n_parts, k_pad in (((160, 144), 208), ((208,), 304), ((144,), 208)):
        kc = lib.dqn_layout_kc(k_pad)
        for rows in n_parts:
            ks = [min(kc, k_pad - k0) for k0 in range(0, k_pad, kc)]
            sizes += [rows * kw * 2 for kw in ks]
            assert all(rows * kw * 2 >= 24 * 1024 for kw in ks[:-1]), (rows, ks)
    stage = (max(sizes) + 127) // 128 * 128
    assert 128 * 208 * 2 + 128 * 304 * 2 + 2 * stage + 8 * 1024 <= 227 * 1024
