"""-m "not gpu": the C-ABI library loads on a CPU-only box and exports exactly what include/rlpt.h declares; calls
that need the GPU fail loudly (no fallback). No compute happens here."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "rlpt.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rlpt_[a-z0-9_]+)\s*\(", text)) - {"rlpt_allreduce_fn"})


def test_header_and_binding_list_agree():
    import rlpt
    assert declared_symbols() == sorted(rlpt.SYMBOLS)


def test_library_exports_every_declared_symbol():
    import rlpt
    if not os.path.exists(rlpt.LIB_PATH):
        pytest.skip("librlpt.so not built (run __graft_entry__.build())")
    lib = rlpt.lib()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    out = subprocess.check_output(["nm", "-D", "--defined-only", rlpt.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "rlpt_" in l.split()[-1] and not l.split()[-1].startswith("_Z")}
    assert exported == set(declared_symbols())


def test_header_compiles_as_plain_c(tmp_path):
    """the boundary is a C ABI: plain pointers and sizes, no C++ or torch types in the signatures"""
    src = tmp_path / "t.c"
    src.write_text('#include "rlpt.h"\nint main(void) { rlpt_config c; rlpt_stats_t s; (void)c; (void)s; return RLPT_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_config_defaults_are_the_reference_settings():
    """G/constants/*.h: MAX_RAY_BOUNCES 80, AREA_PER_SAMPLE 0.001, MAX_DIST 0.003, INITIAL_RADIANCE 100/144, RADIANCE_THRESHOLD 0.8/144, seed 1984"""
    import rlpt
    if not os.path.exists(rlpt.LIB_PATH):
        pytest.skip("librlpt.so not built")
    c = rlpt.default_config()
    assert (c.width, c.height, c.spp, c.max_bounces, c.seed) == (512, 512, 32, 80, 1984)
    assert abs(c.area_per_sample - 0.001) < 1e-9 and abs(c.max_dist - 0.003) < 1e-9
    assert abs(c.initial_radiance - 100.0 / 144.0) < 1e-6 and abs(c.radiance_threshold - 0.8 / 144.0) < 1e-8
    assert ctypes.sizeof(rlpt.Config) == 13 * 4 and ctypes.sizeof(rlpt.Stats) == 23 * 8


def test_no_cpu_fallback():
    """without a GPU the context cannot be created and says so; nothing computes on the host"""
    import torch
    import rlpt
    if not os.path.exists(rlpt.LIB_PATH):
        pytest.skip("librlpt.so not built")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rlpt.RlptError, match="no CUDA device|CUDA"):
        rlpt.Context(0)
