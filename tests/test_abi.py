"""CPU-only checks of the drop-in boundary: libldx.so builds, loads, and exports every symbol
include/ldx.h declares; the ctypes table covers them all; without a GPU the product fails loudly
instead of falling back to anything."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    g.build()
    return os.path.join(ROOT, "ld_tools_b200", "lib", "libldx.so")


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ldx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldx_[a-z0-9_]+)\s*\(", text)))


def test_header_is_plain_c():
    src = os.path.join(ROOT, "tests", "_abi_probe.c")
    with open(src, "w") as fh:
        fh.write('#include "ldx.h"\nint main(void){ldx_hit h; ldx_ld_result r; (void)h; (void)r; return sizeof(ldx_hit)==16 ? 0 : 1;}\n')
    try:
        subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", src], check=True)
    finally:
        os.remove(src)


def test_library_exports_every_declared_symbol(lib_path):
    names = declared_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ldx.h but not exported"
    from ld_tools_b200 import _lib
    assert sorted(list(_lib.SIGNATURES) + ["ldx_last_error"]) == names


def test_built_for_sm_100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_numpy_views(lib_path):
    from ld_tools_b200 import _lib
    assert _lib.HIT_DTYPE.itemsize == 16
    assert _lib.LD_RESULT_DTYPE.itemsize == 6 * 8 + 9 * 8 + 8


def test_no_gpu_means_loud_failure(lib_path):
    """On a box without CUDA the product raises; it never computes on the CPU."""
    import ld_tools_b200
    from ld_tools_b200 import _lib
    n = ctypes.c_int32(-1)
    rc = _lib.load().ldx_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ld_tools_b200.LdxError):
        ld_tools_b200.calc_ld([1, 0, 1], [1, 1, 0])
    with pytest.raises(ld_tools_b200.LdxError):
        ld_tools_b200.Context(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under ld_tools_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "ld_tools_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|ldoracle|oracle/|ld_oracle|calc_ld_port", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f"{os.path.join(dirpath, f)} references the oracle"


def test_format_e4_is_python_str(lib_path):
    """ldx_format_e4 (host code; the digit arithmetic the matrix text kernel uses for a cell) against Python's
    str() of the value the result word stands for, for every representable k = round(x, 4) * 10^4."""
    from ld_tools_b200.engine import format_e4
    for k in range(20000):
        assert format_e4(k) == str(k / 10000.0), k
        assert k / 10000.0 == round(k / 10000.0, 4)
