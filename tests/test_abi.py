"""CPU-only: the C-ABI library builds, loads and exports every symbol include/*.h declares.
No compute entry point is called (no GPU here); without a device they must fail loudly."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_cuda

HEADER = os.path.join(ROOT, "include", "coexpression_b200.h")


@pytest.fixture(scope="module")
def lib():
    from coexpression_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(getPearson|coex_[a-z_0-9]+)\s*\(", src)) - {"coex_ctx", "coex_params"})


def test_header_symbols_exported_and_bound(lib):
    from coexpression_b200 import _lib
    names = declared_functions()
    assert "getPearson" in names and "coex_network_batch" in names and len(names) >= 20
    exported = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH], text=True)
    exported = {line.split()[-1] for line in exported.splitlines() if " T " in line}
    assert set(names) <= exported, set(names) - exported
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    # nothing but the declared C ABI leaks out of the library
    assert exported == set(names), exported - set(names)


def test_library_is_sm100a_tensor_core_code():
    """SASS evidence that the shipped kernels are Blackwell-native: UTCIMMA (tcgen05.mma kind::i8),
    UTMALDG (TMA), LDTM (tcgen05.ld).  cuobjdump runs without a GPU."""
    from coexpression_b200 import _lib
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", _lib.SO_PATH], text=True, stderr=subprocess.STDOUT)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    for mnemonic in ("UTCIMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_abi_version_and_params_layout(lib):
    from coexpression_b200._lib import CoexParams
    assert lib.coex_abi_version() == 2
    assert ctypes.sizeof(CoexParams) == 48          # 6 ints + 2 doubles + 1 long long, no padding surprises
    assert CoexParams.pval_to_join.offset == 24 and CoexParams.soft_separation.offset == 40


@pytest.mark.skipif(has_cuda(), reason="CPU-only behaviour")
def test_no_cpu_fallback_without_device(lib):
    from coexpression_b200 import _lib
    from coexpression_b200.api import CoexContext
    assert lib.coex_device_count() == 0
    h = ctypes.c_void_p()
    assert lib.coex_create(ctypes.byref(h), 0) != 0
    assert b"cuda" in lib.coex_last_error().lower() or b"device" in lib.coex_last_error().lower()
    with pytest.raises(_lib.CoexError):
        CoexContext()


@pytest.mark.skipif(has_cuda(), reason="CPU-only behaviour")
def test_legacy_symbol_aborts_loudly_without_device():
    """getPearson has no error channel (void, reference coexpression_v2.c:22): it must abort,
    never return garbage.  Run in a child process."""
    from coexpression_b200 import _lib
    code = (
        "import ctypes, numpy as np\n"
        f"lib = ctypes.cdll.LoadLibrary({_lib.SO_PATH!r})\n"
        "X = np.ones((4, 8)); r = np.zeros(2); a = np.array([0, 1], dtype=np.int32); b = np.array([2, 3], dtype=np.int32)\n"
        "lib.getPearson(ctypes.c_void_p(X.ctypes.data), ctypes.c_void_p(r.ctypes.data), ctypes.c_void_p(a.ctypes.data),"
        " ctypes.c_void_p(b.ctypes.data), ctypes.c_int(2), ctypes.c_int(8))\n"
        "print('RETURNED')\n")
    p = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert p.returncode != 0 and "RETURNED" not in p.stdout
    assert "getPearson failed" in p.stderr


def test_options_map_to_reference_semantics():
    import math
    from coexpression_b200.api import Options
    p = Options().to_params(20000)
    assert (p.measure, p.use_specific_p, p.anticorrelation, p.noiter, p.useaverage) == (0, 1, 0, 0, 0)
    assert p.selectionthreshold == 0.0 and p.soft_separation == 100000 and p.pval_to_join == 0.1
    assert Options(nsp=0.0).to_params(100).use_specific_p == 0                       # 1-make-network.py:189-193
    assert Options(correlationmeasure="Pearson").to_params(100).measure == 1
    p = Options(selectionthreshold_is_j=True, pval_to_join=0.01).to_params(100)      # -m, :95-96
    assert p.selectionthreshold == -math.log(0.01, 10)
    with pytest.raises(NameError):                                                   # quirk Q8
        Options(nsp=0.5).to_params(100)


def test_headers_are_plain_c_and_the_libraries_link_from_c(tmp_path, lib):
    """The boundary is a C ABI: both headers compile as C99 (no C++ in the signatures), and a C program linked against the
    two libraries gets the ABI version, a loud failure of coex_create without a device, and the host writer's text."""
    from coexpression_b200 import _lib, hostio
    hostio.build()
    inc = os.path.join(ROOT, "include")
    for h in ("coexpression_b200.h", "coexpression_b200_io.h"):
        subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, h)])
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n'
        '#include "coexpression_b200.h"\n#include "coexpression_b200_io.h"\n'
        'int main(void) {\n'
        '    double v[3] = {0.1, -0.0, 1e16};\n'
        '    char out[128];\n'
        '    long long n = coex_io_format_doubles(v, 3, out, sizeof out);\n'
        '    if (n < 0) return 2;\n'
        '    out[n] = 0;\n'
        '    printf("abi=%d devices=%d\\n%s", coex_abi_version(), coex_device_count(), out);\n'
        '    if (coex_device_count() == 0) {\n'
        '        coex_ctx *c = NULL;\n'
        '        if (coex_create(&c, 0) == 0) return 3;\n'
        '        printf("create failed: %s\\n", strlen(coex_last_error()) ? "with a message" : "silently");\n'
        '    }\n'
        '    return 0;\n}\n')
    csrc = os.path.dirname(_lib.SO_PATH)
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe), "-L", csrc,
                           "-l:coexpression_v2.so", "-l:coex_hostio.so", "-Wl,-rpath," + csrc])
    out = subprocess.check_output([str(exe)], text=True)
    lines = out.split("\n")
    assert lines[0].startswith("abi=2 devices=") and lines[1:4] == ["0.1", "-0.0", "1e+16"]
    if not has_cuda():
        assert lines[4] == "create failed: with a message"
