"""The C-ABI library loads on a CPU-only box and exports exactly what include/dhj.h declares; the
product fails loudly (no CPU fallback) when there is no device.  No compute calls here."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, PKG


@pytest.fixture(scope="module")
def lib_path():
    path = os.path.join(PKG, "lib", "libdhj.so")
    if not os.path.exists(path):
        import __graft_entry__ as G
        G.build()
    return path


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dhj.h")).read()
    return sorted(set(re.findall(r"^DHJ_API\s+(?:const\s+char\*|int)\s+(dhj_\w+)\(", text, flags=re.M)))


def test_header_matches_binding(lib_path):
    import dhj
    assert header_symbols() == sorted(dhj.EXPORTS)
    lib = dhj.load_library()
    for name in dhj.EXPORTS:
        assert hasattr(lib, name), name
    assert lib.dhj_abi_version() == 1


def test_exports_are_exactly_the_header(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("dhj_"))
    assert exported == header_symbols()


def test_sass_is_sm100a(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout and "sm_90" not in out.stdout


def test_header_cites_reference():
    text = open(os.path.join(ROOT, "include", "dhj.h")).read()
    for cite in ("160-192", "lbfgs_calibrator.py:118-177", "synthetic_generator.py:123-138",
                 "_numdiff.py"):
        assert cite in text


def test_no_cpu_fallback(lib_path):
    """Without a GPU every entry into the product raises; nothing silently computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import dhj
    with pytest.raises(dhj.NativeError, match="no CPU fallback"):
        dhj.Context(0)
    missing = dhj.load_library  # loading a non-existent library must fail loudly too
    with pytest.raises(dhj.NativeError, match="not found"):
        missing("/nonexistent/libdhj.so")


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under the product package may reference it."""
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("oracle/", "").lower() or f == "__never__", os.path.join(dirpath, f)


def test_header_is_plain_c(tmp_path):
    """include/dhj.h must be consumable by a C compiler (cgo / JNI / ctypes-style bindings): no C++ in the boundary."""
    src = tmp_path / "use_dhj.c"
    src.write_text('#include "dhj.h"\nint main(void) { dhj_ctx* c = 0; (void)c; return dhj_abi_version() == DHJ_ABI_VERSION ? 0 : 1; }\n')
    lib = os.path.join(PKG, "lib")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{os.path.join(ROOT, 'include')}", str(src),
                    f"-L{lib}", "-ldhj", f"-Wl,-rpath,{lib}", "-o", str(tmp_path / "use_dhj")], check=True)
    assert subprocess.run([str(tmp_path / "use_dhj")]).returncode == 0
