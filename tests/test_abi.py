"""The C-ABI library loads and exports every symbol include/lfs_b200.h declares (no compute calls: no GPU here),
the ctypes signature table covers the header, and the product package never touches oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "lfs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(lfs_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if n != "lfs_alloc_fn"))


def test_header_declares_the_expected_surface():
    fns = header_functions()
    for required in ("lfs_projection_ut_3dgs_fused", "lfs_spherical_harmonics_fwd", "lfs_spherical_harmonics_bwd",
                     "lfs_intersect_tile", "lfs_intersect_offset", "lfs_rasterize_to_pixels_from_world_3dgs_fwd",
                     "lfs_rasterize_to_pixels_from_world_3dgs_bwd", "lfs_adam_step", "lfs_adam_step_multi",
                     "lfs_trainer_create", "lfs_trainer_view_forward", "lfs_trainer_view_backward"):
        assert required in fns


def test_library_exports_every_declared_symbol():
    import lichtfeld_studio_b200 as L
    if not os.path.exists(L._lib.LIB_PATH):
        L.build()
    lib = C.CDLL(L._lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/lfs_b200.h but not exported"


def test_ctypes_table_matches_header():
    import lichtfeld_studio_b200 as L
    assert sorted(L._lib.SIGNATURES) == header_functions()
    lib = L.load()
    assert lib.lfs_abi_version() == L._lib.ABI_VERSION == 2
    assert lib.lfs_set_option(b"no_such_option", 1) < 0
    assert b"unknown option" in lib.lfs_last_error()


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "lichtfeld-studio_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                code = "\n".join(l for l in txt.splitlines()
                                 if re.match(r"\s*(from\s+\S+\s+import|import)\s", l) or "#include" in l or "CDLL(" in l)
                assert "oracle" not in code, f"{f} references oracle/ in an import/include/load"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import lichtfeld_studio_b200 as L
    monkeypatch.setattr(L._lib, "_lib", None)
    monkeypatch.setattr(L._lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(FileNotFoundError):
        L._lib.load()


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++-isms, no torch / CUDA types) and a C program
    must link against the library and call a no-compute entry point."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "t.c"
    src.write_text('#include "lfs_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %s", lfs_abi_version(), lfs_last_error()); '
                   'return lfs_set_option("no_such_option", 1) == LFS_ERR_INVALID_ARG ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    libdir = os.path.join(ROOT, "lichtfeld-studio_b200")
    exe = tmp_path / "t"
    r = subprocess.run([cc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-llfs_b200",
                        f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("2 "), (r.returncode, r.stdout, r.stderr)
