"""The two language facades over the C ABI.
  C++      include/digital_filter.hpp, used exactly like the reference's test/cpp-main.cpp
  Fortran  fortran/digital_filtering.f90 -- no Fortran compiler in the image, so its binding is
           exercised by tests/fortran_abi_mimic.c (same struct image, same by-reference calls)
CPU: both compile and link against libdfb200.so.  GPU: both run on the reference's default plane."""
import os
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "tests", "_bin")
LIBDIR = os.path.join(ROOT, "digital-filtering_b200", "lib")


def _build(dfb):
    os.makedirs(BIN, exist_ok=True)
    common = ["-I" + os.path.join(ROOT, "include"), "-L" + LIBDIR, "-ldfb200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", os.path.join(ROOT, "examples", "cpp-main.cpp")] + common +
                   ["-o", os.path.join(BIN, "cpp-test")], check=True, capture_output=True)
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", os.path.join(ROOT, "tests", "fortran_abi_mimic.c")] + common +
                   ["-lm", "-o", os.path.join(BIN, "fortran_abi_mimic")], check=True, capture_output=True)


def test_facades_compile_and_link(dfb):
    _build(dfb)
    assert os.path.exists(os.path.join(BIN, "cpp-test")) and os.path.exists(os.path.join(BIN, "fortran_abi_mimic"))


def test_fortran_module_mirrors_the_c_struct():
    """field order of `type, bind(C) :: dfb_config_c` == field order of `struct dfb_config`"""
    import re
    h = open(os.path.join(ROOT, "include", "dfb200.h")).read()
    body = h[h.index("typedef struct dfb_config {") + len("typedef struct dfb_config {"):h.index("} dfb_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl or decl.startswith("typedef"):
            continue
        names = re.sub(r"^(const\s+)?(unsigned\s+)?\w+\s*\*?", "", decl, count=1)
        c_fields += [n.strip().lstrip("*").strip() for n in names.split(",") if n.strip()]
    f = open(os.path.join(ROOT, "fortran", "digital_filtering.f90")).read()
    fb = f[f.index("type, bind(C) :: dfb_config_c"):f.index("end type dfb_config_c")]
    f_fields = []
    for ln in fb.splitlines()[1:]:
        if "::" in ln:
            f_fields += [n.strip() for n in ln.split("::")[1].split(",")]
    assert [x.lower() for x in c_fields] == [x.lower() for x in f_fields], (c_fields, f_fields)


@pytest.mark.gpu
def test_cpp_facade_runs_the_reference_example(dfb, O):
    if not O.have_ref():
        pytest.skip("needs the data files under oracle/_ref")
    _build(dfb)
    os.makedirs(O.REF_RUN, exist_ok=True)
    r = subprocess.run([os.path.join(BIN, "cpp-test")], cwd=O.REF_RUN, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Ny=510 Nz=400" in r.stdout
    rms = float(r.stdout.split("=")[-1].split()[0])
    assert 20.0 < rms < 80.0          # sqrt(mean R11) of the M6 DNS profile on this plane is ~50 m/s


@pytest.mark.gpu
def test_fortran_calling_convention(dfb, O):
    if not O.have_ref():
        pytest.skip("needs the data files under oracle/_ref")
    _build(dfb)
    r = subprocess.run([os.path.join(BIN, "fortran_abi_mimic"), O.RST_DAT, O.LINE_DAT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    tok = r.stdout.split()
    assert tok[0] == "OK" and (int(tok[1]), int(tok[2])) == (510, 400) and float(tok[3]) > 1.0
