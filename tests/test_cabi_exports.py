"""The drop-in boundary loads and exports every symbol include/dfb200.h declares -- CPU only
(no compute call is made: without a GPU dfb_create must fail loudly, never fall back)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "dfb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(dfb):
    L = dfb.lib()
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_image_matches_the_library(dfb):
    c = dfb.dfb_config()
    assert dfb.lib().dfb_config_init(ctypes.byref(c)) == 0
    assert c.struct_bytes == ctypes.sizeof(dfb.dfb_config)      # ctypes image == sizeof(struct dfb_config)
    assert c.device == -1 and c.grid_file_len == -1


def test_no_gpu_means_loud_failure_not_fallback(dfb):
    import torch
    if torch.cuda.is_available():
        return
    try:
        dfb.DIGITAL_FILTER()
    except dfb.DfbError as e:
        assert e.code == dfb.ERR_CUDA and "no CPU fallback" in str(e)
    else:
        raise AssertionError("DIGITAL_FILTER() succeeded without a GPU")


def test_bad_config_is_rejected_with_status(dfb):
    c = dfb.dfb_config()
    h = ctypes.c_void_p()
    assert dfb.lib().dfb_create(ctypes.byref(c), ctypes.byref(h)) == dfb.ERR_ARG     # struct_bytes == 0
    assert b"struct_bytes" in dfb.lib().dfb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "digital-filtering_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                bad = re.findall(r"(?:^\s*(?:import|from)\s+\S*oracle|#\s*include\s+\S*oracle|libdforacle|libdfref|dlopen)", src, flags=re.M)
                assert not bad, (os.path.join(dp, f), bad)
