"""The C-ABI library loads and exports every symbol include/gsm_b200.h declares; the ctypes structs
match the C layout.  No compute calls (CPU only)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsm_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"GSM_API\s+[\w\s\*]+?\b(gsm_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points(gsm):
    syms = declared_symbols()
    assert set(syms) == set(gsm._lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(gsm):
    lib = gsm._lib.load()
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.gsm_abi_version() == 2


def test_struct_layouts_match_c(gsm, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "gsm_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   "sizeof(gsm_variogram),sizeof(gsm_model),sizeof(gsm_targets),sizeof(gsm_neighbor_opts),"
                   "sizeof(gsm_timings));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    L = gsm._lib
    assert sizes == [ctypes.sizeof(L.Variogram), ctypes.sizeof(L.Model), ctypes.sizeof(L.Targets),
                     ctypes.sizeof(L.NeighborOpts), ctypes.sizeof(L.Timings)]


def test_no_cpu_fallback_without_gpu(gsm):
    """On a box without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gsm.GsmError):
        gsm.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "geostatsmodels.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "gsm_oracle" not in text and "oracle/" not in text, f
