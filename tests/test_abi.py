"""C-ABI boundary: the in-tree shared library loads and exports every symbol declared in
include/*.h; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from poissonrecon_gpu_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("prb.h", "prb_io.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(prb\w*)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_are_exported():
    lib = api.load_library()
    syms = declared_symbols()
    assert len(syms) >= 19, syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported by libprb.so"
    for s in api.EXPORTS:
        assert s in syms


def test_library_is_in_tree():
    assert os.path.dirname(api.lib_path()) == os.path.join(ROOT, "poissonrecon_gpu_b200")
    assert os.path.exists(api.lib_path())


def test_bad_arguments_are_rejected():
    lib = api.load_library()
    h = ctypes.c_void_p()
    assert lib.prb_create(0, 1, ctypes.byref(h)) == -1          # depth < 2
    assert lib.prb_create(0, 13, ctypes.byref(h)) == -1         # depth > 12
    assert b"depth" in lib.prb_last_error()
    assert lib.prb_create(0, 8, None) == -1


def test_no_cpu_fallback():
    """Without a usable sm_100 device prb_create must fail with PRB_ERR_CUDA and a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = api.load_library()
    h = ctypes.c_void_p()
    assert lib.prb_create(0, 8, ctypes.byref(h)) == -2
    assert b"no CPU fallback" in lib.prb_last_error()
    with pytest.raises(api.PrbError):
        api.PoissonRecon(8)


def test_product_does_not_reference_the_oracle():
    """The product path (package, csrc, include, CLI) must not import / link / exec oracle/."""
    pk = os.path.join(ROOT, "poissonrecon_gpu_b200")
    bad = []
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"liborc|orc_run|oracle_binding|oracle/build|/oracle/", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
