"""CPU suite: the C-ABI library builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every symbol that
include/dsdtm_gpu.h declares. No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dsdtm_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(dsdtm_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_and_binding_agree(built):
    from dsdtm_b200 import capi
    assert header_functions() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    from dsdtm_b200 import capi
    L = capi.load()
    for name in header_functions():
        assert hasattr(L, name), name
    assert L.dsdtm_abi_version() == 1


def test_library_is_sm100a_and_has_no_cpu_fallback(built):
    from dsdtm_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # the product library must not link or reference the oracle
    nm = subprocess.run(["nm", "-D", capi.lib_path()], capture_output=True, text=True).stdout
    assert "orc_" not in nm


def test_create_fails_loudly_without_device_or_with_bad_args(built):
    from dsdtm_b200 import capi
    L = capi.load()
    cam = capi.Cam(640, 480, 500, 500, 320, 240, 500)
    bad = capi.Params(0, 15, 300, 300, 2, 1)                    # levels = 0
    assert not L.dsdtm_create(0, ctypes.byref(cam), ctypes.byref(bad))
    assert b"bad" in L.dsdtm_create_error()
    import torch
    if not torch.cuda.is_available():
        good = capi.Params(5, 15, 300, 300, 2, 1)
        assert not L.dsdtm_create(0, ctypes.byref(cam), ctypes.byref(good))
        assert b"no CPU fallback" in L.dsdtm_create_error()
        with pytest.raises(capi.DsdtmError):
            capi.Context(dict(width=640, height=480, fx=500, fy=500, cx=320, cy=240, f=500))


def test_product_sources_do_not_touch_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "dsdtm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "dsdtm_oracle" not in txt and "liboracle" not in txt, os.path.join(d, f)
