"""The C-ABI library loads and exports every symbol include/srt.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "srt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from simple_raytracer_b200 import tracer
    lib = tracer.load_library()
    names = declared_symbols()
    assert len(names) >= 25 and "srt_render" in names and "srt_upload_scene" in names
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/srt.h but not exported"
    assert lib.srt_abi_version() == 1


def test_python_binding_declares_every_symbol():
    from simple_raytracer_b200 import tracer
    lib = tracer.load_library()
    for n in declared_symbols():
        assert getattr(lib, n).argtypes is not None or n in ("srt_abi_version",), n


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "simple_raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "oracle.c" not in src and "oracle_math" not in src, f


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from simple_raytracer_b200.tracer import SrtError, Tracer
    with pytest.raises(SrtError, match="no CUDA device|CUDA"):
        Tracer(8, 8, np.zeros((2, 2, 4), np.float32))


def test_create_rejects_bad_arguments():
    from simple_raytracer_b200 import tracer
    lib = tracer.load_library()
    h = ctypes.c_void_p()
    sky = np.zeros((2, 2, 4), np.float32)
    assert lib.srt_create(0, 8, sky.ctypes.data_as(ctypes.c_void_p), 2, 2, -1, ctypes.byref(h)) == 1
    assert b"bad image size" in lib.srt_last_error(None)
    assert lib.srt_create(8, 8, None, 2, 2, -1, ctypes.byref(h)) == 1
    assert lib.srt_destroy(None) == 0
