"""The C-ABI library loads and exports exactly what include/edxraster_c.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from edxraster_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "edxraster_c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(edx_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(_lib.SYMBOLS) == names


def test_version_names_the_architecture():
    assert b"sm_100a" in _lib.load().edx_version()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert _lib.load().edx_create(0, C.byref(h)) == _lib.EDX_ERR_NO_DEVICE
    assert not h.value
    from edxraster_b200 import renderer
    with pytest.raises(_lib.EdxError):
        renderer.Renderer(0)


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (only tests, smoke and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "edxraster_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle,", "") or fn == "scenes.py", fn
    for fn in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", fn)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read()
