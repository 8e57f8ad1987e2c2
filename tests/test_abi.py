"""The C-ABI library builds, loads, and exports every symbol include/minimagen_b200.h declares; the Python binding
covers exactly that set; and the product path refuses to run without CUDA (no CPU fallback)."""
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "minimagen_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_all_declared_symbols():
    import ctypes
    from minimagen_b200 import build_ext
    lib_path = build_ext.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    lib.mi_abi_version.restype = ctypes.c_int
    assert lib.mi_abi_version() == 2


def test_binding_covers_header():
    from minimagen_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared()


def test_no_cpu_fallback():
    from minimagen_b200 import _native
    from minimagen_b200.ops import NativeOps
    with pytest.raises(RuntimeError, match="not on a CUDA device"):
        _native.ptr(torch.zeros(4))
    ops = NativeOps()
    with pytest.raises(RuntimeError):
        ops.posemb(torch.zeros(2, dtype=torch.int64), 2, 8, torch.zeros(2, 8))


def test_missing_library_fails_loudly(monkeypatch):
    from minimagen_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libminimagen_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.load()


def test_grad_mode_routes_to_the_training_path(emu):
    """Unet.forward under grad mode runs the autograd path (train_path.py), under no_grad the sampling path."""
    from minimagen_b200.Unet import Unet, BaseTest
    u = Unet(**BaseTest.defaults)
    x = torch.randn(1, 3, 32, 32)
    out = u(x, torch.zeros(1, dtype=torch.long), text_embeds=torch.randn(1, 4, 512))   # grad mode on
    assert out.requires_grad and out.grad_fn is not None
    with torch.no_grad():
        out2 = u(x, torch.zeros(1, dtype=torch.long), text_embeds=torch.randn(1, 4, 512))
    assert not out2.requires_grad
