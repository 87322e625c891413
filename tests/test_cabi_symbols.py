"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

import __graft_entry__ as g


def _declared(header):
    text = open(os.path.join(g.ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ucgb200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    names = _declared("ucgb200.h") + _declared("ucgb200_host.h")
    assert len(names) >= 70
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu(pkg):
    """creating a context without a CUDA device must fail loudly (rc -3), never fall back"""
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = pkg.lib().ucgb200_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    try:
        pkg.Context(0)
    except pkg.UCGError as e:
        assert "no CPU fallback" in str(e) or e.rc != 0
    else:
        raise AssertionError("Context() succeeded without a GPU")


def test_product_does_not_reference_oracle():
    """nothing under the product package may import, link or name the oracle"""
    bad = []
    for root, _, files in os.walk(g.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(root, f), errors="ignore").read()
                if re.search(r"ucg_oracle|oracle_binding|libucg_oracle|orc_[a-z]+\(", text):
                    bad.append(os.path.join(root, f))
    assert not bad, bad
