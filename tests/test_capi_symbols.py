"""CPU: the C-ABI shared library loads and exports every function include/b2p.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(b2p_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from palace_b200 import capi

    lib = capi.lib()
    names = declared_functions()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/b2p.h but not exported: {missing}"


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        return
    from palace_b200 import capi

    try:
        capi.Ctx(0)
    except capi.B2PError as e:
        assert "no CPU fallback" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("context creation must fail loudly without a GPU")
