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


def test_palace_adapter_header_compiles_against_the_c_abi(tmp_path):
    """include/b2p_palace.hpp (the reference-side binding) must stay type-correct against include/b2p.h: compiled here
    against a mock <mfem.hpp> (tests/mock_mfem), since MFEM itself is not available in this container."""
    import subprocess

    tu = tmp_path / "adapter_tu.cpp"
    tu.write_text('#include "b2p_palace.hpp"\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "mock_mfem"),
                        "-I", os.path.join(ROOT, "include"), str(tu)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
