"""TEST INFRASTRUCTURE ONLY. Runs the `-m gpu` parity tests on the CPU against tests/emu/libb2p_emu.so, the
host-thread SIMT emulation build of the SAME kernel sources (see cuda_emu.hpp). Enabled only by
B2P_EMU_TESTS=1 in the environment of a pytest run (tests/conftest.py); the product package never imports this
module and libb2p.so never contains emulation code. Purpose: validate indexing / synchronisation / arithmetic of
new or changed kernels before GPU time is spent on them. A pass here is NOT a GPU parity claim."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("B2P_EMU_LIB") or os.path.join(HERE, "libb2p_emu.so")


def build():
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE], stdout=subprocess.DEVNULL)


def enable():
    import torch

    from palace_b200 import capi

    build()
    capi.LIB_PATH = LIB
    capi._lib = None
    capi._stream = lambda stream=None: C.c_void_p(0)
    # "device" tensors are host tensors: the emulated kernels dereference the same pointers
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()  # a copy, like a real host-to-device transfer

    def strip_device(f):
        def g(*a, **k):
            k.pop("device", None)
            return f(*a, **k)

        return g

    for name in ("full", "empty", "zeros", "ones", "tensor", "rand", "randn", "arange"):
        setattr(torch, name, strip_device(getattr(torch, name)))
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.is_available = lambda: True
    torch.cuda.current_stream = lambda *a, **k: type("S", (), {"cuda_stream": 0})()
