"""The Palace-side glue of include/b2p_palace.hpp -- MFEM objects -> b2p descriptors (fem/libceed/basis.cpp:15-85,
restriction.cpp:113-297, fem/mesh.cpp:146-209, fem/bilinearform.cpp:27-107) -- EXECUTED through the C ABI on a two-element
hexahedral mesh held by the mock MFEM of tests/mock_mfem (native ND numbering with -1-d sign flips, H1 nodes in MFEM's vertex
order, column-major DofToQuad tables) and compared with the oracle's dense reference-style apply and diagonal.
CPU: against the SIMT emulation build of the kernel sources (test infrastructure); GPU (-m gpu): against libb2p.so."""
import os
import subprocess

import numpy as np
import pytest

from palace_b200.host import coeff as cf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, libdir, libname):
    exe = str(tmp_path / "glue_exec")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "mock_mfem"),
                           os.path.join(ROOT, "tests", "mock_mfem", "glue_exec.cpp"), "-o", exe, "-L" + libdir, "-l:" + libname,
                           "-L" + os.path.join(ROOT, "oracle"), "-l:liboracle.so", "-Wl,-rpath," + libdir,
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    am, mc = cf.test_suite_coefficient(2, "matrix")
    blob = cf.coeff_ctx_pair(cf.coeff_ctx(am, mc, a=1.3), cf.coeff_ctx(am, mc[::-1].copy(), a=0.7, transpose=True))
    path = str(tmp_path / "coeff.bin")
    with open(path, "wb") as f:
        f.write(np.asarray(blob).tobytes())
    return exe, path


def _run(exe, blob, p, mode="hex"):
    r = subprocess.run([exe, blob, str(p), mode], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GLUE_EXEC OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("p", [1, 2, 3])
def test_glue_executes_on_the_emulation_build(tmp_path, p):
    from oracle import pyoracle as O
    from tests.emu import emu_mode

    O.lib()
    emu_mode.build()
    exe, blob = _build(tmp_path, os.path.join(ROOT, "tests", "emu"), "libb2p_emu.so")
    _run(exe, blob, p)


@pytest.mark.parametrize("mode", ["dense", "dense_co"])
def test_dense_element_glue_executes_on_the_emulation_build(tmp_path, mode):
    """The non-tensor glue (GatherDenseNDSpace / CreateGeneralGeometry / CreateDenseNDIntegrator: FULL DofToQuad tables, element
    transformations for the geometry factors, DofTransformation columns -> int8 tridiagonal rows as restriction.cpp:301-329)."""
    from oracle import pyoracle as O
    from tests.emu import emu_mode

    O.lib()
    emu_mode.build()
    exe, blob = _build(tmp_path, os.path.join(ROOT, "tests", "emu"), "libb2p_emu.so")
    _run(exe, blob, 2, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("p", [2, 3, 4])
def test_glue_executes_on_the_gpu(tmp_path, p):
    if os.environ.get("B2P_EMU_TESTS") == "1":
        pytest.skip("emulation run: covered by the CPU test")
    from oracle import pyoracle as O

    O.lib()
    exe, blob = _build(tmp_path, os.path.join(ROOT, "palace_b200"), "libb2p.so")
    _run(exe, blob, p)
