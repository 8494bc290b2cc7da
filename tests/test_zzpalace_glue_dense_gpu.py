"""GPU: the non-tensor Palace-side glue of include/b2p_palace.hpp (GatherDenseNDSpace / CreateGeneralGeometry /
CreateDenseNDIntegrator; fem/libceed/basis.cpp:40-85, restriction.cpp:207-385, fem/mesh.cpp:146-209) executed against libb2p.so on
the mock-MFEM mesh of tests/mock_mfem; the CPU twin (emulation build) is tests/test_palace_glue.py."""
import os

import pytest

from tests import test_palace_glue as g

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["dense", "dense_co"])
def test_dense_element_glue_executes_on_the_gpu(tmp_path, mode):
    if os.environ.get("B2P_EMU_TESTS") == "1":
        pytest.skip("emulation run: covered by the CPU test")
    from oracle import pyoracle as O

    O.lib()
    exe, blob = g._build(tmp_path, os.path.join(g.ROOT, "palace_b200"), "libb2p.so")
    g._run(exe, blob, 3, mode)
