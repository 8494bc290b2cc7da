"""GPU: the known-answer cases of the reference's own unit tests for the vector layer, at one rank
(/root/reference/test/unit/test-vector.cpp "Vector Sum - Real" :17-40, "Vector Sum - Complex" :76-110, "Sqrt function" :316-347;
test-orthog.cpp "OrthogonalizeColumn Parameterized - Real 2" :164-232 for MGS / CGS / CGS2), through the C ABI."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def test_vector_sum_real_and_complex(b2p_ctx):
    from palace_b200 import capi

    rank, size = 0, 1
    v = _dev([rank * 3 + i + 1 for i in range(3)])                       # rank 0: [1, 2, 3]
    assert capi.vec_sum(b2p_ctx, v) == 9.0 * size * (size - 1) / 2.0 + 6.0 * size
    re, im = _dev([rank, rank]), _dev([rank + 0, rank + 1])              # ComplexVector(2): (rank, rank + i)
    assert capi.vec_sum(b2p_ctx, re) == 0.0 and capi.vec_sum(b2p_ctx, im) == 1.0


def test_sqrt_function(b2p_ctx):
    from palace_b200 import capi

    v = _dev([4.0, 9.0, 16.0, 25.0])
    capi.flux_sqrt_scale(b2p_ctx, 4, 1.0, v)
    assert v.cpu().tolist() == [2.0, 3.0, 4.0, 5.0]
    v = _dev([1.0, 4.0, 9.0])
    capi.flux_sqrt_scale(b2p_ctx, 3, 4.0, v)                              # sqrt(4 x)
    assert v.cpu().tolist() == [2.0, 4.0, 6.0]


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_orthogonalize_column_real_2(b2p_ctx, kind):
    from palace_b200 import capi

    V = [_dev([1.0, 0.0, 0.0, 0.0]), _dev([0.0, 1.0, 0.0, 0.0])]
    w = _dev([0.0, 1.0, 0.0, 0.0])
    capi.vec_orthogonalize(b2p_ctx, kind, V[:1], w)                       # exact in double: multiply by zero
    assert w.cpu().tolist() == [0.0, 1.0, 0.0, 0.0]
    w = _dev([0.0, 1.0, 2.0, 3.0])                                        # d_v[i] = mpi_rank + i
    H = capi.vec_orthogonalize(b2p_ctx, kind, V, w)
    wh = w.cpu().numpy()
    assert abs(wh[0]) < 1e-12 and abs(wh[1]) < 1e-12 and wh[2] == 2.0 and wh[3] == 3.0
    assert np.allclose(H, [0.0, 1.0], rtol=0, atol=1e-15)                 # size (size - 1) / 2, size (size + 1) / 2 at size = 1
