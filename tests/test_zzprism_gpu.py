"""GPU: one ParOperator over two element geometry types -- hexahedra through the sum-factorised kernel, prisms through the
dense-basis (DMMA) kernel -- on a conforming mixed mesh, as BilinearForm::PartialAssemble builds one sub-operator per geometry
type (/root/reference/palace/fem/bilinearform.cpp:56-101). Against the oracle's assembled sum; a Jacobi-PCG solve of the mixed
system. Mesh, order-1 prism tables and numbering from palace_b200/host/prism.py (checked on the CPU in tests/test_prism_cpu.py)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import pyoracle as O
from oracle import solvers as S
from palace_b200.host import hexspace as hs
from palace_b200.host import prism as pr
from tests import common

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def mixed(b2p_ctx):
    from palace_b200 import capi

    mesh = pr.mixed_box_mesh((5, 3, 2), 2, n_attr=3)   # 12 hexahedra + 36 prisms (four full batches of 8 and a ragged one)
    sp_ = pr.build_mixed_nd_space(mesh)
    kind = O.CURLCURL_MASS
    blob = common.coefficient(kind, 3, "matrix", a_mass=1.3, a_curl=0.7)
    nodes = hs.gauss_lobatto(2)
    xe = mesh.hexes.node_coords(1, nodes)
    qx, qw = hs.gauss_legendre(2)
    nB, nG = hs.lagrange_table(nodes, qx)
    gh = capi.Geom.hex(b2p_ctx, xe, mesh.hexes.attr, 1, 2, nB, nG, qw)
    t = hs.tables_1d(1, 2)
    idx_h, ori_h = sp_.hex_space.native_restriction()
    oph = capi.Op.create(b2p_ctx, gh, kind, 1, sp_.ndofs, idx_h, ori_h, sp_.hex_space.dof_map, t.Bo, t.Bc, t.Gc, blob, False)
    pts, w = pr.prism_quadrature()
    ip, cp_ = pr.prism_tables(pts)
    qd_p = pr.prism_qdata(mesh, w)
    gp = capi.Geom.general(b2p_ctx, qd_p)
    opp = capi.Op.create_dense(b2p_ctx, gp, kind, sp_.ndofs, sp_.prism_idx, sp_.prism_orient, ip, cp_, blob)
    A = capi.Operator.par(b2p_ctx, sp_.ndofs, sp_.ndofs, [oph, opp], [1.0, 1.0], sp_.ess_dofs, diag_policy=1)
    # oracle: assembled sum of the two blocks
    qd_h = O.geom_hex_qdata(xe, mesh.hexes.attr, 1, 2)
    ih, ch, _ = O.nd_hex_tables(1, 2)
    Ah = S.assemble_sparse(O.element_matrices(kind, ih, ch, ori_h, qd_h, blob, 12), idx_h.astype(np.int64), sp_.ndofs)
    Ap = S.assemble_sparse(O.element_matrices(kind, ip, cp_, sp_.prism_orient, qd_p, blob, 9), sp_.prism_idx.astype(np.int64), sp_.ndofs)
    Ao = S.eliminate((Ah + Ap).tocsr(), sp_.ess_dofs)
    return dict(A=A, Ao=Ao, sp=sp_, opp=opp, Ap=Ap)


def test_prism_block_alone_matches_the_oracle(mixed):
    n = mixed["sp"].ndofs
    x = np.random.default_rng(0).random(n) - 0.5
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    mixed["opp"].apply(_dev(x), y)
    assert _rel(y.cpu().numpy(), mixed["Ap"] @ x) < RTOL


def test_mixed_geometry_operator_matches_the_assembled_sum(mixed):
    A, Ao, n = mixed["A"], mixed["Ao"], mixed["sp"].ndofs
    assert not A.is_fused()
    x = np.random.default_rng(1).random(n) - 0.5
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    A.mult(_dev(x), y)
    assert _rel(y.cpu().numpy(), Ao @ x) < RTOL
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    A.assemble_diagonal(d)
    assert _rel(d.cpu().numpy(), Ao.diagonal()) < RTOL


def test_pcg_on_the_mixed_geometry_system(mixed, b2p_ctx):
    from palace_b200 import capi

    A, Ao, sp_ = mixed["A"], mixed["Ao"], mixed["sp"]
    n = sp_.ndofs
    b = np.random.default_rng(2).random(n) - 0.5
    b[sp_.ess_dofs] = 0.0
    pc = capi.Solver.jacobi(b2p_ctx)
    pc.set_operator(A)
    cg = capi.Solver.krylov(b2p_ctx, capi.CG, rel_tol=1e-12, max_it=1000)
    cg.set_operator(A)
    cg.set_preconditioner(pc)
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    cg.mult(_dev(b), x)
    assert cg.stats()["converged"]
    assert _rel(x.cpu().numpy(), spla.spsolve(Ao.tocsc(), b)) < 1e-9
