"""Shared problem builders for the parity tests: caller-side mesh/space data (palace_b200.host), the
oracle's dense tables (oracle.pyoracle) and the conversion of both into C-ABI operator descriptors."""
from __future__ import annotations

import dataclasses

import numpy as np

from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import hexmesh as hm
from palace_b200.host import hexspace as hs


@dataclasses.dataclass
class Problem:
    mesh: hm.HexMesh
    topo: hs.HexTopology
    p: int
    q1d: int
    mesh_order: int
    xe: np.ndarray
    nd: hs.HexSpace
    h1: hs.HexSpace
    tabs: hs.Tables1D
    node_tabs: tuple  # (B, G) of the mesh nodal basis at the quadrature points
    qdata_ref: np.ndarray  # oracle q-data [ne][11][Q]


def make_problem(n=(3, 2, 2), p=2, q1d=None, mesh_order=2, warp=0.04, scramble=3, n_attr=3, size=(1.0, 0.7, 0.9)) -> Problem:
    mesh = hm.box_mesh(n, size, warp_amp=warp, scramble_seed=scramble, n_attr=n_attr)
    return problem_on_mesh(mesh, p, q1d, mesh_order)


def problem_on_mesh(mesh, p=2, q1d=None, mesh_order=2) -> Problem:
    """Spaces, tables and oracle q-data on a given hexahedral mesh (conforming numbering of every entity the mesh has)."""
    topo = hs.build_topology(mesh)
    q1d = p + 1 if q1d is None else q1d
    nodes = hs.gauss_lobatto(mesh_order + 1)
    xe = mesh.node_coords(mesh_order, nodes)
    qx, qw = hs.gauss_legendre(q1d)
    nB, nG = hs.lagrange_table(nodes, qx)
    nd = hs.build_nd_space(mesh, topo, p)
    h1 = hs.build_h1_space(mesh, topo, p)
    qd = O.geom_hex_qdata(xe, mesh.attr, mesh_order, q1d)
    return Problem(mesh, topo, p, q1d, mesh_order, xe, nd, h1, hs.tables_1d(p, q1d), (nB, nG), qd)


def eigsh_above(A, M, k, sigma, extra=6, vectors=False, tol=1e-13):
    """The k eigenvalues of A x = lam M x just above sigma (shift-invert ARPACK, as Palace orders its modes), DETERMINISTICALLY: a
    seeded start vector instead of ARPACK's random one, a few extra modes and a generous Krylov space, so that one copy of a
    (nearly) multiple eigenvalue cannot be missed from one run to the next."""
    import scipy.sparse.linalg as spla

    n = A.shape[0]
    v0 = np.random.default_rng(20260923).standard_normal(n)
    kk = min(k + extra, n - 2)
    r = spla.eigsh(A, k=kk, M=M, sigma=sigma, which="LA", tol=tol, v0=v0, ncv=min(n - 1, max(4 * kk, 60)), return_eigenvectors=vectors)
    if vectors:
        lam, V = r
        o = np.argsort(lam)[:k]
        return lam[o], V[:, o]
    return np.sort(r)[:k]


def coefficient(kind, n_attr, coeff_type="matrix", a_mass=1.0, a_curl=1.0):
    """Coefficient context blob for operator ``kind`` in the style of the reference unit tests."""
    if coeff_type == "const":
        first = cf.coeff_ctx(a=a_mass)
        second = cf.coeff_ctx(a=a_curl)
    else:
        am, mc = cf.test_suite_coefficient(n_attr, "scalar" if coeff_type == "scalar" else "matrix")
        first = cf.coeff_ctx(am, mc, a=a_mass)
        # a different (transposed, rescaled) coefficient for the curl part of pair contexts
        second = cf.coeff_ctx(am, mc[::-1].copy() if len(mc) > 1 else mc, a=a_curl, transpose=True)
    if kind == O.CURLCURL_MASS:
        return cf.coeff_ctx_pair(first, second)
    if kind == O.CURLCURL:
        return second
    return first


def oracle_apply(prob: Problem, kind, ctx_blob, x, space=None, q1d=None):
    """y = A x through the dense reference-style path (oracle)."""
    q1d = prob.q1d if q1d is None else q1d
    if kind == O.H1_DIFFUSION:
        sp = prob.h1 if space is None else space
        _, grad, _ = O.h1_hex_tables(sp.p, q1d)
        idx = sp.lex_gid.astype(np.int32)
        y = np.zeros(sp.ndofs)
        return O.apply_add(kind, None, grad, idx, None, prob.qdata_ref, ctx_blob, np.ascontiguousarray(x), y)
    sp = prob.nd if space is None else space
    interp, curl, _ = O.nd_hex_tables(sp.p, q1d)
    idx, ori = sp.native_restriction()
    y = np.zeros(sp.ndofs)
    return O.apply_add(kind, interp, curl, idx, ori, prob.qdata_ref, ctx_blob, np.ascontiguousarray(x), y)


def oracle_diag(prob: Problem, kind, ctx_blob, space=None, q1d=None):
    q1d = prob.q1d if q1d is None else q1d
    if kind == O.H1_DIFFUSION:
        sp = prob.h1 if space is None else space
        _, grad, _ = O.h1_hex_tables(sp.p, q1d)
        return O.diag_add(kind, None, grad, sp.lex_gid.astype(np.int32), prob.qdata_ref, ctx_blob, np.zeros(sp.ndofs))
    sp = prob.nd if space is None else space
    interp, curl, _ = O.nd_hex_tables(sp.p, q1d)
    idx, _ = sp.native_restriction()
    return O.diag_add(kind, interp, curl, idx, prob.qdata_ref, ctx_blob, np.zeros(sp.ndofs))


def gpu_geom(ctx, prob: Problem):
    from palace_b200 import capi

    nB, nG = prob.node_tabs
    return capi.Geom.hex(ctx, prob.xe, prob.mesh.attr, prob.mesh_order, prob.q1d, nB, nG, prob.tabs.qw)


def gpu_op(ctx, geom, prob: Problem, kind, ctx_blob, assemble=False, space=None):
    """Operator through the C ABI, fed exactly what MFEM/Palace would supply: native-order idx +
    orientation, GetDofMap(), 1-D tables."""
    from palace_b200 import capi

    if kind == O.H1_DIFFUSION:
        sp = prob.h1 if space is None else space
        t = hs.tables_1d(sp.p, prob.q1d)
        return capi.Op.create(ctx, geom, kind, sp.p, sp.ndofs, sp.lex_gid.astype(np.int32), None, None, None, t.Bc, t.Gc,
                              ctx_blob, assemble)
    sp = prob.nd if space is None else space
    t = hs.tables_1d(sp.p, prob.q1d)
    idx, ori = sp.native_restriction()
    return capi.Op.create(ctx, geom, kind, sp.p, sp.ndofs, idx, ori, sp.dof_map, t.Bo, t.Bc, t.Gc, ctx_blob, assemble)


# ------------------------------------------------------------------------------------------------
# Solver-level helpers: oracle sparse matrices and the matching C-ABI operators
# ------------------------------------------------------------------------------------------------


def oracle_matrix(prob: Problem, kind, ctx_blob, space=None, q1d=None, eliminate=True):
    """Assembled sparse matrix of the oracle operator (optionally with essential dofs eliminated the
    ParOperator way)."""
    from oracle import solvers as S

    q1d = prob.q1d if q1d is None else q1d
    if kind == O.H1_DIFFUSION:
        sp_ = prob.h1 if space is None else space
        _, grad, _ = O.h1_hex_tables(sp_.p, q1d)
        Ae = O.element_matrices(kind, None, grad, None, prob.qdata_ref, ctx_blob, sp_.P)
        A = S.assemble_sparse(Ae, sp_.lex_gid, sp_.ndofs)
    else:
        sp_ = prob.nd if space is None else space
        interp, curl, _ = O.nd_hex_tables(sp_.p, q1d)
        idx, ori = sp_.native_restriction()
        Ae = O.element_matrices(kind, interp, curl, ori, prob.qdata_ref, ctx_blob, sp_.P)
        A = S.assemble_sparse(Ae, idx.astype(np.int64), sp_.ndofs)
    return S.eliminate(A, sp_.ess_dofs) if eliminate else A


def oracle_interp(in_space, out_space, I_loc):
    from oracle import solvers as S

    return S.interp_matrix(I_loc, in_space.lex_gid, in_space.lex_sign.astype(float), out_space.lex_gid,
                           out_space.lex_sign.astype(float), in_space.ndofs, out_space.ndofs)


def gpu_par_operator(ctx, geom, prob, kind, ctx_blob, space=None, fine_op=None, assemble=False):
    """ParOperator over one local operator with the space's essential dofs (DIAG_ONE).
    With ``fine_op`` the local operator is p-coarsened from it (shares quadrature and coefficient)."""
    from palace_b200 import capi

    if kind == O.H1_DIFFUSION:
        sp_ = prob.h1 if space is None else space
    else:
        sp_ = prob.nd if space is None else space
    if fine_op is None:
        op = gpu_op(ctx, geom, prob, kind, ctx_blob, assemble, space=sp_)
    else:
        t = hs.tables_1d(sp_.p, prob.q1d)
        if kind == O.H1_DIFFUSION:
            op = fine_op.coarsen(sp_.p, sp_.ndofs, sp_.lex_gid.astype(np.int32), None, None, None, t.Bc, t.Gc)
        else:
            idx, ori = sp_.native_restriction()
            op = fine_op.coarsen(sp_.p, sp_.ndofs, idx, ori, sp_.dof_map, t.Bo, t.Bc, t.Gc)
    A = capi.Operator.par(ctx, sp_.ndofs, sp_.ndofs, [op], None, sp_.ess_dofs, diag_policy=1)
    A.local_op = op
    return A


def gpu_interp(ctx, in_space, out_space, comps):
    from palace_b200 import capi
    from palace_b200.host import assemble as asm

    it = capi.Interp(ctx, asm.space_dict(in_space), asm.space_dict(out_space), comps)
    return capi.Operator.interp(ctx, it)
