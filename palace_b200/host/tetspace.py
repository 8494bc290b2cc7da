"""Caller-side stand-in for what MFEM/Palace supply at the boundary for TETRAHEDRAL meshes: conforming tet
meshes, the Nedelec (first kind) tetrahedron of order p as FULL DofToQuad tables in native dof order, the
element restriction with its tridiagonal "curl-oriented" transformation, and geometry q-data. This is the
input of the dense-basis operator path (b2p_op_create_dense), i.e. what
/root/reference/palace/fem/libceed/basis.cpp:40-85 (InitNonTensorBasis) and restriction.cpp:207-385
(InitNativeRestr, curl_orients) hand to libCEED for every simplex.

The element follows the published definition of MFEM's ND_TetrahedronElement (open basis = Gauss-Legendre,
/root/reference/palace/fem/multigrid.hpp:35): dofs are tangential components at p Gauss-Legendre points per
edge (edge order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3), tangent from the first to the second vertex), two
tangential components per face point (faces (1,2,3),(0,3,2),(0,1,3),(0,2,1), tangents v1-v0 and v2-v0 of the
face's vertex order, the pair adjacent in the dof order) and three Cartesian components per interior point;
shape functions are the dual basis of these functionals in the first-kind space P_{p-1}^3 + S_p. MFEM itself
is not available in the build container, so the restatement is pinned by identities (tests/test_tet_cpu.py):
unisolvence, exact reproduction of P_{p-1}^3 fields on randomly ordered/oriented tets (which fails for any
wrong node position, tangent, sign or face-pair transformation), energies against direct integration, and
curl(grad) = 0.

Where MFEM resolves a face shared by two differently ordered tets through its DofTransformation (2x2
blocks on the face-dof pairs, restriction.cpp:301-329 stores them as int8 tridiagonal rows), this module
derives the same kind of transformation from first principles: a face's global dofs are defined on its
vertices sorted by global id; a local face with another vertex order sees tangents that are small-integer
combinations of the global ones.
"""
from __future__ import annotations

import dataclasses
import itertools

import numpy as np
from numpy.polynomial import legendre as npleg

TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
TET_FACES = ((1, 2, 3), (0, 3, 2), (0, 1, 3), (0, 2, 1))
_REF_VERTS = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])


def nd_tet_ndof(p: int) -> int:
    return p * (p + 2) * (p + 3) // 2


def _gl01(n):
    """n Gauss-Legendre points on [0, 1] (MFEM poly1d.OpenPoints(n - 1, GaussLegendre))."""
    if n <= 0:
        return np.zeros(0)
    x, _ = npleg.leggauss(n)
    return 0.5 * (x + 1.0)


# ------------------------------------------------------------------------------------------------
# Quadrature: collapsed Gauss-Jacobi (Stroud conical product), exact to the requested degree.
# (At the boundary the rule is caller data: Palace passes MFEM's qX / qW, basis.cpp:51-62.)
# ------------------------------------------------------------------------------------------------


def tet_quadrature(degree: int):
    from scipy.special import roots_jacobi

    n = degree // 2 + 1
    x1, w1 = roots_jacobi(n, 2.0, 0.0)
    x2, w2 = roots_jacobi(n, 1.0, 0.0)
    x3, w3 = roots_jacobi(n, 0.0, 0.0)
    a, wa = 0.5 * (x1 + 1), w1 / 8.0   # weight (1-a)^2
    b, wb = 0.5 * (x2 + 1), w2 / 4.0   # weight (1-b)
    c, wc = 0.5 * (x3 + 1), w3 / 2.0
    A, B, Cc = np.meshgrid(a, b, c, indexing="ij")
    W = wa[:, None, None] * wb[None, :, None] * wc[None, None, :]
    X = A
    Y = B * (1 - A)
    Z = Cc * (1 - A) * (1 - B)
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    return pts, W.ravel()


def tet_quadrature_symmetric6():
    """The 24-point, degree-6 symmetric rule on the reference tetrahedron (Keast 1986; orbits 4 + 4 + 4 + 12, positive weights):
    the rule MFEM's IntRules returns for a tetrahedron at order 6, i.e. what the reference integrates order-3 forms with
    (q_order = 2 p, /root/reference/palace/fem/integrator.cpp:14-22). Exact to 3e-17 on every monomial of degree <= 6
    (tests/test_tet_cpu.py); with curved elements the integrand is not a polynomial and the choice of rule shows at the 1e-8 level,
    which is why the pin against the reference's stored capacitances uses this rule (tests/test_spheres_golden.py)."""
    import itertools

    pts, w = [], []
    for a, wt in ((0.21460287125915202, 0.0066537917096945820), (0.040673958534611353, 0.0016795351758867738),
                  (0.32233789014227551, 0.0092261969239424536)):
        b = 1.0 - 3.0 * a
        pts += [(a, a, a), (a, a, b), (a, b, a), (b, a, a)]
        w += [wt] * 4
    a, b = 0.063661001875017525, 0.26967233145831580
    orbit = sorted(set(itertools.permutations((a, a, b, 1.0 - 2.0 * a - b))))
    pts += [q[:3] for q in orbit]
    w += [0.0080357142857142857] * len(orbit)
    return np.array(pts), np.array(w)


def tet_quadrature_symmetric8():
    """A 43-point, degree-8 symmetric rule on the reference tetrahedron (orbits 4 + 4 + 12 + 6 + 12 + 4 + centroid, the centroid
    weight negative: the structure of Keast's degree-8 rule). The constants solve the moment equations up to degree 8 for this orbit
    structure (residual 3e-17; exactness is tested in tests/test_tet_cpu.py). Kept as a second degree-8 rule beside the conical
    tet_quadrature(8): error-type quantities on CURVED elements move at the 1e-4 ... 1e-3 level between rules of one degree
    (tests/test_cylinder_tet_indicator_golden.py measures exactly that against the reference's stored indicators), eigenvalues at 1e-8.
    Whether MFEM's order-8 tetrahedron rule is this one cannot be checked here (MFEM is not in /root/reference)."""
    import itertools

    pts, w = [], []

    def add(orbit, wt):
        pts.extend(orbit)
        w.extend([wt] * len(orbit))

    o4 = lambda a: [(a, a, a), (a, a, 1 - 3 * a), (a, 1 - 3 * a, a), (1 - 3 * a, a, a)]
    o6 = lambda a: [q[:3] for q in sorted(set(itertools.permutations((a, a, 0.5 - a, 0.5 - a))))]
    o12 = lambda a, b: [q[:3] for q in sorted(set(itertools.permutations((a, a, b, 1 - 2 * a - b))))]
    add(o4(0.005781950502546093), 0.00016983410907009535)
    add(o4(0.08210358830838656), 0.001967033313071645)
    add(o12(0.03660774955326096, 0.1904860419344086), 0.002140519141167885)
    add(o6(0.050532740018802424), 0.0045796838244578746)
    add(o12(0.22906653611675284, 0.0356395827889085), 0.005704485808728209)
    add(o4(0.2068299316088912), 0.014250305821778431)
    add([(0.25, 0.25, 0.25)], -0.020500188654514428)
    return np.array(pts), np.array(w)


# ------------------------------------------------------------------------------------------------
# Reference element: dof functionals and shape functions
# ------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class NDTetElement:
    p: int
    P: int
    nodes: np.ndarray      # [P][3] reference points of the dof functionals
    tangents: np.ndarray   # [P][3] reference tangent of each functional
    coef: np.ndarray       # [P][M] shape function d = sum_m coef[d][m] * phi_m
    edge_off: int = 0
    face_off: int = 0
    int_off: int = 0
    cond: float = 0.0

    def tabulate(self, pts):
        """interp[3][n][P], curl[3][n][P]: reference shape functions / reference curls at pts."""
        Phi, Curl = _modal_basis(self.p, np.asarray(pts, dtype=np.float64))
        interp = np.einsum("dm,mnc->cnd", self.coef, Phi)
        curl = np.einsum("dm,mnc->cnd", self.coef, Curl)
        return np.ascontiguousarray(interp), np.ascontiguousarray(curl)


def face_node_index(p):
    """(i, j) -> position of the face point in the (j outer, i inner) enumeration, i + j <= p - 2."""
    out, o = {}, 0
    for j in range(p - 1):
        for i in range(p - 1 - j):
            out[(i, j)] = o
            o += 1
    return out


def _dof_functionals(p):
    eop, fop, iop = _gl01(p), _gl01(p - 1), _gl01(p - 2)
    V = _REF_VERTS
    nodes, tans = [], []
    for (a, b) in TET_EDGES:
        for i in range(p):
            nodes.append(V[a] + eop[i] * (V[b] - V[a]))
            tans.append(V[b] - V[a])
    face_off = len(nodes)
    pm2 = p - 2
    for (a, b, c) in TET_FACES:
        for j in range(p - 1):
            for i in range(p - 1 - j):
                w = fop[i] + fop[j] + fop[pm2 - i - j]
                x = (fop[pm2 - i - j] * V[a] + fop[i] * V[b] + fop[j] * V[c]) / w
                nodes += [x, x]
                tans += [V[b] - V[a], V[c] - V[a]]
    int_off = len(nodes)
    pm3 = p - 3
    for k in range(p - 2):
        for j in range(p - 2 - k):
            for i in range(p - 2 - k - j):
                w = iop[i] + iop[j] + iop[k] + iop[pm3 - i - j - k]
                x = np.array([iop[i], iop[j], iop[k]]) / w
                for d in range(3):
                    nodes.append(x)
                    tans.append(np.eye(3)[d])
    return np.array(nodes), np.array(tans), face_off, int_off


def _leg01(n, x):
    """Shifted Legendre L_0..L_n on [0, 1] and their derivatives at x: two [len(x)][n + 1] arrays."""
    X = 2.0 * x - 1.0
    V = npleg.legvander(X, n)
    dV = np.zeros_like(V)
    for k in range(1, n + 1):
        c = np.zeros(k + 1)
        c[k] = 1.0
        dV[:, k] = 2.0 * npleg.legval(X, npleg.legder(c))
    return V, dV


def _modal_basis(p, pts):
    """A basis of the first-kind Nedelec space R_p = P_{p-1}^3 + S_p (same construction as MFEM's: full
    degree p-1 vector polynomials, then degree-(p-1) scalars times the three rotation fields about the
    centroid) and its curl, at pts: Phi[M][n][3], Curl[M][n][3]."""
    n = pts.shape[0]
    pm1 = p - 1
    Lx, dLx = _leg01(pm1, pts[:, 0])
    Ly, dLy = _leg01(pm1, pts[:, 1])
    Lz, dLz = _leg01(pm1, pts[:, 2])
    c = 0.25
    x, y, z = pts[:, 0] - c, pts[:, 1] - c, pts[:, 2] - c
    Phi, Curl = [], []

    def scalar(i, j, k):
        s = Lx[:, i] * Ly[:, j] * Lz[:, k]
        g = np.stack([dLx[:, i] * Ly[:, j] * Lz[:, k], Lx[:, i] * dLy[:, j] * Lz[:, k], Lx[:, i] * Ly[:, j] * dLz[:, k]], axis=1)
        return s, g

    zero = np.zeros(n)
    for k in range(p):
        for j in range(p - k):
            for i in range(p - k - j):
                s, g = scalar(i, j, k)
                for d in range(3):
                    v = np.zeros((n, 3))
                    v[:, d] = s
                    e = np.zeros((n, 3))
                    e[:, d] = 1.0
                    Phi.append(v)
                    Curl.append(np.cross(g, e))
    w0 = np.stack([y, -x, zero], axis=1)
    w1 = np.stack([z, zero, -x], axis=1)
    w2 = np.stack([zero, z, -y], axis=1)
    cw0, cw1, cw2 = np.array([0.0, 0, -2]), np.array([0.0, 2, 0]), np.array([-2.0, 0, 0])
    for k in range(p):
        for j in range(p - k):
            i = pm1 - j - k
            s, g = scalar(i, j, k)
            for w, cw in ((w0, cw0), (w1, cw1)):
                Phi.append(s[:, None] * w)
                Curl.append(np.cross(g, w) + s[:, None] * cw[None, :])
    for k in range(p):
        s, g = scalar(0, pm1 - k, k)
        Phi.append(s[:, None] * w2)
        Curl.append(np.cross(g, w2) + s[:, None] * cw2[None, :])
    return np.array(Phi), np.array(Curl)


_ELEMENTS = {}


def nd_tet_element(p: int) -> NDTetElement:
    if p in _ELEMENTS:
        return _ELEMENTS[p]
    nodes, tans, face_off, int_off = _dof_functionals(p)
    P = nd_tet_ndof(p)
    assert nodes.shape[0] == P, (nodes.shape, P)
    Phi, _ = _modal_basis(p, nodes)                # [M][P][3]
    assert Phi.shape[0] == P
    T = np.einsum("mdc,dc->md", Phi, tans)         # T[m][d] = functional d of phi_m
    coef = np.linalg.inv(T)                        # shape_d = sum_m coef[d][m] phi_m
    el = NDTetElement(p, P, nodes, tans, coef, 0, face_off, int_off, float(np.linalg.cond(T)))
    _ELEMENTS[p] = el
    return el


def nd_tet_tables(p: int, degree: int | None = None):
    """(interp[3][Q][P], curl[3][Q][P], qpts[Q][3], qw[Q]) with a rule exact to `degree` (default 2p, the
    reference's default for this operator, /root/reference/palace/fem/integrator.cpp:14-22)."""
    pts, w = tet_quadrature(2 * p if degree is None else degree)
    interp, curl = nd_tet_element(p).tabulate(pts)
    return interp, curl, pts, w


# ------------------------------------------------------------------------------------------------
# Meshes
# ------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class TetMesh:
    verts: np.ndarray   # [nv][3]
    elems: np.ndarray   # [ne][4] vertex ids, positively oriented
    attr: np.ndarray    # [ne] 1-based attributes
    warp: object = None  # optional smooth map R^3 -> R^3 applied to the high-order nodes (curved tets)

    @property
    def ne(self):
        return int(self.elems.shape[0])

    def node_coords(self, order: int) -> np.ndarray:
        """xe[ne][3][Nn]: nodes of the order-`order` nodal geometry (equispaced lattice, a + b + c <= order,
        a fastest), component-major like mfem::Ordering::byNODES (mesh.hpp:31-33)."""
        lat = tet_lattice(order)
        lam = np.concatenate([1.0 - lat.sum(axis=1, keepdims=True), lat], axis=1)  # barycentric
        X = np.einsum("nv,evc->enc", lam, self.verts[self.elems])                 # [ne][Nn][3]
        if self.warp is not None:
            X = self.warp(X.reshape(-1, 3)).reshape(X.shape)
        return np.ascontiguousarray(np.transpose(X, (0, 2, 1)))


def tet_lattice(order: int) -> np.ndarray:
    pts = [(a, b, c) for c in range(order + 1) for b in range(order + 1 - c) for a in range(order + 1 - b - c)]
    return np.array(pts, dtype=np.float64) / max(order, 1)


def _tet_det(verts, el):
    v = verts[el]
    return np.linalg.det(np.stack([v[1] - v[0], v[2] - v[0], v[3] - v[0]], axis=1))


def box_tet_mesh(n, size=(1.0, 1.0, 1.0), *, jitter=0.0, scramble_seed=None, n_attr=1, warp_amp=0.0) -> TetMesh:
    """n[0] x n[1] x n[2] cubes, each split into the six Kuhn tetrahedra around the main diagonal (conforming).
    `jitter` moves interior vertices (fraction of the cell size); `scramble_seed` permutes every tet's local
    vertex order (keeping it positively oriented), so that all face orientations occur; `warp_amp` > 0 attaches
    a smooth map that curves the high-order geometry."""
    if np.isscalar(n):
        n = (n, n, n)
    nx, ny, nz = n
    rng = np.random.default_rng(scramble_seed if scramble_seed is not None else 0)
    gx, gy, gz = (np.linspace(0, size[d], n[d] + 1) for d in range(3))
    vid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i
    verts = np.array([[gx[i], gy[j], gz[k]] for k in range(nz + 1) for j in range(ny + 1) for i in range(nx + 1)])
    if jitter > 0.0:
        h = np.array([size[d] / n[d] for d in range(3)])
        for k in range(1, nz):
            for j in range(1, ny):
                for i in range(1, nx):
                    verts[vid(i, j, k)] += jitter * h * (rng.random(3) - 0.5)
    elems, attr = [], []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                for perm in itertools.permutations(range(3)):
                    c = [i, j, k]
                    vs = [vid(*c)]
                    for a in perm:
                        c[a] += 1
                        vs.append(vid(*c))
                    elems.append(vs)
                    attr.append(1 + (i + j + k) % n_attr)
    elems = np.array(elems, dtype=np.int64)
    for e in range(elems.shape[0]):
        if scramble_seed is not None:
            elems[e] = elems[e][rng.permutation(4)]
        if _tet_det(verts, elems[e]) < 0:
            elems[e, [2, 3]] = elems[e, [3, 2]]
    warp = None
    if warp_amp > 0.0:
        L = np.array(size, dtype=np.float64)

        def warp(X, amp=warp_amp, L=L):
            s = X / L
            d = np.stack([np.sin(2 * np.pi * s[:, 1]) * np.sin(np.pi * s[:, 2]), np.sin(2 * np.pi * s[:, 2]) * np.sin(np.pi * s[:, 0]),
                          np.sin(2 * np.pi * s[:, 0]) * np.sin(np.pi * s[:, 1])], axis=1)
            return X + amp * L * d

    return TetMesh(verts, elems, np.array(attr, dtype=np.int32), warp)


# ------------------------------------------------------------------------------------------------
# Geometry q-data (reference layout, /root/reference/palace/fem/mesh.cpp:146-209, qfunctions/33/geom_33_qf.h:9-34)
# ------------------------------------------------------------------------------------------------


def _lagrange_tet_grad(order, pts):
    """Gradients of the order-`order` nodal (lattice) basis at pts: dN[n][Q][3]."""
    lat = tet_lattice(order)
    expo = [(a, b, c) for c in range(order + 1) for b in range(order + 1 - c) for a in range(order + 1 - b - c)]

    def vander(X, deriv=None):
        cols = []
        for (a, b, c) in expo:
            e = [a, b, c]
            coef = 1.0
            if deriv is not None:
                if e[deriv] == 0:
                    cols.append(np.zeros(X.shape[0]))
                    continue
                coef = e[deriv]
                e[deriv] -= 1
            cols.append(coef * X[:, 0] ** e[0] * X[:, 1] ** e[1] * X[:, 2] ** e[2])
        return np.stack(cols, axis=1)

    Vinv = np.linalg.inv(vander(lat))            # N_n(x) = sum_m mono_m(x) Vinv[m][n]
    return np.stack([vander(pts, d) @ Vinv for d in range(3)], axis=2).transpose(1, 0, 2)  # [n][Q][3]


def geom_qdata(xe: np.ndarray, attr: np.ndarray, order: int, qpts: np.ndarray, qw: np.ndarray) -> np.ndarray:
    """qdata[ne][11][Q] = {attr, w detJ, (adj(J)^T / detJ) column-major} for order-`order` tets."""
    dN = _lagrange_tet_grad(order, qpts)                      # [n][Q][3]
    J = np.einsum("ecn,nqd->eqcd", xe, dN)                    # J[c][d] = d x_c / d xi_d
    det = np.linalg.det(J)
    JinvT = np.transpose(np.linalg.inv(J), (0, 1, 3, 2))      # J^-T = adj(J)^T / detJ
    ne, Q = det.shape
    qd = np.empty((ne, 11, Q))
    qd[:, 0, :] = attr[:, None]
    qd[:, 1, :] = qw[None, :] * det
    for r in range(3):
        for c in range(3):
            qd[:, 2 + r + 3 * c, :] = JinvT[:, :, r, c]
    return qd


# ------------------------------------------------------------------------------------------------
# Global space: numbering, restriction, orientation transformations
# ------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class TetSpace:
    p: int
    P: int
    ndofs: int
    idx: np.ndarray           # [ne][P] int32 global dof of every local dof (native order)
    curl_orient: np.ndarray   # [ne][P][3] int8 row-major tridiagonal T_e: x_e = T_e x[idx_e]
    ess_dofs: np.ndarray      # boundary (PEC) dofs
    edges: dict               # (lo, hi) -> edge number
    faces: dict               # (g0, g1, g2) -> face number
    n_edges: int = 0
    n_faces: int = 0

    def orient_signs(self):
        """Sign-only orientation (valid when the transformation is diagonal, always at p = 1)."""
        assert not self.curl_orient[:, :, 0].any() and not self.curl_orient[:, :, 2].any()
        return np.ascontiguousarray(self.curl_orient[:, :, 1])

    def dense_T(self, e):
        P = self.P
        T = np.zeros((P, P))
        co = self.curl_orient[e]
        T[np.arange(P), np.arange(P)] = co[:, 1]
        T[np.arange(1, P), np.arange(P - 1)] = co[1:, 0]
        T[np.arange(P - 1), np.arange(1, P)] = co[:-1, 2]
        return T


def build_nd_tet_space(mesh: TetMesh, p: int) -> TetSpace:
    el = nd_tet_element(p)
    P, ne = el.P, mesh.ne
    edges, faces = {}, {}
    for e in range(ne):
        v = mesh.elems[e]
        for (a, b) in TET_EDGES:
            edges.setdefault((min(v[a], v[b]), max(v[a], v[b])), len(edges))
        for f in TET_FACES:
            faces.setdefault(tuple(sorted(int(v[t]) for t in f)), len(faces))
    n_edges, n_faces = len(edges), len(faces)
    nfd = p * (p - 1)                 # dofs per face
    nid = p * (p - 1) * (p - 2) // 2  # dofs per interior
    face_base = p * n_edges
    int_base = face_base + nfd * n_faces
    ndofs = int_base + nid * ne
    fidx = face_node_index(p)
    idx = np.zeros((ne, P), dtype=np.int32)
    co = np.zeros((ne, P, 3), dtype=np.int8)
    face_count = {}
    E2 = {0: np.array([0, 0]), 1: np.array([1, 0]), 2: np.array([0, 1])}  # x_v - x_g0 in (T1, T2) coordinates
    for e in range(ne):
        v = mesh.elems[e]
        o = 0
        for (a, b) in TET_EDGES:
            ga, gb = int(v[a]), int(v[b])
            base = p * edges[(min(ga, gb), max(ga, gb))]
            for i in range(p):
                if ga < gb:
                    idx[e, o], co[e, o, 1] = base + i, 1
                else:  # local direction against the global one: mirrored point, opposite tangent
                    idx[e, o], co[e, o, 1] = base + (p - 1 - i), -1
                o += 1
        for f in TET_FACES:
            g = [int(v[t]) for t in f]
            key = tuple(sorted(g))
            face_count[key] = face_count.get(key, 0) + 1
            base = face_base + nfd * faces[key]
            rank = [key.index(t) for t in g]        # position of local vertices a, b, c in the sorted face
            # local tangents in terms of the global ones: t1 = x_b - x_a, t2 = x_c - x_a
            M = np.array([E2[rank[1]] - E2[rank[0]], E2[rank[2]] - E2[rank[0]]])
            for j in range(p - 1):
                for i in range(p - 1 - j):
                    trip = (p - 2 - i - j, i, j)     # lattice indices attached to local vertices a, b, c
                    gtrip = [0, 0, 0]
                    for t in range(3):
                        gtrip[rank[t]] = trip[t]
                    n_glob = fidx[(gtrip[1], gtrip[2])]
                    idx[e, o], idx[e, o + 1] = base + 2 * n_glob, base + 2 * n_glob + 1
                    co[e, o, 1], co[e, o, 2] = M[0, 0], M[0, 1]
                    co[e, o + 1, 0], co[e, o + 1, 1] = M[1, 0], M[1, 1]
                    o += 2
        for t in range(nid):
            idx[e, o], co[e, o, 1] = int_base + nid * e + t, 1
            o += 1
        assert o == P
    # essential (PEC) dofs: edges and faces on the boundary
    bfaces = [k for k, c in face_count.items() if c == 1]
    ess = set()
    for k in bfaces:
        base = face_base + nfd * faces[k]
        ess.update(range(base, base + nfd))
        for (a, b) in ((0, 1), (0, 2), (1, 2)):
            eb = p * edges[(k[a], k[b])]
            ess.update(range(eb, eb + p))
    return TetSpace(p, P, ndofs, idx, co, np.array(sorted(ess), dtype=np.int32), edges, faces, n_edges, n_faces)


def interpolate(mesh: TetMesh, space: TetSpace, field) -> np.ndarray:
    """Global dof vector of a vector field through the GLOBAL functionals (straight-sided tets): an edge's
    dofs use the tangent from its lower to its higher vertex, a face's dofs the tangents of its sorted vertices."""
    p = space.p
    eop, fop, iop = _gl01(p), _gl01(p - 1), _gl01(p - 2)
    x = np.zeros(space.ndofs)
    X = mesh.verts
    for (lo, hi), k in space.edges.items():
        t = X[hi] - X[lo]
        for i in range(p):
            x[p * k + i] = t @ field(X[lo] + eop[i] * t)
    nfd = p * (p - 1)
    face_base = p * space.n_edges
    fidx = face_node_index(p)
    for (g0, g1, g2), k in space.faces.items():
        T1, T2 = X[g1] - X[g0], X[g2] - X[g0]
        for (i, j), n in fidx.items():
            w = fop[i] + fop[j] + fop[p - 2 - i - j]
            pt = (fop[p - 2 - i - j] * X[g0] + fop[i] * X[g1] + fop[j] * X[g2]) / w
            E = field(pt)
            x[face_base + nfd * k + 2 * n] = T1 @ E
            x[face_base + nfd * k + 2 * n + 1] = T2 @ E
    nid = p * (p - 1) * (p - 2) // 2
    int_base = face_base + nfd * space.n_faces
    for e in range(mesh.ne):
        v = X[mesh.elems[e]]
        Jm = np.stack([v[1] - v[0], v[2] - v[0], v[3] - v[0]], axis=1)
        o = 0
        for k in range(p - 2):
            for j in range(p - 2 - k):
                for i in range(p - 2 - k - j):
                    w = iop[i] + iop[j] + iop[k] + iop[p - 3 - i - j - k]
                    E = field(v[0] + Jm @ (np.array([iop[i], iop[j], iop[k]]) / w))
                    for d in range(3):
                        x[int_base + nid * e + o] = Jm[:, d] @ E
                        o += 1
    return x


# ------------------------------------------------------------------------------------------------
# H1 tetrahedron (auxiliary space of the Hiptmair smoother) and the element-local transfer matrices
# (discrete gradient G, p-prolongations): what Palace builds with ProjectGrad / GetTransferMatrix
# (/root/reference/palace/fem/libceed/basis.cpp:116-165, fem/bilinearform.cpp:203-282)
# ------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class H1TetElement:
    p: int
    P: int
    nodes: np.ndarray   # [P][3] reference nodes: vertices, edge interiors, face interiors, interior
    coef: np.ndarray    # [P][M] nodal basis in the modal (Legendre product) basis

    def tabulate(self, pts):
        """(values[n][P], grad[3][n][P]) of the nodal basis at pts."""
        V, dV = _h1_modal(self.p, np.asarray(pts, dtype=np.float64))
        return V @ self.coef.T, np.ascontiguousarray(np.einsum("nmd,pm->dnp", dV, self.coef))


def _h1_modal(p, pts):
    Lx, dLx = _leg01(p, pts[:, 0])
    Ly, dLy = _leg01(p, pts[:, 1])
    Lz, dLz = _leg01(p, pts[:, 2])
    V, dV = [], []
    for c in range(p + 1):
        for b in range(p + 1 - c):
            for a in range(p + 1 - b - c):
                V.append(Lx[:, a] * Ly[:, b] * Lz[:, c])
                dV.append(np.stack([dLx[:, a] * Ly[:, b] * Lz[:, c], Lx[:, a] * dLy[:, b] * Lz[:, c], Lx[:, a] * Ly[:, b] * dLz[:, c]], axis=1))
    return np.stack(V, axis=1), np.stack(dV, axis=1)   # [n][M], [n][M][3]


def _h1_nodes(p):
    V = _REF_VERTS
    nodes = [V[i] for i in range(4)]
    for (a, b) in TET_EDGES:
        for m in range(1, p):
            nodes.append(V[a] + (m / p) * (V[b] - V[a]))
    for (a, b, c) in TET_FACES:
        for j in range(1, p):
            for i in range(1, p - j):
                nodes.append(((p - i - j) * V[a] + i * V[b] + j * V[c]) / p)
    for k in range(1, p):
        for j in range(1, p - k):
            for i in range(1, p - k - j):
                nodes.append(np.array([i, j, k]) / p)
    return np.array(nodes)


_H1_ELEMENTS = {}


def h1_tet_element(p: int) -> H1TetElement:
    if p not in _H1_ELEMENTS:
        nodes = _h1_nodes(p)
        P = (p + 1) * (p + 2) * (p + 3) // 6
        assert nodes.shape[0] == P
        V, _ = _h1_modal(p, nodes)
        _H1_ELEMENTS[p] = H1TetElement(p, P, nodes, np.linalg.inv(V).T)
    return _H1_ELEMENTS[p]


@dataclasses.dataclass
class H1TetSpace:
    p: int
    P: int
    ndofs: int
    idx: np.ndarray      # [ne][P] int32 (no orientation: nodal values)
    ess_dofs: np.ndarray


def face_interior_index(p):
    out, o = {}, 0
    for j in range(1, p):
        for i in range(1, p - j):
            out[(i, j)] = o
            o += 1
    return out


def build_h1_tet_space(mesh: TetMesh, nd_space: TetSpace, p: int) -> H1TetSpace:
    """Order-p nodal space on the same mesh (shares the edge / face numbering of the ND space)."""
    el = h1_tet_element(p)
    ne = mesh.ne
    nv = mesh.verts.shape[0]
    n_e, n_f = p - 1, (p - 1) * (p - 2) // 2
    n_i = (p - 1) * (p - 2) * (p - 3) // 6
    edge_base = nv
    face_base = edge_base + n_e * nd_space.n_edges
    int_base = face_base + n_f * nd_space.n_faces
    ndofs = int_base + n_i * ne
    fidx = face_interior_index(p)
    idx = np.zeros((ne, el.P), dtype=np.int32)
    face_count = {}
    for e in range(ne):
        v = mesh.elems[e]
        idx[e, :4] = v
        o = 4
        for (a, b) in TET_EDGES:
            ga, gb = int(v[a]), int(v[b])
            base = edge_base + n_e * nd_space.edges[(min(ga, gb), max(ga, gb))]
            for m in range(1, p):
                idx[e, o] = base + ((m if ga < gb else p - m) - 1)
                o += 1
        for f in TET_FACES:
            g = [int(v[t]) for t in f]
            key = tuple(sorted(g))
            face_count[key] = face_count.get(key, 0) + 1
            base = face_base + n_f * nd_space.faces[key]
            rank = [key.index(t) for t in g]
            for j in range(1, p):
                for i in range(1, p - j):
                    trip = (p - i - j, i, j)
                    gtrip = [0, 0, 0]
                    for t in range(3):
                        gtrip[rank[t]] = trip[t]
                    idx[e, o] = base + fidx[(gtrip[1], gtrip[2])]
                    o += 1
        for t in range(n_i):
            idx[e, o] = int_base + n_i * e + t
            o += 1
        assert o == el.P
    ess = set()
    for k, c in face_count.items():
        if c != 1:
            continue
        ess.update(k)
        base = face_base + n_f * nd_space.faces[k]
        ess.update(range(base, base + n_f))
        for (a, b) in ((0, 1), (0, 2), (1, 2)):
            eb = edge_base + n_e * nd_space.edges[(k[a], k[b])]
            ess.update(range(eb, eb + n_e))
    return H1TetSpace(p, el.P, ndofs, idx, np.array(sorted(ess), dtype=np.int32))


def nd_tet_prolongation(pc: int, pf: int) -> np.ndarray:
    """[P_f][P_c]: fine functionals of the coarse shape functions (nested spaces: exact)."""
    ec, ef = nd_tet_element(pc), nd_tet_element(pf)
    interp, _ = ec.tabulate(ef.nodes)                       # [3][P_f nodes][P_c]
    return np.einsum("cnd,nc->nd", interp, ef.tangents)


def h1_tet_prolongation(pc: int, pf: int) -> np.ndarray:
    vals, _ = h1_tet_element(pc).tabulate(h1_tet_element(pf).nodes)
    return vals                                              # [P_f][P_c]


def tet_discrete_gradient(p: int) -> np.ndarray:
    """[P_nd][P_h1]: ND functionals of the gradients of the order-p nodal basis (grad H1_p is a subspace of ND_p)."""
    nd = nd_tet_element(p)
    _, grad = h1_tet_element(p).tabulate(nd.nodes)           # [3][P_nd nodes][P_h1]
    return np.einsum("cnj,nc->nj", grad, nd.tangents)


def dual_orient(space: TetSpace) -> np.ndarray:
    """Range-side transformation of interpolators: int8 rows of T^-T, so that the transposed restriction applies T^-1
    (the reference fills it from InvTransformDual, restriction.cpp:309-317). Blocks are 1x1 (signs) or unimodular 2x2."""
    ne, P = space.curl_orient.shape[:2]
    out = np.zeros_like(space.curl_orient)
    for e in range(ne):
        T = space.dense_T(e)
        D = np.rint(np.linalg.inv(T).T).astype(np.int64)
        assert np.abs(D).max() <= 1
        out[e, :, 1] = np.diag(D)
        out[e, 1:, 0] = np.diag(D, -1)
        out[e, :-1, 2] = np.diag(D, 1)
        chk = np.diag(np.diag(D)) + np.diag(np.diag(D, -1), -1) + np.diag(np.diag(D, 1), 1)
        assert (chk == D).all()
    return out


def global_interp_matrix(I_loc, in_idx, in_T, out_idx, out_Tdual, n_in, n_out):
    """Oracle-side assembled interpolator y = (1/mult) sum_e E_out^T (I_loc E_in x) with the curl-oriented
    restrictions (dense T per element or None), as ceed::Operator::Mult does for interpolators
    (/root/reference/palace/fem/libceed/operator.cpp:182-190)."""
    import scipy.sparse as sp

    ne = in_idx.shape[0]
    rows, cols, vals = [], [], []
    mult = np.zeros(n_out)
    for e in range(ne):
        M = I_loc if in_T is None else I_loc @ in_T(e)
        if out_Tdual is not None:
            M = out_Tdual(e).T @ M
        r, c = np.meshgrid(out_idx[e], in_idx[e], indexing="ij")
        rows.append(r.ravel())
        cols.append(c.ravel())
        vals.append(M.ravel())
        np.add.at(mult, out_idx[e], 1.0)
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_out, n_in))
    return sp.diags(1.0 / mult) @ A


# ------------------------------------------------------------------------------------------------
# Lowest-order Raviart-Thomas space and the discrete curl of the order-1 Nedelec space (the flux spaces of the error
# estimators on simplices, /root/reference/palace/linalg/errorestimator.cpp: B = curl E lives in RT, D = eps E is projected to it)
# ------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class RT0TetSpace:
    ndofs: int
    idx: np.ndarray      # [ne][4] face number of the local face opposite vertex f
    orient: np.ndarray   # [ne][4] +1 / -1: outward normal against the global one (vertices sorted by global id, right-hand rule)


def rt0_tet_tables(pts):
    """interp[3][Q][4]: phi_f(x) = 2 (x - v_f), unit outward flux through the face opposite reference vertex f, zero through the others."""
    pts = np.asarray(pts, dtype=np.float64)
    interp = np.zeros((3, len(pts), 4))
    for f in range(4):
        interp[:, :, f] = 2.0 * (pts - _REF_VERTS[f][None]).T
    return interp


def build_rt0_tet_space(mesh: TetMesh, nd: TetSpace) -> RT0TetSpace:
    ne = mesh.ne
    idx = np.zeros((ne, 4), dtype=np.int32)
    ori = np.zeros((ne, 4), dtype=np.int8)
    X = mesh.verts
    for e in range(ne):
        v = mesh.elems[e]
        for f in range(4):
            g = sorted(int(v[t]) for t in range(4) if t != f)
            idx[e, f] = nd.faces[tuple(g)]
            n_glob = np.cross(X[g[1]] - X[g[0]], X[g[2]] - X[g[0]])
            outward = X[g[0]] - X[int(v[f])]          # from the opposite vertex towards the face
            ori[e, f] = 1 if n_glob @ outward > 0 else -1
    return RT0TetSpace(nd.n_faces, idx, ori)


def discrete_curl_p1(nd: TetSpace, rt: RT0TetSpace):
    """Sparse [n_faces x n_edges]: the flux of curl u through a face is the circulation of u around it (order-1 ND dofs are edge
    circulations from the lower to the higher vertex id; a face's loop runs g0 -> g1 -> g2 -> g0 over its sorted vertices)."""
    import scipy.sparse as sp

    assert nd.p == 1
    rows, cols, vals = [], [], []
    for g, F in nd.faces.items():
        for (a, b, s) in ((g[0], g[1], 1.0), (g[1], g[2], 1.0), (g[0], g[2], -1.0)):
            rows.append(F)
            cols.append(nd.edges[(a, b)])
            vals.append(s)
    return sp.csr_matrix((vals, (rows, cols)), shape=(rt.ndofs, nd.ndofs))


# ------------------------------------------------------------------------------------------------
# Raviart-Thomas tetrahedra RT_k (k = p - 1: the flux space of order-p Nedelec fields) and the element-local discrete curl
# ------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class RTTetElement:
    k: int
    P: int
    nodes: np.ndarray     # [P][3] reference points of the dof functionals
    dirs: np.ndarray      # [P][3] functional d(u) = u(node) . dir: twice the outward area vector on faces, unit vectors inside
    coef: np.ndarray      # [P][M] shape function d = sum_m coef[d][m] phi_m
    face_off: int = 0     # faces first ((k+1)(k+2)/2 each, TET_FACES order), then the interior
    int_off: int = 0
    cond: float = 0.0

    def tabulate(self, pts):
        """interp[3][n][P]: reference shape functions at pts (contravariant Piola: u = J u^ / detJ)."""
        Phi = _rt_modal_basis(self.k, np.asarray(pts, dtype=np.float64))
        return np.ascontiguousarray(np.einsum("dm,mnc->cnd", self.coef, Phi))


def rt_tet_ndof(k: int) -> int:
    return (k + 1) * (k + 2) * (k + 4) // 2


def _rt_modal_basis(k, pts):
    """A basis of RT_k = P_k^3 + (x - c) P~_k at pts: Phi[M][n][3] (products of shifted Legendre polynomials; the last block uses the
    products of total degree exactly k, whose leading parts span the homogeneous polynomials)."""
    n = pts.shape[0]
    Lx, _ = _leg01(k, pts[:, 0])
    Ly, _ = _leg01(k, pts[:, 1])
    Lz, _ = _leg01(k, pts[:, 2])
    Phi = []
    for l in range(k + 1):
        for j in range(k + 1 - l):
            for i in range(k + 1 - l - j):
                s = Lx[:, i] * Ly[:, j] * Lz[:, l]
                for d in range(3):
                    v = np.zeros((n, 3))
                    v[:, d] = s
                    Phi.append(v)
    xc = pts - 0.25
    for l in range(k + 1):
        for j in range(k + 1 - l):
            i = k - l - j
            Phi.append((Lx[:, i] * Ly[:, j] * Lz[:, l])[:, None] * xc)
    return np.array(Phi)


def rt_face_index(k):
    """(i1, i2) -> running index of the (k + 1)(k + 2) / 2 points of a face's lattice (i1 + i2 <= k; weights of the second and third
    vertex of the face in its canonical, sorted-by-global-id order)."""
    out, o = {}, 0
    for j in range(k + 1):
        for i in range(k + 1 - j):
            out[(i, j)] = o
            o += 1
    return out


def _rt_functionals(k):
    V = _REF_VERTS
    nodes, dirs = [], []
    for (a, b, c) in TET_FACES:
        nf = np.cross(V[b] - V[a], V[c] - V[a])             # twice the outward area vector
        for j in range(k + 1):
            for i in range(k + 1 - j):
                w = np.array([k - i - j + 1.0, i + 1.0, j + 1.0]) / (k + 3.0)    # open lattice: strictly inside the face
                nodes.append(w[0] * V[a] + w[1] * V[b] + w[2] * V[c])
                dirs.append(nf)
    int_off = len(nodes)
    for l in range(k):
        for j in range(k - l):
            for i in range(k - l - j):
                x = np.array([i + 1.0, j + 1.0, l + 1.0]) / (k + 3.0)
                for d in range(3):
                    nodes.append(x)
                    dirs.append(np.eye(3)[d])
    return np.array(nodes), np.array(dirs), int_off


_RT_ELEMENTS = {}


def rt_tet_element(k: int) -> RTTetElement:
    if k in _RT_ELEMENTS:
        return _RT_ELEMENTS[k]
    nodes, dirs, int_off = _rt_functionals(k)
    P = rt_tet_ndof(k)
    assert nodes.shape[0] == P, (nodes.shape, P)
    Phi = _rt_modal_basis(k, nodes)
    assert Phi.shape[0] == P
    T = np.einsum("mdc,dc->md", Phi, dirs)
    el = RTTetElement(k, P, nodes, dirs, np.linalg.inv(T), 0, int_off, float(np.linalg.cond(T)))
    _RT_ELEMENTS[k] = el
    return el


@dataclasses.dataclass
class RTTetSpace:
    k: int
    P: int
    ndofs: int
    idx: np.ndarray      # [ne][P]
    orient: np.ndarray   # [ne][P] +1 / -1 (face dofs: local outward normal against the global one; interior: +1)


def _perm_parity(seq):
    s, seq = 1, list(seq)
    for i in range(len(seq)):
        for j in range(i + 1, len(seq)):
            if seq[i] > seq[j]:
                s = -s
    return s


def build_rt_tet_space(mesh: TetMesh, nd: TetSpace, k: int) -> RTTetSpace:
    """Global RT_k space on the faces numbered by `nd` (any order of nd): a face's dofs sit on its lattice in the canonical order of its
    vertices sorted by global id, with the normal of that order (right-hand rule); a tetrahedron sees them through the permutation
    of its own vertex order and the sign of that permutation."""
    el = rt_tet_element(k)
    nfd = (k + 1) * (k + 2) // 2
    nid = el.P - 4 * nfd
    fidx = rt_face_index(k)
    ne = mesh.ne
    int_base = nfd * nd.n_faces
    idx = np.zeros((ne, el.P), dtype=np.int32)
    ori = np.ones((ne, el.P), dtype=np.int8)
    for e in range(ne):
        v = mesh.elems[e]
        o = 0
        for f in TET_FACES:
            g = [int(v[t]) for t in f]
            key = tuple(sorted(g))
            rank = [key.index(t) for t in g]
            sgn = _perm_parity(rank)
            base = nfd * nd.faces[key]
            for j in range(k + 1):
                for i in range(k + 1 - j):
                    trip = (k - i - j, i, j)          # lattice weights on the local vertices a, b, c
                    gtrip = [0, 0, 0]
                    for t in range(3):
                        gtrip[rank[t]] = trip[t]
                    idx[e, o], ori[e, o] = base + fidx[(gtrip[1], gtrip[2])], sgn
                    o += 1
        for t in range(nid):
            idx[e, o] = int_base + nid * e + t
            o += 1
        assert o == el.P
    return RTTetSpace(k, el.P, int_base + nid * ne, idx, ori)


def tet_discrete_curl(p: int) -> np.ndarray:
    """[P_rt x P_nd] element matrix of the discrete curl ND_p -> RT_{p-1} in reference coordinates: entry (i, j) is the i-th RT dof
    functional of curl phi_j (the curl of a covariantly mapped field is the contravariantly mapped reference curl, so one matrix
    serves every element)."""
    nd, rt = nd_tet_element(p), rt_tet_element(p - 1)
    _, curl = nd.tabulate(rt.nodes)                       # [3][P_rt][P_nd]
    return np.einsum("cij,ic->ij", curl, rt.dirs)
