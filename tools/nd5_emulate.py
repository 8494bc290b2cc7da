"""Lane-level NumPy emulation of nd_hex_apply5_kernel (one element per warp, the two half-warps share every
(qy,qz) line / item through the mirror symmetry of the 1-D tables). Design aid: validates the index maps, the
mirrored contractions and the shared-memory bank behaviour before the CUDA version goes to the GPU.
SIMT semantics: every per-lane quantity is an array of 32; shfl_xor(v, 16) == v[lane ^ 16]."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import pyoracle as O
from palace_b200.host import coeff as cf
from palace_b200.host import hexspace as hs
from tests import common
from tools.smem_sim import wavefronts, ideal

LANE = np.arange(32)
H = LANE >> 4          # half-warp
T = LANE & 15          # index inside the half


class Smem:
    """Shared array of doubles that records wavefronts (any-lane model) of each warp-wide access."""

    def __init__(self, n):
        self.a = np.full(n, np.nan)
        self.stats = {}

    def _acc(self, name, idx, mask):
        ad = [(int(l), int(8 * idx[l])) for l in range(32) if mask[l]]
        if not ad:
            return
        s = self.stats.setdefault(name, [0, 0])
        s[0] += wavefronts(ad, 8, "any")
        s[1] += ideal(ad, 8)

    def ld(self, name, idx, mask=None):
        mask = np.ones(32, bool) if mask is None else mask
        self._acc(name, idx, mask)
        return np.where(mask, self.a[np.where(mask, idx, 0)], 0.0)

    def st(self, name, idx, val, mask=None):
        mask = np.ones(32, bool) if mask is None else mask
        self._acc(name, idx, mask)
        for l in range(32):
            if mask[l]:
                self.a[idx[l]] = val[l]


def layout(p, q, MASS, CURL, PADA, PADB, PADY):
    n = p + 1
    NXA, NNA = p * q, n * q
    L = {}
    L["A_XA"] = 0; L["A_XB"] = NXA; L["A_ZA"] = L["A_XB"] + (NXA if CURL else 0); L["RSA"] = L["A_ZA"] + NNA + PADA
    L["B_YA"] = 0; L["B_YB"] = NNA; L["RSB"] = L["B_YB"] + (NNA if CURL else 0) + PADB
    L["ZA0"] = 0; L["ZB0"] = n * L["RSA"]; L["ZSZ"] = L["ZB0"] + p * L["RSB"]
    o = 0
    for nm, on, sz in (("Y_X1", MASS, NXA), ("Y_X2", CURL, NXA), ("Y_X3", CURL, NXA), ("Y_Y1", True, NNA), ("Y_Y2", CURL, NNA),
                       ("Y_Z1", True, NNA), ("Y_Z3", CURL, NNA)):
        L[nm] = o
        o += sz if on else 0
    L["RSY"] = o + PADY
    L["Y0"] = L["ZSZ"]; L["WTOT"] = L["Y0"] + q * L["RSY"]
    return L


def kernel_pads(p, q, MASS, CURL):
    """Row paddings exactly as ND5Layout computes them (RSA, RSB = 1 mod 16; RSY = 4 mod 8)."""
    n = p + 1
    LA = p * q * (2 if CURL else 1) + n * q
    LB = n * q * (2 if CURL else 1)
    LY = p * q * ((1 if MASS else 0) + (2 if CURL else 0)) + n * q * (2 + (2 if CURL else 0))
    return (1 - LA) % 16, (1 - LB) % 16, (0 if LY % 8 == 4 else (4 - LY) % 16)


def element_apply(p, q, KIND, Bo, Bc, Gc, u_e, D_apply, pads=None):
    """u_e: [P] signed element dofs in lexicographic order; D_apply(u[3,Q], c[3,Q]) -> (v, w) in logical point order
    (x fastest). Returns y_e [P] and the Smem statistics."""
    assert q == 4
    n = p + 1
    MASS, CURL = KIND in (1, 2), KIND in (0, 2)
    D3 = p * n * n
    L = layout(p, q, MASS, CURL, *(pads or kernel_pads(p, q, MASS, CURL)))
    RSA, RSB, RSY = L["RSA"], L["RSB"], L["RSY"]
    sW = Smem(L["WTOT"])
    sgn = np.where(H == 1, -1.0, 1.0)
    y_e = np.zeros_like(u_e)

    # ------------------------------------------------------------------ phase Z
    # x-/y-directed items: half 0 lane t < p*n is x-directed dof t = i + p*j ; half 1 lane t is y-directed dof t = i + n*j
    IPX = p * n
    vxy = T < IPX
    t = np.where(vxy, T, 0)
    ux = np.stack([u_e[H * D3 + t + IPX * k] for k in range(n)])          # [n][32]
    rowx = np.where(H == 0, L["ZA0"] + (t // p) * RSA + L["A_XA"] + q * (t % p), L["ZB0"] + (t // n) * RSB + L["B_YA"] + q * (t % n))
    offb = np.where(H == 0, L["A_XB"] - L["A_XA"], L["B_YB"] - L["B_YA"])
    for qz in range(q):
        a = sum(Bc[qz, k] * ux[k] for k in range(n))
        sW.st("Z st xy", rowx + qz, a, vxy)
        if CURL:
            b = sum(Gc[qz, k] * ux[k] for k in range(n))
            sW.st("Z st xy", rowx + offb + qz, b, vxy)
    # z-directed items tz = i + n*j (16 per half): half h handles qz = c (h=0) or q-1-c (h=1) with mirrored k
    IPZ = n * n
    vz = T < IPZ
    tz = np.where(vz, T, 0)
    uz = np.stack([u_e[2 * D3 + tz + IPZ * np.where(H == 0, k, p - 1 - k)] for k in range(p)])
    rowz = L["ZA0"] + (tz // n) * RSA + L["A_ZA"] + q * (tz % n)
    for c in range(q // 2):
        a = sum(Bo[c, k] * uz[k] for k in range(p))
        sW.st("Z st z", rowz + np.where(H == 0, c, q - 1 - c), a, vz)

    # ------------------------------------------------------------------ phase Y : item (qz,i) = column w ; half -> 2 of the qy
    IX, IN = p * q, n * q
    vx, vn = T < IX, T < IN
    wx, wn = np.where(vx, T, 0), np.where(vn, T, 0)
    jn = [np.where(H == 0, j, n - 1 - j) for j in range(n)]
    jp = [np.where(H == 0, j, p - 1 - j) for j in range(p)]
    xa = [sW.ld("Y ld", L["ZA0"] + jn[j] * RSA + L["A_XA"] + wx) for j in range(n)]
    xb = [sW.ld("Y ld", L["ZA0"] + jn[j] * RSA + L["A_XB"] + wx) for j in range(n)] if CURL else None
    ya = [sW.ld("Y ld", L["ZB0"] + jp[j] * RSB + L["B_YA"] + wn) for j in range(p)]
    yb = [sW.ld("Y ld", L["ZB0"] + jp[j] * RSB + L["B_YB"] + wn) for j in range(p)] if CURL else None
    za = [sW.ld("Y ld", L["ZA0"] + jn[j] * RSA + L["A_ZA"] + wn) for j in range(n)]
    for c in range(q // 2):
        row = L["Y0"] + np.where(H == 0, c, q - 1 - c) * RSY
        if MASS:
            sW.st("Y st", row + L["Y_X1"] + wx, sum(Bc[c, j] * xa[j] for j in range(n)), vx)
        if CURL:
            sW.st("Y st", row + L["Y_X2"] + wx, sum(Bc[c, j] * xb[j] for j in range(n)), vx)
            sW.st("Y st", row + L["Y_X3"] + wx, sgn * sum(Gc[c, j] * xa[j] for j in range(n)), vx)
        sW.st("Y st", row + L["Y_Y1"] + wn, sum(Bo[c, j] * ya[j] for j in range(p)), vn)
        if CURL:
            sW.st("Y st", row + L["Y_Y2"] + wn, sum(Bo[c, j] * yb[j] for j in range(p)), vn)
        sW.st("Y st", row + L["Y_Z1"] + wn, sum(Bc[c, j] * za[j] for j in range(n)), vn)
        if CURL:
            sW.st("Y st", row + L["Y_Z3"] + wn, sgn * sum(Gc[c, j] * za[j] for j in range(n)), vn)

    # ------------------------------------------------------------------ phase XDX : line s = qy + q*qz = T ; half -> qx = c or q-1-c
    s = T
    qy, qz = s % q, s // q
    WXb = L["Y0"] + qy * RSY + qz
    ip = [np.where(H == 0, i, p - 1 - i) for i in range(p)]
    inn = [np.where(H == 0, i, n - 1 - i) for i in range(n)]
    z = np.zeros(32)
    x1 = [sW.ld("XDX ld", WXb + L["Y_X1"] + q * ip[i]) for i in range(p)] if MASS else [z] * p
    x2 = [sW.ld("XDX ld", WXb + L["Y_X2"] + q * ip[i]) for i in range(p)] if CURL else [z] * p
    x3 = [sW.ld("XDX ld", WXb + L["Y_X3"] + q * ip[i]) for i in range(p)] if CURL else [z] * p
    y1 = [sW.ld("XDX ld", WXb + L["Y_Y1"] + q * inn[i]) for i in range(n)]
    y2 = [sW.ld("XDX ld", WXb + L["Y_Y2"] + q * inn[i]) for i in range(n)] if CURL else [z] * n
    z1 = [sW.ld("XDX ld", WXb + L["Y_Z1"] + q * inn[i]) for i in range(n)]
    z3 = [sW.ld("XDX ld", WXb + L["Y_Z3"] + q * inn[i]) for i in range(n)] if CURL else [z] * n
    uu = np.zeros((2, 3, 32)); cc = np.zeros((2, 3, 32))
    for c in range(2):
        u0 = sum(Bo[c, i] * x1[i] for i in range(p))
        dzux = sum(Bo[c, i] * x2[i] for i in range(p))
        dyux = sum(Bo[c, i] * x3[i] for i in range(p))
        u1 = sum(Bc[c, i] * y1[i] for i in range(n))
        dzuy = sum(Bc[c, i] * y2[i] for i in range(n))
        dxuy = sum(Gc[c, i] * y1[i] for i in range(n))
        u2 = sum(Bc[c, i] * z1[i] for i in range(n))
        dyuz = sum(Bc[c, i] * z3[i] for i in range(n))
        dxuz = sum(Gc[c, i] * z1[i] for i in range(n))
        uu[c] = [u0, u1, u2]
        cc[c] = [dyuz - dzuy, dzux - sgn * dxuz, sgn * dxuy - dyux]
    # pointwise D at (qx, qy, qz), logical point index iq = qx + q*(qy + q*qz)
    Q = q ** 3
    U = np.zeros((3, Q)); C = np.zeros((3, Q))
    for l in range(32):
        for c in range(2):
            qx = c if H[l] == 0 else q - 1 - c
            iq = qx + q * (qy[l] + q * qz[l])
            U[:, iq] = uu[c, :, l]; C[:, iq] = cc[c, :, l]
    V, Wc = D_apply(U, C)
    vv = np.zeros((2, 3, 32)); cw = np.zeros((2, 3, 32))
    for l in range(32):
        for c in range(2):
            qx = c if H[l] == 0 else q - 1 - c
            iq = qx + q * (qy[l] + q * qz[l])
            vv[c, :, l] = V[:, iq]; cw[c, :, l] = Wc[:, iq]
    g1, g2 = sgn * cw[:, 1], sgn * cw[:, 2]
    # transposed x-contraction: partial sums over this lane's two points, for the (mirrored) output index i'
    a1 = [sum(Bo[c, i] * vv[c, 0] for c in range(2)) for i in range(p)]
    a2 = [sum(Bo[c, i] * cw[c, 1] for c in range(2)) for i in range(p)]
    a3 = [-sum(Bo[c, i] * cw[c, 2] for c in range(2)) for i in range(p)]
    b1 = [sum(Bc[c, i] * vv[c, 1] + Gc[c, i] * g2[c] for c in range(2)) for i in range(n)]
    b2 = [-sum(Bc[c, i] * cw[c, 0] for c in range(2)) for i in range(n)]
    c1 = [sum(Bc[c, i] * vv[c, 2] - Gc[c, i] * g1[c] for c in range(2)) for i in range(n)]
    c3 = [sum(Bc[c, i] * cw[c, 0] for c in range(2)) for i in range(n)]

    def finish(acc, ln, off, on=True):
        if not on:
            return
        for i in range((ln + 1) // 2):
            tot = acc[i] + acc[ln - 1 - i][LANE ^ 16]
            iout = np.where(H == 0, i, ln - 1 - i)
            sW.st("XDX st", WXb + off + q * iout, tot)
    finish(a1, p, L["Y_X1"], MASS); finish(a2, p, L["Y_X2"], CURL); finish(a3, p, L["Y_X3"], CURL)
    finish(b1, n, L["Y_Y1"]); finish(b2, n, L["Y_Y2"], CURL); finish(c1, n, L["Y_Z1"]); finish(c3, n, L["Y_Z3"], CURL)

    # ------------------------------------------------------------------ phase Yt : partial over the lane's two qy, mirrored j'
    rows = [L["Y0"] + np.where(H == 0, c, q - 1 - c) * RSY for c in range(2)]
    X1 = [sW.ld("Yt ld", rows[c] + L["Y_X1"] + wx) for c in range(2)] if MASS else [z, z]
    X2 = [sW.ld("Yt ld", rows[c] + L["Y_X2"] + wx) for c in range(2)] if CURL else [z, z]
    X3 = [sgn * sW.ld("Yt ld", rows[c] + L["Y_X3"] + wx) for c in range(2)] if CURL else [z, z]
    Y1 = [sW.ld("Yt ld", rows[c] + L["Y_Y1"] + wn) for c in range(2)]
    Y2 = [sW.ld("Yt ld", rows[c] + L["Y_Y2"] + wn) for c in range(2)] if CURL else [z, z]
    Z1 = [sW.ld("Yt ld", rows[c] + L["Y_Z1"] + wn) for c in range(2)]
    Z3 = [sgn * sW.ld("Yt ld", rows[c] + L["Y_Z3"] + wn) for c in range(2)] if CURL else [z, z]
    pa = [sum(Bc[c, j] * X1[c] + Gc[c, j] * X3[c] for c in range(2)) for j in range(n)]
    pb = [sum(Bc[c, j] * X2[c] for c in range(2)) for j in range(n)]
    qa = [sum(Bo[c, j] * Y1[c] for c in range(2)) for j in range(p)]
    qb = [sum(Bo[c, j] * Y2[c] for c in range(2)) for j in range(p)]
    ra = [sum(Bc[c, j] * Z1[c] + Gc[c, j] * Z3[c] for c in range(2)) for j in range(n)]

    def finish_z(acc, ln, base, stride, col, mask, on=True):
        if not on:
            return
        for j in range((ln + 1) // 2):
            tot = acc[j] + acc[ln - 1 - j][LANE ^ 16]
            jout = np.where(H == 0, j, ln - 1 - j)
            sW.st("Yt st", base + jout * stride + col, tot, mask)
    finish_z(pa, n, L["ZA0"] + L["A_XA"], RSA, wx, vx); finish_z(pb, n, L["ZA0"] + L["A_XB"], RSA, wx, vx, CURL)
    finish_z(qa, p, L["ZB0"] + L["B_YA"], RSB, wn, vn); finish_z(qb, p, L["ZB0"] + L["B_YB"], RSB, wn, vn, CURL)
    finish_z(ra, n, L["ZA0"] + L["A_ZA"], RSA, wn, vn)

    # ------------------------------------------------------------------ phase Zt
    XA = [sW.ld("Zt ld", rowx + qz_) for qz_ in range(q)]
    XB = [sW.ld("Zt ld", rowx + offb + qz_) for qz_ in range(q)] if CURL else [z] * q
    for k in range(n):
        o = sum(Bc[qz_, k] * XA[qz_] + Gc[qz_, k] * XB[qz_] for qz_ in range(q))
        for l in range(32):
            if vxy[l]:
                y_e[H[l] * D3 + t[l] + IPX * k] += o[l]
    vz0 = vz & (H == 0)
    ZA = [sW.ld("Zt ld", rowz + qz_, vz0) for qz_ in range(q)]
    for k in range(p):
        o = sum(Bo[qz_, k] * ZA[qz_] for qz_ in range(q))
        for l in range(32):
            if vz0[l]:
                y_e[2 * D3 + tz[l] + IPZ * k] += o[l]
    return y_e, sW.stats, L


def main():
    p, q = 3, 4
    for KIND in (2, 0, 1):
        prob = common.make_problem(n=(2, 1, 1), p=p, n_attr=2)
        sp = prob.nd
        blob = common.coefficient(KIND, 2, "matrix", a_mass=0.7, a_curl=1.3)
        rng = np.random.default_rng(0)
        x = rng.standard_normal(sp.ndofs)
        y_ref = common.oracle_apply(prob, KIND, blob, x)
        t1 = hs.tables_1d(p, q)
        Bo, Bc, Gc = (np.asarray(a).reshape(q, -1) for a in (t1.Bo, t1.Bc, t1.Gc))
        # symmetry the kernel relies on
        assert np.allclose(Bo, Bo[::-1, ::-1], atol=1e-14) and np.allclose(Bc, Bc[::-1, ::-1], atol=1e-14)
        assert np.allclose(Gc, -Gc[::-1, ::-1], atol=1e-13)
        y = np.zeros(sp.ndofs)
        stats = None
        for e in range(prob.mesh.ne if hasattr(prob.mesh, "ne") else sp.lex_gid.shape[0]):
            gid, sg = sp.lex_gid[e], sp.lex_sign[e].astype(float)
            u_e = sg * x[gid]
            qd = prob.qdata_ref[e]
            y_e, stats, L = element_apply(p, q, KIND, Bo, Bc, Gc, u_e, lambda U, C: O.apply_D(KIND, blob, qd, U, C))
            np.add.at(y, gid, sg * y_e)
        err = np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref)
        print(f"kind {KIND}: rel err vs oracle {err:.2e}   WTOT {L['WTOT']} doubles  RSA {L['RSA']} RSB {L['RSB']} RSY {L['RSY']}")
        tot = [0, 0]
        for k, v in stats.items():
            print(f"   {k:10s} wavefronts {v[0]:4d} ideal {v[1]:4d}")
            tot[0] += v[0]; tot[1] += v[1]
        print("   total", tot)


if __name__ == "__main__":
    main()
