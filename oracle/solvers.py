"""TEST INFRASTRUCTURE ONLY. NumPy/SciPy restatement of the reference's smoother / multigrid /
Krylov algorithms on explicitly assembled (sparse) matrices, used to check the device-resident
solver loop. Each function cites the reference lines it follows."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def assemble_sparse(Ae, gid, n):
    """Global sparse matrix from element matrices in the (sign-folded) element basis."""
    P = gid.shape[1]
    rows = np.repeat(gid, P, axis=1).ravel()
    cols = np.tile(gid, (1, P)).ravel()
    return sp.coo_matrix((Ae.ravel(), (rows, cols)), shape=(n, n)).tocsr()


def eliminate(A, ess, diag_one=True):
    """ParOperator essential-dof handling (/root/reference/palace/linalg/rap.cpp:195-234): masked
    input, output rows replaced by x (DIAG_ONE) or 0."""
    A = A.tolil(copy=True)
    A[ess, :] = 0
    A[:, ess] = 0
    if diag_one:
        A[ess, ess] = 1.0
    return A.tocsr()


def interp_matrix(I_loc, in_gid, in_sign, out_gid, out_sign, n_in, n_out):
    """Sparse interpolator with the reference's multiplicity scaling
    (/root/reference/palace/fem/libceed/operator.cpp:182-212): y = (1/mult) .* sum_e E_out^T I E_in x."""
    ne, Po = out_gid.shape
    Pi = in_gid.shape[1]
    vals = (out_sign[:, :, None] * I_loc[None, :, :] * in_sign[:, None, :]).ravel()
    rows = np.repeat(out_gid, Pi, axis=1).ravel()
    cols = np.tile(in_gid, (1, Po)).ravel()
    M = sp.coo_matrix((vals, (rows, cols)), shape=(n_out, n_in)).tocsr()
    mult = np.bincount(out_gid.ravel(), minlength=n_out).astype(float)
    return sp.diags(1.0 / np.maximum(mult, 1.0)) @ M


# ---- chebyshev.cpp:191-220 (4th kind) and :261-293 (1st kind) ---------------------------------
def chebyshev(A, dinv, lambda_max, order, x, y, initial_guess, fourth_kind=True, pc_it=1, sf_min=0.0):
    y = y.copy()
    for it in range(pc_it):
        if initial_guess or it > 0:
            r = x - A @ y
        else:
            r = x.copy()
            y[:] = 0.0
        if fourth_kind:
            d = 4.0 / (3.0 * lambda_max) * dinv * r
            for k in range(1, order):
                y += d
                r -= A @ d
                sd = (2.0 * k - 1.0) / (2.0 * k + 3.0)
                sr = (8.0 * k + 4.0) / ((2.0 * k + 3.0) * lambda_max)
                d = sd * d + sr * dinv * r
        else:
            sfm = sf_min if sf_min > 0 else 1.69 / (order ** 1.68 + 2.11 * order + 1.98)  # chebyshev.cpp:243-247
            lmin = sfm * lambda_max
            theta, delta = 0.5 * (lambda_max + lmin), 0.5 * (lambda_max - lmin)
            d = dinv * r / theta
            rhop = delta / theta
            for k in range(1, order):
                y += d
                r -= A @ d
                rho = 1.0 / (2.0 * theta / delta - rhop)
                d = rho * rhop * d + 2.0 * rho / delta * dinv * r
                rhop = rho
        y += d
    return y


def power_iteration(A, dinv, tol=1e-4, max_it=1000, seed=0):
    """linalg/operator.cpp:583-631 applied to D^-1 A (chebyshev.cpp:14-28)."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1, 1, A.shape[0])
    u /= np.linalg.norm(u)
    l0 = 0.0
    for it in range(max_it):
        u = dinv * (A @ u)
        l = np.linalg.norm(u)
        u /= l
        if it > 0 and abs(l - l0) / l0 < tol:
            break
        l0 = l
    return l


# ---- distrelaxation.cpp:99-151 -------------------------------------------------------------------
class DistRelax:
    def __init__(self, A, A_G, G, ess_G, lam, lam_G, order, fourth_kind=True, pc_it=1):
        self.A, self.A_G, self.G, self.ess_G = A, A_G, G, ess_G
        self.dinv, self.dinv_G = 1.0 / A.diagonal(), 1.0 / A_G.diagonal()
        self.lam, self.lam_G, self.order, self.fk, self.pc_it = lam, lam_G, order, fourth_kind, pc_it

    def mult2(self, x, y, initial_guess):
        for it in range(self.pc_it):
            y = chebyshev(self.A, self.dinv, self.lam, self.order, x, y, initial_guess or it > 0, self.fk)
            r = x - self.A @ y
            xg = self.G.T @ r
            xg[self.ess_G] = 0.0
            yg = chebyshev(self.A_G, self.dinv_G, self.lam_G, self.order, xg, np.zeros_like(xg), False, self.fk)
            y = y + self.G @ yg
        return y

    def mult_transpose2(self, x, y, initial_guess):
        for it in range(self.pc_it):
            if initial_guess or it > 0:
                r = x - self.A @ y
                xg = self.G.T @ r
            else:
                y = np.zeros_like(y)
                xg = self.G.T @ x
            xg[self.ess_G] = 0.0
            yg = chebyshev(self.A_G, self.dinv_G, self.lam_G, self.order, xg, np.zeros_like(xg), False, self.fk)
            y = y + self.G @ yg
            y = chebyshev(self.A, self.dinv, self.lam, self.order, x, y, True, self.fk)
        return y


# ---- gmg.cpp:126-142,172-205 ---------------------------------------------------------------------
class Gmg:
    def __init__(self, A, P, smoothers, coarse_solve, ess, cycle_it=1):
        """A[l] level matrices (coarse->fine), P[l]: l -> l+1, smoothers[l] objects with mult2 /
        mult_transpose2 (l >= 1), coarse_solve(x), ess[l] essential dofs per level."""
        self.A, self.P, self.B, self.coarse, self.ess, self.cycle_it = A, P, smoothers, coarse_solve, ess, cycle_it

    def mult(self, x):
        L = len(self.A) - 1
        self.X = [None] * (L + 1)
        self.Y = [np.zeros(a.shape[0]) for a in self.A]
        self.X[L] = x.copy()
        for it in range(self.cycle_it):
            self.vcycle(L, it > 0)
        return self.Y[L].copy()

    def vcycle(self, l, initial_guess):
        if l == 0:
            self.Y[0] = self.coarse(self.X[0])
            return
        self.Y[l] = self.B[l].mult2(self.X[l], self.Y[l], initial_guess)
        R = self.X[l] - self.A[l] @ self.Y[l]
        self.X[l - 1] = self.P[l - 1].T @ R
        self.X[l - 1][self.ess[l - 1]] = 0.0
        self.vcycle(l - 1, False)
        self.Y[l] = self.Y[l] + self.P[l - 1] @ self.Y[l - 1]
        self.Y[l] = self.B[l].mult_transpose2(self.X[l], self.Y[l], True)


class ChebSmoother:
    """Plain Chebyshev level smoother (gmg.cpp:52-63) with the Solver::Mult2 interface."""

    def __init__(self, A, lam, order, fourth_kind=True, pc_it=1):
        self.A, self.dinv, self.lam, self.order, self.fk, self.pc_it = A, 1.0 / A.diagonal(), lam, order, fourth_kind, pc_it

    def mult2(self, x, y, initial_guess):
        return chebyshev(self.A, self.dinv, self.lam, self.order, x, y, initial_guess, self.fk, self.pc_it)

    mult_transpose2 = mult2


# ---- orthog.hpp:41-89 ----------------------------------------------------------------------------
def orthogonalize(kind, V, w):
    w = w.copy()
    m = len(V)
    H = np.zeros(m)
    if kind == 0:
        for j in range(m):
            H[j] = w @ V[j]
            w -= H[j] * V[j]
    else:
        H = np.array([w @ V[j] for j in range(m)])
        for j in range(m):
            w -= H[j] * V[j]
        if kind == 2:
            dH = np.array([w @ V[j] for j in range(m)])
            for j in range(m):
                w -= dH[j] * V[j]
            H = H + dH
    return H, w


def _plane_rotation(dx, dy):
    """iterative.cpp:73-110 (well-scaled branch)."""
    if dy == 0.0:
        return 1.0, 0.0
    if dx == 0.0:
        return 0.0, np.copysign(1.0, dy)
    d = np.hypot(dx, dy)
    return abs(dx) / d, dy / np.copysign(d, dx)


def _apply_rot(dx, dy, cs, sn):
    return cs * dx + sn * dy, -sn * dx + cs * dy


# ---- iterative.cpp:361-486 -----------------------------------------------------------------------
def cg(A, b, B=None, rel_tol=1e-6, abs_tol=0.0, max_it=100, x0=None):
    if x0 is not None:
        x = x0.copy()
        r = b - A @ x
    else:
        x = np.zeros_like(b)
        r = b.copy()
    z = B(r) if B else r.copy()
    beta = z @ r
    res = np.sqrt(abs(beta))
    if x0 is not None:
        beta_rhs = (B(b) @ b) if B else np.linalg.norm(b)
        initial_res = np.sqrt(abs(beta_rhs))
    else:
        initial_res = res
    eps = max(rel_tol * initial_res, abs_tol)
    it, hist = 0, []
    p = None
    beta_prev = 0.0
    while it < max_it and not res < eps:
        hist.append(res)
        p = z.copy() if it == 0 else z + (beta / beta_prev) * p
        z = A @ p
        alpha = beta / (z @ p)
        x += alpha * p
        r -= alpha * z
        beta_prev = beta
        z = B(r) if B else r.copy()
        beta = z @ r
        res = np.sqrt(abs(beta))
        it += 1
    hist.append(res)
    return x, it, hist


# ---- iterative.cpp:544-705 (GMRES, right/left PC) and :734-871 (FGMRES) ---------------------------
def gmres(A, b, B=None, rel_tol=1e-6, abs_tol=0.0, max_it=100, max_dim=None, orthog=0, flexible=False, right=True, x0=None):
    n = b.size
    mdim = max_it if max_dim is None or max_dim < 0 else max_dim
    x = np.zeros(n) if x0 is None else x0.copy()
    it, restart, hist = 0, 0, []
    beta = 0.0
    converged = False
    right = right or flexible
    while it < max_it:
        ig = (x0 is not None) or restart > 0
        if B and not right:
            r = B(b - A @ x) if ig else B(b)
        else:
            r = (b - A @ x) if ig else b.copy()
        if not ig:
            x[:] = 0.0
        true_beta = np.linalg.norm(r)
        if it == 0:
            if x0 is not None:
                initial_res = np.linalg.norm(B(b)) if (B and not right) else np.linalg.norm(b)
            else:
                initial_res = true_beta
            eps = max(rel_tol * initial_res, abs_tol)
        beta = true_beta
        if beta < eps:
            converged = True
            break
        V = [r / beta]
        Z = []
        H = np.zeros((mdim + 1, mdim))
        s = np.zeros(mdim + 1)
        cs, sn = np.zeros(mdim + 1), np.zeros(mdim + 1)
        s[0] = beta
        j = 0
        while True:
            hist.append(beta)
            if B and not right:
                w = B(A @ V[j])
            elif B:
                zj = B(V[j])
                if flexible:
                    Z.append(zj)
                w = A @ zj
            else:
                w = A @ V[j]
            Hj, w = orthogonalize(orthog, V, w)
            H[: j + 1, j] = Hj
            H[j + 1, j] = np.linalg.norm(w)
            V.append(w / H[j + 1, j])
            for k in range(j):
                H[k, j], H[k + 1, j] = _apply_rot(H[k, j], H[k + 1, j], cs[k], sn[k])
            cs[j], sn[j] = _plane_rotation(H[j, j], H[j + 1, j])
            H[j, j], H[j + 1, j] = _apply_rot(H[j, j], H[j + 1, j], cs[j], sn[j])
            s[j], s[j + 1] = _apply_rot(s[j], s[j + 1], cs[j], sn[j])
            beta = abs(s[j + 1])
            converged = beta < eps
            if converged or j + 1 == mdim or it + 1 == max_it:
                it += 1
                break
            j += 1
            it += 1
        for i in range(j, -1, -1):
            s[i] /= H[i, i]
            for k in range(i - 1, -1, -1):
                s[k] -= H[k, i] * s[i]
        if flexible:
            for k in range(j + 1):
                x += s[k] * Z[k]
        elif (not B) or (not right):
            for k in range(j + 1):
                x += s[k] * V[k]
        else:
            rr = sum(s[k] * V[k] for k in range(j + 1))
            x += B(rr)
        if converged:
            break
        restart += 1
    hist.append(beta)
    return x, it, hist


# ---- complex variants (OperType = ComplexOperator) ------------------------------------------------
def _cplane_rotation(dx, dy):
    """iterative.cpp:112-226 (well-scaled branches) -> (cs real, sn complex)."""
    if dy == 0.0:
        return 1.0, 0.0 + 0.0j
    if dx == 0.0:
        return 0.0, np.conj(dy) / abs(dy)
    dx2, dy2 = abs(dx) ** 2, abs(dy) ** 2
    dz2 = dx2 + dy2
    return np.sqrt(dx2 / dz2), np.conj(dy) * (dx / np.sqrt(dx2 * dz2))


def _capply_rot(dx, dy, cs, sn):
    return cs * dx + sn * dy, -np.conj(sn) * dx + cs * dy


def corthogonalize(kind, V, w):
    """orthog.hpp:41-89 with Dot(w, V_j) = V_j^H w (vector.cpp:674-685)."""
    w = w.copy()
    m = len(V)
    H = np.zeros(m, dtype=complex)
    if kind == 0:
        for j in range(m):
            H[j] = np.vdot(V[j], w)
            w -= H[j] * V[j]
    else:
        H = np.array([np.vdot(V[j], w) for j in range(m)])
        for j in range(m):
            w -= H[j] * V[j]
        if kind == 2:
            dH = np.array([np.vdot(V[j], w) for j in range(m)])
            for j in range(m):
                w -= dH[j] * V[j]
            H = H + dH
    return H, w


def cgmres(A, b, B=None, rel_tol=1e-6, abs_tol=0.0, max_it=100, max_dim=None, orthog=0, flexible=False, right=True):
    """GmresSolver / FgmresSolver<ComplexOperator> (iterative.cpp:544-871), zero initial guess."""
    n = b.size
    mdim = max_it if max_dim is None or max_dim < 0 else max_dim
    x = np.zeros(n, dtype=complex)
    it, restart = 0, 0
    right = right or flexible
    beta, eps, converged = 0.0, 0.0, False
    while it < max_it:
        ig = restart > 0
        if B and not right:
            r = B(b - A @ x) if ig else B(b)
        else:
            r = (b - A @ x) if ig else b.astype(complex)
        true_beta = np.linalg.norm(r)
        if it == 0:
            eps = max(rel_tol * true_beta, abs_tol)
        beta = true_beta
        if beta < eps:
            converged = True
            break
        V, Z = [r / beta], []
        H = np.zeros((mdim + 1, mdim), dtype=complex)
        s = np.zeros(mdim + 1, dtype=complex)
        cs, sn = np.zeros(mdim + 1), np.zeros(mdim + 1, dtype=complex)
        s[0] = beta
        j = 0
        while True:
            if B and not right:
                w = B(A @ V[j])
            elif B:
                zj = B(V[j])
                if flexible:
                    Z.append(zj)
                w = A @ zj
            else:
                w = A @ V[j]
            Hj, w = corthogonalize(orthog, V, w)
            H[: j + 1, j] = Hj
            H[j + 1, j] = np.linalg.norm(w)
            V.append(w / H[j + 1, j])
            for k in range(j):
                H[k, j], H[k + 1, j] = _capply_rot(H[k, j], H[k + 1, j], cs[k], sn[k])
            cs[j], sn[j] = _cplane_rotation(H[j, j], H[j + 1, j])
            H[j, j], H[j + 1, j] = _capply_rot(H[j, j], H[j + 1, j], cs[j], sn[j])
            s[j], s[j + 1] = _capply_rot(s[j], s[j + 1], cs[j], sn[j])
            beta = abs(s[j + 1])
            converged = beta < eps
            if converged or j + 1 == mdim or it + 1 == max_it:
                it += 1
                break
            j += 1
            it += 1
        for i in range(j, -1, -1):
            s[i] /= H[i, i]
            for k in range(i - 1, -1, -1):
                s[k] -= H[k, i] * s[i]
        if flexible:
            for k in range(j + 1):
                x += s[k] * Z[k]
        elif (not B) or (not right):
            for k in range(j + 1):
                x += s[k] * V[k]
        else:
            x += B(sum(s[k] * V[k] for k in range(j + 1)))
        if converged:
            break
        restart += 1
    return x, it, converged
