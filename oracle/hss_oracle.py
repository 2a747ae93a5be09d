"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product.

CPU (numpy, fp64) restatement of the reference's HSS hot path, working on the
generators read by ``oracle/hss_file.py``.  Each function follows the
reference routine cited in its docstring step by step (same order of
operations, same intermediate quantities ``tmp1/tmp2``, ``Dt/Vt1``,
``z/ft1/y``), so intermediate values can be compared, not just results.

Pinned (tests/test_oracle.py) against the reference itself run in this
container through ``oracle/ref.py``: ``mult`` / ``factor``+``solve`` agree with
``HSSMatrix<double>::mult/solve`` to ~1e-13 on the same generators.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg may import this.
"""
import numpy as np
import scipy.linalg as sla
from .hss_file import ipiv_to_gather


# ---------------------------------------------------------------- ID basis
def basis_apply(P, E, b):
    """U*b = P [b; E b]           (HSSBasisID::apply, HSSBasisID.hpp:155-169:
    copy b on top, E*b below, then laswp backwards)."""
    g = ipiv_to_gather(P)
    c = np.vstack([b, E @ b]) if E.shape[0] else b.copy()
    out = np.empty_like(c)
    out[g, :] = c          # inverse of the gather  (laswp fwd=false)
    return out


def basis_applyC(P, E, b):
    """U^H*b = top(P^T b) + E^H bottom(P^T b)   (HSSBasisID::applyC,
    HSSBasisID.hpp:189-203)."""
    g = ipiv_to_gather(P)
    pb = b[g, :]
    r = E.shape[1]
    if E.shape[0] == 0:
        return pb
    return pb[:r, :] + E.conj().T @ pb[r:, :]


def basis_dense(P, E):
    """dense(U) = P [I; E]        (HSSBasisID::dense, HSSBasisID.hpp:144-152)."""
    r = E.shape[1]
    return basis_apply(P, E, np.eye(r))


# ------------------------------------------------------------------- apply
def apply(nodes, x, trans=False):
    """y = op(H) x.  apply_HSS (HSSMatrix.cpp:419-435) = apply_fwd
    (HSSMatrix.apply.hpp:55-82) + apply_bwd (:84-136); the transposed sweep is
    applyT_fwd/applyT_bwd (:139-220)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    s = x.shape[1]
    root = nodes[0]
    y = np.zeros(((root.cols if trans else root.rows), s))
    tmp1 = [None] * len(nodes)
    tmp2 = [None] * len(nodes)
    flops = 0

    def in_basis(nd):   # the basis used on the way up
        return (nd.Pu, nd.Eu) if trans else (nd.Pv, nd.Ev)

    def out_basis(nd):  # the basis used on the way down
        return (nd.Pv, nd.Ev) if trans else (nd.Pu, nd.Eu)

    def fwd(i, isroot):
        nonlocal flops
        nd = nodes[i]
        P, E = in_basis(nd)
        if nd.leaf:
            if not isroot:
                off = nd.row_off if trans else nd.col_off
                m = nd.rows if trans else nd.cols
                tmp1[i] = basis_applyC(P, E, x[off:off + m, :])
                flops += 2 * E.size * s
        else:
            for c in nd.ch:
                fwd(c, False)
            if not isroot:
                cat = np.vstack([tmp1[nd.ch[0]], tmp1[nd.ch[1]]])
                tmp1[i] = basis_applyC(P, E, cat)
                flops += 2 * E.size * s

    def bwd(i, isroot):
        nonlocal flops
        nd = nodes[i]
        P, E = out_basis(nd)
        if nd.leaf:
            off = nd.col_off if trans else nd.row_off
            xoff = nd.row_off if trans else nd.col_off
            m = nd.cols if trans else nd.rows
            mx = nd.rows if trans else nd.cols
            D = nd.D.conj().T if trans else nd.D
            y[off:off + m, :] = D @ x[xoff:xoff + mx, :]
            flops += 2 * D.size * s
            if E.shape[1] and not isroot:
                y[off:off + m, :] += basis_apply(P, E, tmp2[i])
                flops += 2 * E.size * s
        else:
            c0, c1 = nd.ch
            if trans:
                B0, B1 = nd.B10.conj().T, nd.B01.conj().T
            else:
                B0, B1 = nd.B01, nd.B10
            t0 = B0 @ tmp1[c1]
            t1 = B1 @ tmp1[c0]
            flops += 2 * (B0.size + B1.size) * s
            if not (isroot or E.shape[1] == 0):
                u = basis_apply(P, E, tmp2[i])
                flops += 2 * E.size * s
                r0 = t0.shape[0]
                t0 = u[:r0, :] + t0
                t1 = u[r0:, :] + t1
            tmp2[c0], tmp2[c1] = t0, t1
            bwd(c0, False)
            bwd(c1, False)

    fwd(0, True)
    bwd(0, True)
    apply.last_flops = flops
    return y


def nnz_generators(nodes):
    """Sum of |D|+|E_u|+|E_v|+|B01|+|B10| (nnz(H) without the
    sizeof(*this) bookkeeping of HSSMatrix.cpp:316-323)."""
    return sum(n.D.size + n.Eu.size + n.Ev.size + n.B01.size + n.B10.size
               for n in nodes)


# ---------------------------------------------------------------- ULV factor
class ULV:
    """Per-node factors, as HSSFactors (HSSExtra.hpp:197-212): L_, Q_, W1_,
    Vt0_; root: D_ (LU) + piv_."""

    def __init__(self, n):
        self.L = [None] * n
        self.Q = [None] * n
        self.W1 = [None] * n
        self.Vt0 = [None] * n
        self.lu = None
        self.flops = 0


def factor(nodes):
    """HSSMatrix::factor_recursive (HSSMatrix.factor.hpp:51-147)."""
    f = ULV(len(nodes))
    Dt = [None] * len(nodes)
    Vt1 = [None] * len(nodes)

    def gemm_flops(m, n, k):
        return 2 * m * n * k

    def rec(i, isroot):
        nd = nodes[i]
        if not nd.leaf:
            c0, c1 = nd.ch
            rec(c0, False)
            rec(c1, False)
            r0, r1 = nodes[c0].U_rank, nodes[c1].U_rank
            Df = np.zeros((r0 + r1, r0 + r1))
            Df[:r0, :r0] = Dt[c0]
            Df[r0:, r0:] = Dt[c1]
            Df[:r0, r0:] = nd.B01 @ Vt1[c1].conj().T        # :74-75
            Df[r0:, :r0] = nd.B10 @ Vt1[c0].conj().T        # :76-77
            f.flops += gemm_flops(r0, r1, nd.B01.shape[1]) + \
                gemm_flops(r1, r0, nd.B10.shape[1])
            if not isroot:
                V = basis_dense(nd.Pv, nd.Ev)                # :86
                rv0 = nodes[c0].V_rank
                Vh = np.vstack([Vt1[c0] @ V[:rv0, :], Vt1[c1] @ V[rv0:, :]])
                f.flops += gemm_flops(r0, V.shape[1], rv0) + \
                    gemm_flops(r1, V.shape[1], V.shape[0] - rv0)
            Dt[c0] = Dt[c1] = Vt1[c0] = Vt1[c1] = None
        else:
            Df = nd.D.copy()
            Vh = basis_dense(nd.Pv, nd.Ev) if not isroot else None
        if isroot:
            f.lu = sla.lu_factor(Df)                         # :104-107
            n = Df.shape[0]
            f.flops += int(2 * n ** 3 / 3)
            return
        g = ipiv_to_gather(nd.Pu)
        Dp = Df[g, :]                                        # laswp fwd :109
        r, m = nd.U_rank, nd.U_rows
        if m > r:
            W1 = Dp[:r, :]
            W0 = Dp[r:, :] - nd.Eu @ W1                      # :116-118
            f.flops += gemm_flops(m - r, m, r)
            # W0 = [L 0] Q   (DenseMatrix::LQ, DenseMatrix.cpp:693-719)
            Qh, R = np.linalg.qr(W0.conj().T, mode="complete")
            L = R[:m - r, :].conj().T
            Q = Qh.conj().T
            f.flops += int(4 * (m - r) * m * m)  # ~gelqf+orglq, informative
            Q0, Q1 = Q[:m - r, :], Q[m - r:, :]
            f.L[i], f.Q[i], f.W1[i] = L, Q, W1
            f.Vt0[i] = Q0 @ Vh                               # :126-129
            Vt1[i] = Q1 @ Vh                                 # :130-131
            Dt[i] = W1 @ Q1.conj().T                         # :135-137
            f.flops += 2 * gemm_flops(m - r, Vh.shape[1], m) + \
                gemm_flops(r, r, m)
        else:                                                # :142-145
            Vt1[i] = Vh
            Dt[i] = Dp

    rec(0, True)
    return f


# ----------------------------------------------------------------- ULV solve
def solve(nodes, f, b):
    """HSSMatrix::solve = solve_fwd (HSSMatrix.solve.hpp:69-197) + solve_bwd
    (:199-238)."""
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = b[:, None]
    s = b.shape[1]
    x = np.zeros_like(b)
    n = len(nodes)
    z, ft1, y, xs = [None] * n, [None] * n, [None] * n, [None] * n

    def fwd(i, isroot):
        nd = nodes[i]
        if nd.leaf:
            fv = b[nd.row_off:nd.row_off + nd.rows, :].copy()
        else:
            c0, c1 = nd.ch
            fwd(c0, False)
            fwd(c1, False)
            f0 = ft1[c0] - nd.B01 @ z[c1]                    # :92-95
            f1 = ft1[c1] - nd.B10 @ z[c0]
            for c, fc in ((c0, f0), (c1, f1)):               # :101-128
                cn = nodes[c]
                if cn.U_rows > cn.U_rank:
                    Q0 = f.Q[c][:cn.U_rows - cn.U_rank, :]
                    fc -= f.W1[c] @ (Q0.conj().T @ y[c])
            fv = np.vstack([f0, f1])
        if isroot:
            xs[i] = sla.lu_solve(f.lu, fv)                   # :133-135
            return
        g = ipiv_to_gather(nd.Pu)
        fv = fv[g, :]                                        # :153
        r, m = nd.U_rank, nd.U_rows
        ft1[i] = fv[:r, :]
        if m > r:
            yy = fv[r:, :] - nd.Eu @ ft1[i]                  # :158-159
            y[i] = sla.solve_triangular(f.L[i], yy, lower=True)   # :160-161
            zz = f.Vt0[i].conj().T @ y[i]                    # :172/:178
            if not nd.leaf:
                zz = zz + basis_applyC(nd.Pv, nd.Ev,
                                       np.vstack([z[nd.ch[0]], z[nd.ch[1]]]))
            z[i] = zz
        else:
            if not nd.leaf:
                z[i] = basis_applyC(nd.Pv, nd.Ev,
                                    np.vstack([z[nd.ch[0]], z[nd.ch[1]]]))
            else:
                z[i] = np.zeros((nd.V_rank, s))

    def bwd(i):
        nd = nodes[i]
        if nd.leaf:
            x[nd.row_off:nd.row_off + nd.rows, :] = xs[i]    # :202
            return
        c0, c1 = nd.ch
        r0 = nodes[c0].U_rank
        parts = (xs[i][:r0, :], xs[i][r0:, :])
        for c, xc in zip((c0, c1), parts):                   # :209-224
            cn = nodes[c]
            if cn.U_rows > cn.U_rank:
                xs[c] = f.Q[c].conj().T @ np.vstack([y[c], xc])
            else:
                xs[c] = xc.copy()
        bwd(c0)
        bwd(c1)

    fwd(0, True)
    bwd(0)
    return x


def to_dense(nodes):
    """dense(H) by applying to the identity (tests only, small N)."""
    return apply(nodes, np.eye(nodes[0].cols))
