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
# The sweeps are written per node; the classic routines below run them over
# all nodes (children have larger pre-order indices than their parent, so
# descending index order is a valid bottom-up order), the sharded oracle
# (class ShardOracle) over the owned / replicated subsets.

class _ApplyState:
    def __init__(self, nodes, x, trans):
        self.nodes, self.x, self.trans = nodes, x, trans
        root = nodes[0]
        self.y = np.zeros(((root.cols if trans else root.rows), x.shape[1]))
        self.tmp1 = [None] * len(nodes)
        self.tmp2 = [None] * len(nodes)
        self.flops = 0

    def _in(self, nd):   # basis used on the way up
        return (nd.Pu, nd.Eu) if self.trans else (nd.Pv, nd.Ev)

    def _out(self, nd):  # basis used on the way down
        return (nd.Pv, nd.Ev) if self.trans else (nd.Pu, nd.Eu)

    def fwd_node(self, i):
        """apply_fwd / applyT_fwd at one node (HSSMatrix.apply.hpp:55-82,
        :139-166); the root has no basis."""
        nd = self.nodes[i]
        if i == 0:
            return
        P, E = self._in(nd)
        s = self.x.shape[1]
        if nd.leaf:
            off = nd.row_off if self.trans else nd.col_off
            m = nd.rows if self.trans else nd.cols
            self.tmp1[i] = basis_applyC(P, E, self.x[off:off + m, :])
        else:
            cat = np.vstack([self.tmp1[nd.ch[0]], self.tmp1[nd.ch[1]]])
            self.tmp1[i] = basis_applyC(P, E, cat)
        self.flops += 2 * E.size * s

    def bwd_node(self, i):
        """apply_bwd / applyT_bwd at one node (HSSMatrix.apply.hpp:84-136,
        :168-220)."""
        nd = self.nodes[i]
        isroot = (i == 0)
        P, E = self._out(nd)
        x, trans, s = self.x, self.trans, self.x.shape[1]
        if nd.leaf:
            off = nd.col_off if trans else nd.row_off
            xoff = nd.row_off if trans else nd.col_off
            m = nd.cols if trans else nd.rows
            mx = nd.rows if trans else nd.cols
            D = nd.D.conj().T if trans else nd.D
            self.y[off:off + m, :] = D @ x[xoff:xoff + mx, :]
            self.flops += 2 * D.size * s
            if E.shape[1] and not isroot:
                self.y[off:off + m, :] += basis_apply(P, E, self.tmp2[i])
                self.flops += 2 * E.size * s
        else:
            c0, c1 = nd.ch
            if trans:
                B0, B1 = nd.B10.conj().T, nd.B01.conj().T
            else:
                B0, B1 = nd.B01, nd.B10
            t0 = B0 @ self.tmp1[c1]
            t1 = B1 @ self.tmp1[c0]
            self.flops += 2 * (B0.size + B1.size) * s
            if not (isroot or E.shape[1] == 0):
                u = basis_apply(P, E, self.tmp2[i])
                self.flops += 2 * E.size * s
                r0 = t0.shape[0]
                t0 = u[:r0, :] + t0
                t1 = u[r0:, :] + t1
            self.tmp2[c0], self.tmp2[c1] = t0, t1


def apply(nodes, x, trans=False):
    """y = op(H) x.  apply_HSS (HSSMatrix.cpp:419-435) = apply_fwd
    (HSSMatrix.apply.hpp:55-82) + apply_bwd (:84-136); the transposed sweep is
    applyT_fwd/applyT_bwd (:139-220)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    st = _ApplyState(nodes, x, trans)
    for i in range(len(nodes) - 1, -1, -1):
        st.fwd_node(i)
    for i in range(len(nodes)):
        st.bwd_node(i)
    apply.last_flops = st.flops
    return st.y


def nnz_generators(nodes):
    """Sum of |D|+|E_u|+|E_v|+|B01|+|B10| (nnz(H) without the
    sizeof(*this) bookkeeping of HSSMatrix.cpp:316-323)."""
    return sum(n.D.size + n.Eu.size + n.Ev.size + n.B01.size + n.B10.size
               for n in nodes)


# ---------------------------------------------------------------- ULV factor
class ULV:
    """Per-node factors, as HSSFactors (HSSExtra.hpp:197-212): L_, Q_, W1_,
    Vt0_; root: D_ (LU) + piv_; plus the WorkFactor hand-offs Dt, Vt1
    (HSSExtra.hpp:142-147)."""

    def __init__(self, n):
        self.L = [None] * n
        self.Q = [None] * n
        self.W1 = [None] * n
        self.Vt0 = [None] * n
        self.Dt = [None] * n
        self.Vt1 = [None] * n
        self.lu = None
        self.flops = 0


def _factor_node(nodes, f, i, root=0, partial=False):
    """One node of HSSMatrix::factor_recursive (HSSMatrix.factor.hpp:51-147).
    root/partial: partial_factor() runs the recursion on child(0) with
    isroot = partial = true (factor.hpp:44-50)."""
    nd = nodes[i]
    isroot = (i == root)

    def gemm_flops(m, n, k):
        return 2 * m * n * k

    if not nd.leaf:
        c0, c1 = nd.ch
        r0, r1 = nodes[c0].U_rank, nodes[c1].U_rank
        Df = np.zeros((r0 + r1, r0 + r1))
        Df[:r0, :r0] = f.Dt[c0]
        Df[r0:, r0:] = f.Dt[c1]
        Df[:r0, r0:] = nd.B01 @ f.Vt1[c1].conj().T        # :74-75
        Df[r0:, :r0] = nd.B10 @ f.Vt1[c0].conj().T        # :76-77
        f.flops += gemm_flops(r0, r1, nd.B01.shape[1]) + \
            gemm_flops(r1, r0, nd.B10.shape[1])
        if not isroot or partial:                          # :81
            V = basis_dense(nd.Pv, nd.Ev)                  # :86
            rv0 = nodes[c0].V_rank
            Vh = np.vstack([f.Vt1[c0] @ V[:rv0, :], f.Vt1[c1] @ V[rv0:, :]])
            f.flops += gemm_flops(r0, V.shape[1], rv0) + \
                gemm_flops(r1, V.shape[1], V.shape[0] - rv0)
    else:
        Df = nd.D.copy()
        Vh = basis_dense(nd.Pv, nd.Ev) if (not isroot or partial) else None
    if isroot:
        f.lu = sla.lu_factor(Df)                           # :104-107
        n = Df.shape[0]
        f.flops += int(2 * n ** 3 / 3)
        if partial:
            f.Vhat = Vh                                    # :107
        return
    g = ipiv_to_gather(nd.Pu)
    Dp = Df[g, :]                                          # laswp fwd :109
    r, m = nd.U_rank, nd.U_rows
    if m > r:
        W1 = Dp[:r, :]
        W0 = Dp[r:, :] - nd.Eu @ W1                        # :116-118
        f.flops += gemm_flops(m - r, m, r)
        # W0 = [L 0] Q   (DenseMatrix::LQ, DenseMatrix.cpp:693-719)
        Qh, R = np.linalg.qr(W0.conj().T, mode="complete")
        L = R[:m - r, :].conj().T
        Q = Qh.conj().T
        f.flops += int(4 * (m - r) * m * m)  # ~gelqf+orglq, informative
        Q0, Q1 = Q[:m - r, :], Q[m - r:, :]
        f.L[i], f.Q[i], f.W1[i] = L, Q, W1
        f.Vt0[i] = Q0 @ Vh                                 # :126-129
        f.Vt1[i] = Q1 @ Vh                                 # :130-131
        f.Dt[i] = W1 @ Q1.conj().T                         # :135-137
        f.flops += 2 * gemm_flops(m - r, Vh.shape[1], m) + \
            gemm_flops(r, r, m)
    else:                                                  # :142-145
        f.Vt1[i] = Vh
        f.Dt[i] = Dp


def factor(nodes):
    """HSSMatrix::factor_recursive (HSSMatrix.factor.hpp:51-147)."""
    f = ULV(len(nodes))
    for i in range(len(nodes) - 1, -1, -1):
        _factor_node(nodes, f, i)
    return f


# ------------------------------------------- Schur complement (HSS fronts)
def _subtree(nodes, i):
    out, stack = [], [i]
    while stack:
        j = stack.pop()
        out.append(j)
        if not nodes[j].leaf:
            stack.extend(nodes[j].ch)
    return sorted(out)


def partial_factor(nodes):
    """HSSMatrix::partial_factor (HSSMatrix.factor.hpp:44-50): ULV of the
    subtree of child(0) with child(0) as root; keeps Vhat = Vh(child 0)."""
    f = ULV(len(nodes))
    c0 = nodes[0].ch[0]
    for i in reversed(_subtree(nodes, c0)):
        _factor_node(nodes, f, i, root=c0, partial=True)
    return f


def _sub_apply(nodes, top, t2_top, x, trans):
    """Down-sweep of the subtree of `top` started from t2(top) = t2_top with
    the up-sweep results of `x` (None: x = 0, i.e. apply_UV_big,
    HSSMatrix.Schur.hpp:248-320); returns (rows of the result that belong to
    `top`, t1(top))."""
    nd = nodes[top]
    n = nodes[0].rows
    s = t2_top.shape[1]
    xin = np.zeros((n, s)) if x is None else x
    st = _ApplyState(nodes, xin, trans)
    sub = _subtree(nodes, top)
    for i in reversed(sub):
        st.fwd_node(i)
    t1_top = st.tmp1[top]
    st.tmp2[top] = t2_top
    for i in sub:
        st.bwd_node(i)
    off = nd.col_off if trans else nd.row_off
    cnt = nd.cols if trans else nd.rows
    return st.y[off:off + cnt, :], t1_top


def schur_update(nodes, f):
    """HSSMatrix::Schur_update (HSSMatrix.Schur.hpp:40-59): Theta = U1big B10,
    DUB01 = D0^{-1} U0 B01, Phi = V1big DUB01^H."""
    root = nodes[0]
    c0, c1 = root.ch
    n0 = nodes[c0]
    DUB01 = sla.lu_solve(f.lu, basis_apply(n0.Pu, n0.Eu, root.B01))
    Theta, _ = _sub_apply(nodes, c1, root.B10, None, False)
    Phi, _ = _sub_apply(nodes, c1, DUB01.conj().T.copy(), None, True)
    return Theta, DUB01, Phi


def schur_product_direct(nodes, f, Theta, DUB01, Phi, R):
    """HSSMatrix::Schur_product_direct (HSSMatrix.Schur.hpp:73-137):
    Sr = H11 R - Theta Vhat^H DUB01 (V1big^H R), Sc = H11^H R - Phi Vhat B10^H (U1big^H R)."""
    root = nodes[0]
    c1 = root.ch[1]
    n1 = nodes[c1]
    n = root.rows
    s = R.shape[1]
    Rf = np.zeros((n, s))
    Rf[n1.col_off:n1.col_off + n1.cols, :] = R
    Sr, t1r = _sub_apply(nodes, c1, np.zeros((n1.U_rank, s)), Rf, False)
    Rf = np.zeros((n, s))
    Rf[n1.row_off:n1.row_off + n1.rows, :] = R
    Sc, t1c = _sub_apply(nodes, c1, np.zeros((n1.V_rank, s)), Rf, True)
    Sr = Sr - Theta @ (f.Vhat.conj().T @ (DUB01 @ t1r))
    Sc = Sc - Phi @ (f.Vhat @ (root.B10.conj().T @ t1c))
    return Sr, Sc


# ----------------------------------------------------------------- ULV solve
class _SolveState:
    def __init__(self, nodes, f, b):
        n = len(nodes)
        self.nodes, self.f, self.b = nodes, f, b
        self.x = np.zeros_like(b)
        self.z, self.ft1, self.y, self.xs, self.xc = ([None] * n for _ in range(5))
        # nodes whose ft1 already contains the "- W1 Q0^H y" term of
        # solve.hpp:101-128 (the sharded protocol folds it in at the child,
        # as the GPU engine does, because Q, W1, y of a remote child are not
        # available to the parent)
        self.folded = set()
        self.root = 0    # the node that is LU-solved (child 0 of the root in a partial solve)

    def fwd_node(self, i):
        """solve_fwd at one node (HSSMatrix.solve.hpp:69-197)."""
        nodes, f = self.nodes, self.f
        z, ft1, y = self.z, self.ft1, self.y
        nd = nodes[i]
        s = self.b.shape[1]
        if nd.leaf:
            fv = self.b[nd.row_off:nd.row_off + nd.rows, :].copy()
        else:
            c0, c1 = nd.ch
            f0 = ft1[c0] - nd.B01 @ z[c1]                    # :92-95
            f1 = ft1[c1] - nd.B10 @ z[c0]
            for c, fc in ((c0, f0), (c1, f1)):               # :101-128
                cn = nodes[c]
                if c not in self.folded and cn.U_rows > cn.U_rank:
                    Q0 = f.Q[c][:cn.U_rows - cn.U_rank, :]
                    fc -= f.W1[c] @ (Q0.conj().T @ y[c])
            fv = np.vstack([f0, f1])
        if i == self.root:
            self.xs[i] = sla.lu_solve(f.lu, fv)              # :133-135
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

    def bwd_node(self, i):
        """solve_bwd at one node (HSSMatrix.solve.hpp:199-238).  The reference
        applies Q(c)^H [y(c); x_c] while visiting the parent (:209-224); here
        the same product is done when the child itself is visited, so that a
        sharded run only touches Q/y of nodes it owns."""
        nodes, f = self.nodes, self.f
        nd = nodes[i]
        if i != self.root:
            if nd.U_rows > nd.U_rank:
                self.xs[i] = f.Q[i].conj().T @ np.vstack([self.y[i], self.xc[i]])
            else:
                self.xs[i] = self.xc[i].copy()
        if nd.leaf:
            self.x[nd.row_off:nd.row_off + nd.rows, :] = self.xs[i]    # :202
            return
        c0, c1 = nd.ch
        r0 = nodes[c0].U_rank
        self.xc[c0], self.xc[c1] = self.xs[i][:r0, :], self.xs[i][r0:, :]


def solve(nodes, f, b):
    """HSSMatrix::solve = solve_fwd (HSSMatrix.solve.hpp:69-197) + solve_bwd
    (:199-238)."""
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        b = b[:, None]
    st = _SolveState(nodes, f, b)
    for i in range(len(nodes) - 1, -1, -1):
        st.fwd_node(i)
    for i in range(len(nodes)):
        st.bwd_node(i)
    return st.x


def partial_forward_solve(nodes, f, b0):
    """child(0)->forward_solve(w, b0, partial=true) after partial_factor
    (HSSMatrix.solve.hpp:52-60, :133-152; FrontHSS.cpp:452-462): returns the
    solve state (w) and reduced_rhs = Vhat^H x + V^H [z0; z1]."""
    b0 = np.asarray(b0, dtype=np.float64)
    if b0.ndim == 1:
        b0 = b0[:, None]
    c0 = nodes[0].ch[0]
    n0 = nodes[c0]
    b = np.zeros((nodes[0].rows, b0.shape[1]))
    b[:n0.rows, :] = b0
    st = _SolveState(nodes, f, b)
    st.root = c0
    for i in reversed(_subtree(nodes, c0)):
        st.fwd_node(i)
    red = f.Vhat.conj().T @ st.xs[c0]
    if not n0.leaf:
        red = red + basis_applyC(n0.Pv, n0.Ev, np.vstack([st.z[n0.ch[0]], st.z[n0.ch[1]]]))
    return st, red


def partial_backward_solve(nodes, st):
    """child(0)->backward_solve(w, x0) (HSSMatrix.solve.hpp:62-66, :199-238);
    st.xs[st.root] may have been updated in between (FrontHSS.cpp:487-495)."""
    for i in _subtree(nodes, st.root):
        st.bwd_node(i)
    return st.x[:nodes[st.root].rows, :]


def to_dense(nodes):
    """dense(H) by applying to the identity (tests only, small N)."""
    return apply(nodes, np.eye(nodes[0].cols))


# ------------------------------------------------------ subtree-sharded oracle
class ShardOracle:
    """CPU engine with the begin/end protocol of strumpack_b200.dist (same
    payload layout as the C ABI SB200_d_hss_dist_*): rank `rank` of `world`
    owns the subtree of the rank-th node at depth log2(world), the nodes above
    the cut are replicated.  Built from the per-node sweeps above, i.e. the
    arithmetic is the reference's; only the order/ownership differs.  Buffers
    are torch CPU tensors (gloo)."""

    def __init__(self, nodes, world, rank):
        import torch
        self.torch = torch
        self.nodes, self.world, self.rank = nodes, world, rank
        depth = {0: 0}
        for i, n in enumerate(nodes):
            for c in n.ch:
                depth[c] = depth[i] + 1
        d = int(np.log2(world))
        assert 1 << d == world
        self.cut = [i for i in range(len(nodes)) if depth[i] == d]
        assert len(self.cut) == world
        self.top = [i for i in range(len(nodes)) if depth[i] < d]
        lo = self.cut[rank]
        self.own = [lo] + [i for i in range(lo + 1, len(nodes))
                           if depth[i] > d and all(depth[j] > d for j in range(lo + 1, i + 1))]
        n = nodes[lo]
        self.lo, self.hi = n.row_off, n.row_off + n.rows
        self.f = None

    def sizes(self, s):
        c = [self.nodes[i] for i in self.cut]
        return [max(max(n.U_rank, n.V_rank) for n in c) * s,
                max(n.U_rank * (n.V_rank + n.U_rank) for n in c),
                max(n.V_rank + n.U_rank for n in c) * s]

    def new_buffer(self, n):
        return self.torch.zeros(n, dtype=self.torch.float64)

    # ---- apply (non-transposed) -------------------------------------------
    def mult_begin(self, xT, send, trans="N"):
        assert trans == "N"
        self._ap = _ApplyState(self.nodes, xT.numpy().T.copy(), False)
        for i in reversed(self.own):
            self._ap.fwd_node(i)
        t = self._ap.tmp1[self.cut[self.rank]]
        send.zero_()
        send[:t.size] = self.torch.from_numpy(t.ravel(order="F").copy())

    def mult_end(self, xT, yT, recv, trans="N"):
        s = xT.shape[0]
        st = self.sizes(s)[0]
        for c, i in enumerate(self.cut):
            r = self.nodes[i].V_rank
            self._ap.tmp1[i] = recv[c * st:c * st + r * s].numpy().reshape((r, s), order="F").copy()
        for i in reversed(self.top):
            self._ap.fwd_node(i)
        for i in self.top:
            self._ap.bwd_node(i)
        for i in self.own:
            self._ap.bwd_node(i)
        yT[:, self.lo:self.hi] = self.torch.from_numpy(self._ap.y[self.lo:self.hi, :].T.copy())

    # ---- ULV factor ---------------------------------------------------------
    def factor_begin(self, send):
        self.f = ULV(len(self.nodes))
        for i in reversed(self.own):
            _factor_node(self.nodes, self.f, i)
        i = self.cut[self.rank]
        blk = np.hstack([self.f.Vt1[i], self.f.Dt[i].T])     # r x (rv + r)
        send.zero_()
        send[:blk.size] = self.torch.from_numpy(blk.ravel(order="F").copy())

    def factor_end(self, recv):
        st = self.sizes(1)[1]
        for c, i in enumerate(self.cut):
            n = self.nodes[i]
            r, rv = n.U_rank, n.V_rank
            blk = recv[c * st:c * st + r * (rv + r)].numpy().reshape((r, rv + r), order="F")
            self.f.Vt1[i] = blk[:, :rv].copy()
            self.f.Dt[i] = blk[:, rv:].T.copy()
        for i in reversed(self.top):
            _factor_node(self.nodes, self.f, i)

    # ---- ULV solve ----------------------------------------------------------
    def solve_begin(self, bT, send):
        self._sv = _SolveState(self.nodes, self.f, bT.numpy().T.copy())
        for i in reversed(self.own):
            self._sv.fwd_node(i)
        i = self.cut[self.rank]
        n = self.nodes[i]
        ft1 = self._sv.ft1[i].copy()
        if n.U_rows > n.U_rank:   # fold "- W1 Q0^H y" (solve.hpp:101-128) into ft1
            Q0 = self.f.Q[i][:n.U_rows - n.U_rank, :]
            ft1 -= self.f.W1[i] @ (Q0.conj().T @ self._sv.y[i])
        blk = np.vstack([self._sv.z[i], ft1])                # (rv + r) x s
        send.zero_()
        send[:blk.size] = self.torch.from_numpy(blk.ravel(order="F").copy())

    def solve_end(self, bT, recv):
        s = bT.shape[0]
        st = self.sizes(s)[2]
        for c, i in enumerate(self.cut):
            n = self.nodes[i]
            r, rv = n.U_rank, n.V_rank
            blk = recv[c * st:c * st + (rv + r) * s].numpy().reshape((rv + r, s), order="F")
            self._sv.z[i] = blk[:rv, :].copy()
            self._sv.ft1[i] = blk[rv:, :].copy()
            self._sv.folded.add(i)
        for i in reversed(self.top):
            self._sv.fwd_node(i)
        for i in self.top:
            self._sv.bwd_node(i)
        for i in self.own:
            self._sv.bwd_node(i)
        bT[:, self.lo:self.hi] = self.torch.from_numpy(self._sv.x[self.lo:self.hi, :].T.copy())
