"""Synthetic frontal matrices for the BLR path (BASELINE.json configs[3]).

``laplacian_root_front(k)``: the root frontal matrix (Schur complement onto the
middle-plane separator) of the 7-point Laplacian on a k x k x k grid under
geometric nested dissection -- what the reference's sparse solver hands to
``BLRMatrix::construct_and_partial_factor`` for its top-level front
(reference src/sparse/fronts/FrontBLR.cpp:262-336).  The operator is separable,
so with the 2-D sine transform Q = Q1 (x) Q1 of the plane

    F = Q diag(s_ab) Q^T,   s_ab = 2 + lam_ab - 2 [T_ab^{-1}]_{nn},
    lam_ab = 4 - 2 cos(a pi/(k+1)) - 2 cos(b pi/(k+1)),
    T_ab = tridiag(-1, 2 + lam_ab, -1) of order n = (k-1)/2 (one half of the grid),

which needs no sparse solver.  ``plane_bisection_order``
orders the plane by recursive coordinate bisection so that consecutive index
ranges are compact clusters (the separator ordering the reference obtains from
its own nested dissection of the separator, FrontBLR.cpp:78-150).

Test/bench plumbing: numpy on the host, torch on the device for the large case.
"""
import numpy as np


def _mode_schur(k):
    """s_ab for a, b = 1..k (k odd)."""
    assert k % 2 == 1, "k must be odd: the separator is the middle plane"
    n = (k - 1) // 2
    th = np.arange(1, k + 1) * np.pi / (k + 1)
    lam1 = 2.0 - 2.0 * np.cos(th)
    lam = lam1[:, None] + lam1[None, :]
    d = 2.0 + lam
    r = np.zeros_like(d)                 # [T^{-1}]_{ii} by the continued fraction r_i = 1/(d - r_{i-1})
    for _ in range(n):
        r = 1.0 / (d - r)
    return d - 2.0 * r


def _sine_matrix(k):
    j = np.arange(1, k + 1)
    return np.sqrt(2.0 / (k + 1)) * np.sin(np.outer(j, j) * np.pi / (k + 1))


def plane_bisection_order(k, leaf):
    """Permutation of the k*k plane points (index p*k+q) by recursive coordinate
    bisection down to clusters of at most `leaf` points."""
    pts = np.stack(np.meshgrid(np.arange(k), np.arange(k), indexing="ij"), -1).reshape(-1, 2)
    out = []

    def rec(idx):
        if len(idx) <= leaf:
            out.extend(idx.tolist())
            return
        p = pts[idx]
        ax = int(np.argmax(p.max(0) - p.min(0)))
        o = idx[np.argsort(p[:, ax], kind="stable")]
        rec(o[: len(o) // 2]); rec(o[len(o) // 2:])

    rec(np.arange(k * k))
    return np.asarray(out)


def laplacian_root_front(k, leaf=256, device=None):
    """(F, perm): F = the k^2 x k^2 root front in the bisection ordering.
    device=None: numpy (small k); otherwise a torch device (fp64, row-major ==
    column-major since F is symmetric)."""
    s = _mode_schur(k)
    Q1 = _sine_matrix(k)
    perm = plane_bisection_order(k, leaf)
    if device is None:
        Q = np.kron(Q1, Q1)
        F = (Q * s.reshape(-1)) @ Q.T
        return F[np.ix_(perm, perm)], perm
    import torch
    Q1t = torch.tensor(Q1, device=device)
    st = torch.tensor(s, device=device)
    n = k * k
    F = torch.empty((n, n), dtype=torch.float64, device=device)
    Fv = F.view(k, k, k, k)              # [p, q, p', q']
    for p in range(k):
        # t[p', b] = sum_a Q1[p,a] Q1[p',a] s[a,b]; block (p, p') = Q1 diag(t[p']) Q1^T
        t = (Q1t[p][None, :] * Q1t) @ st                      # k x k  (p', b)
        blk = torch.einsum("qb,pb,rb->pqr", Q1t, t, Q1t)     # (p', q, q')
        Fv[p] = blk.permute(1, 0, 2)                           # [q, p', q']
    pt = torch.tensor(perm, device=device)
    return F[pt][:, pt].contiguous(), perm
