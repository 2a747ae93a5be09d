"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product.

numpy/scipy restatement of the reference's BLR path,
routine by routine:

  tiles                ClusterTree(n).refine(leaf)       src/structured/ClusterTree.hpp:104-114
  low_rank (RRQR)      DenseMatrix::low_rank             src/dense/DenseMatrix.cpp:792-810
                       geqp3tol stopping rule            src/dense/lapack/dgeqp3tol.f:203-209
  create_LR_tile       keep low rank iff r (m+n) <= m n  src/BLR/BLRMatrix.cpp:563-570
  compress_and_factor  right-looking tile LU             src/BLR/BLRMatrix.cpp:113-185
  construct_and_partial_factor (RL)                      src/BLR/BLRMatrix.cpp:739-905
  solve                laswp + trsm L + trsm U           src/BLR/BLRMatrix.hpp:118-122, .cpp:1667-1709
  trsmLNU_gemm / gemm_trsmUNN                            src/BLR/BLRMatrix.cpp:1552-1665

Pinned by tests/test_oracle.py against golden outputs of the reference itself
(tests/golden/blr_*.npz, generator tests/golden/make_golden_blr.py) and against
the live reference library when it is present.  Tiles are kept as (U, V) pairs
(U m x r, V r x n) or dense arrays; pivoting is LAPACK getrf per diagonal tile.
"""
import numpy as np
import scipy.linalg as sla


def refine(n, leaf):
    """ClusterTree(n).refine(leaf).leaf_sizes()"""
    if n >= 2 * leaf:
        return refine(n // 2, leaf) + refine(n - n // 2, leaf)
    return [n]


def low_rank(T, rel_tol, abs_tol, max_rank=5000):
    """DenseMatrix::low_rank: column-pivoted QR truncated by the geqp3tol rule
    (stop at the first c with |R_cc| / |R_00| <= rel_tol or |R_cc| <= abs_tol).
    Returns (U = Q[:, :r], V = R[:r, :] P^T)."""
    m, n = T.shape
    if min(m, n) == 0:
        return np.zeros((m, 0)), np.zeros((0, n))
    Q, R, piv = sla.qr(T, mode="economic", pivoting=True)
    d = np.abs(np.diag(R))
    r = 0
    while r < min(len(d), max_rank):
        if d[r] <= abs_tol or (d[0] > 0 and d[r] / d[0] <= rel_tol):
            break
        r += 1
    V = np.zeros((r, n))
    V[:, piv] = np.triu(R[:r, :])
    return Q[:, :r].copy(), V


class Tile:
    """LRTile (U, V) or DenseTile D."""

    def __init__(self, D=None, U=None, V=None):
        self.D, self.U, self.V = D, U, V

    @property
    def lr(self):
        return self.D is None

    @property
    def rank(self):
        return self.U.shape[1] if self.lr else 0      # DenseTile::maximum_rank() = 0

    def dense(self):
        return self.U @ self.V if self.lr else self.D

    def nonzeros(self):
        return self.U.size + self.V.size if self.lr else self.D.size


def make_tile(T, opts, admissible=True):
    """create_LR_tile / create_dense_tile (BLRMatrix.cpp:146-147, 563-570)."""
    if not admissible:
        return Tile(D=T.copy())
    U, V = low_rank(T, opts["rel_tol"], opts["abs_tol"], opts.get("max_rank", 5000))
    m, n = T.shape
    if U.shape[1] * (m + n) > m * n:
        return Tile(D=T.copy())
    return Tile(U=U, V=V)


def _left(t, f):
    """apply f to the row space carrier of a tile: f(U) for LR, f(D) for dense"""
    if t.lr:
        t.U = f(t.U)
    else:
        t.D = f(t.D)


def _right(t, f):
    if t.lr:
        t.V = f(t.V)
    else:
        t.D = f(t.D)


class BLRFactors:
    """Result of compress_and_factor / construct_and_partial_factor (RL)."""

    def __init__(self, A, tiles, nsteps, opts, admissible=None):
        A = np.array(A, dtype=np.float64, order="F")
        self.off = np.concatenate([[0], np.cumsum(tiles)]).astype(int)
        nb = len(tiles)
        self.nb, self.nsteps, self.n = nb, nsteps, A.shape[0]
        self.t = {}
        self.lu = {}
        o = self.off
        blk = lambda i, j: A[o[i]:o[i + 1], o[j]:o[j + 1]]
        for i in range(nsteps):
            lu, piv = sla.lu_factor(blk(i, i))               # DenseTile::LU (BLRMatrix.cpp:131-135)
            self.lu[i] = (lu, piv)
            L = np.tril(lu, -1) + np.eye(lu.shape[0])
            Uu = np.triu(lu)
            perm = np.arange(lu.shape[0])
            for q, p in enumerate(piv):
                perm[q], perm[p] = perm[p], perm[q]
            for j in range(i + 1, nb):
                adm = True if admissible is None or max(i, j) >= admissible.shape[0] else bool(admissible[i, j])
                tij = make_tile(blk(i, j), opts, adm)         # :146-155
                _left(tij, lambda X: sla.solve_triangular(L, X[perm, :], lower=True, unit_diagonal=True))
                self.t[(i, j)] = tij
                adm = True if admissible is None or max(i, j) >= admissible.shape[0] else bool(admissible[j, i])
                tji = make_tile(blk(j, i), opts, adm)         # :160-166
                _right(tji, lambda X: sla.solve_triangular(Uu, X.T, trans="T", lower=False).T)
                self.t[(j, i)] = tji
            for j in range(i + 1, nb):                        # Schur updates, always into full rank (:170-185)
                for k in range(i + 1, nb):
                    a, b = self.t[(k, i)], self.t[(i, j)]
                    if a.lr and b.lr:
                        upd = a.U @ ((a.V @ b.U) @ b.V)
                    else:
                        upd = a.dense() @ b.dense()
                    blk(k, j)[...] -= upd
        self.A = A     # the trailing block holds the Schur complement of a partial factorization

    # -- statistics as BLRMatrix::rank / nonzeros -------------------------------
    def rank(self):
        return max([t.rank for t in self.t.values()] + [0])

    def schur(self):
        s = self.off[self.nsteps]
        return self.A[s:, s:]

    # -- solves -------------------------------------------------------------------
    def forward(self, b):
        """laswp + trsm(L) over the eliminated block rows; the remaining rows get
        b_k -= F21_kj b_j (trsmLNU_gemm)."""
        x = np.array(b, dtype=np.float64)
        o = self.off
        for i in range(self.nb):
            for j in range(min(i, self.nsteps)):
                x[o[i]:o[i + 1]] -= self.t[(i, j)].dense() @ x[o[j]:o[j + 1]]
            if i < self.nsteps:
                lu, piv = self.lu[i]
                xi = x[o[i]:o[i + 1]]
                for q, p in enumerate(piv):
                    xi[[q, p]] = xi[[p, q]]
                L = np.tril(lu, -1) + np.eye(lu.shape[0])
                x[o[i]:o[i + 1]] = sla.solve_triangular(L, xi, lower=True, unit_diagonal=True)
        return x

    def backward(self, y):
        """y_i <- U_ii^{-1} (y_i - sum_{j>i} T_ij y_j) for the eliminated rows (gemm_trsmUNN)."""
        x = np.array(y, dtype=np.float64)
        o = self.off
        for i in range(self.nsteps - 1, -1, -1):
            for j in range(i + 1, self.nb):
                x[o[i]:o[i + 1]] -= self.t[(i, j)].dense() @ x[o[j]:o[j + 1]]
            x[o[i]:o[i + 1]] = sla.solve_triangular(np.triu(self.lu[i][0]), x[o[i]:o[i + 1]], lower=False)
        return x

    def solve(self, b):
        assert self.nsteps == self.nb
        return self.backward(self.forward(b))


def compress_and_factor(A, leaf, rel_tol, abs_tol=1e-12, admissible=None):
    tiles = refine(A.shape[0], leaf)
    return BLRFactors(A, tiles, len(tiles), dict(rel_tol=rel_tol, abs_tol=abs_tol), admissible)


def construct_and_partial_factor(A11, A12, A21, A22, leaf, rel_tol, abs_tol=1e-12):
    n1, n2 = A11.shape[0], A22.shape[0]
    A = np.block([[A11, A12], [A21, A22]])
    t1 = refine(n1, leaf)
    tiles = t1 + (refine(n2, leaf) if n2 else [])
    return BLRFactors(A, tiles, len(t1), dict(rel_tol=rel_tol, abs_tol=abs_tol))
