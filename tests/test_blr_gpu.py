"""GPU parity tests of the BLR path (kernels K11-K14)
through the C ABI, following the reference's own test/test_BLR_seq.cpp:
Toeplitz matrix, tiles from ClusterTree(n).refine(leaf), weak admissibility,
compress_and_factor, then ||X - B\\(A X)||_F/||X||_F <= 1e2*max(rtol, atol)
(:182-196) -- and against the reference library run on the same matrix."""
import numpy as np
import pytest

from conftest import have_ref

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def toeplitz(n):
    i = np.arange(n)
    return 1.0 / (1.0 + np.abs(i[:, None] - i[None, :]))


@pytest.mark.parametrize("n,leaf,tol", [(2048, 256, 1e-4), (1000, 128, 1e-6), (777, 64, 1e-8)])
def test_blr_lu_solve_toeplitz(built, n, leaf, tol):
    sb = built
    A = toeplitz(n)
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    B = sb.BLRMatrix.compress_and_factor(A, o)
    assert (B.rows, B.cols) == (n, n)
    X = np.random.default_rng(0).standard_normal((n, 10))
    Y = A @ X
    Xs = B.solve(Y)
    assert rel(Xs, X) <= 1e2 * max(tol, 1e-12)        # test_BLR_seq.cpp:192
    assert 0 < B.rank < leaf                           # off-diagonal tiles are low rank
    assert B.nonzeros < 0.6 * n * n
    assert B.launches > 0


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_blr_matches_reference(built):
    from oracle import ref
    sb = built
    n, leaf, tol = 2048, 256, 1e-4
    A = toeplitz(n)
    R = ref.RefBLR(A, f"--blr_leaf_size {leaf} --blr_rel_tol {tol}")
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    B = sb.BLRMatrix.compress_and_factor(A, o)
    assert B.tiles == len(R.tiles)
    X = np.random.default_rng(1).standard_normal((n, 4))
    Y = A @ X
    # both are approximate inverses at the compression tolerance: compare to the
    # 10*eps_compress bound of the north star and to each other's error level
    xr, xs = R.solve(Y), B.solve(Y)
    assert rel(xs, X) <= 1e2 * tol and rel(xr, X) <= 1e2 * tol
    assert rel(xs, xr) <= 10 * 1e2 * tol
    assert abs(B.rank - R.info()["rank"]) <= 3


def test_blr_compress_mult(built):
    sb = built
    n = 1500
    A = toeplitz(n)
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=1e-6, abs_tol=1e-12, leaf_size=128)
    B = sb.StructuredMatrix.from_dense(A, o)      # compress only (construct_from_dense, BLR)
    x = np.random.default_rng(2).standard_normal((n, 3))
    assert rel(B.mult(x), A @ x) <= 1e2 * 1e-6
    with pytest.raises(RuntimeError):
        B.factor()                                # as in the reference: not supported


def test_blr_upper_triangular_zero_tiles(built):
    """rank-0 tiles (strictly lower part is zero)."""
    sb = built
    n = 1024
    A = np.triu(toeplitz(n))
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=1e-6, abs_tol=1e-12, leaf_size=128)
    B = sb.BLRMatrix.compress_and_factor(A, o)
    X = np.random.default_rng(3).standard_normal((n, 2))
    assert rel(B.solve(A @ X), X) <= 1e-4


def _with_noise_tiles(n, leaf, blocks, seed=5):
    """Toeplitz + full-rank noise in the listed off-diagonal tiles + a strong diagonal."""
    A = toeplitz(n) + 4.0 * np.eye(n)
    rng = np.random.default_rng(seed)
    for (i, j) in blocks:
        A[i * leaf:(i + 1) * leaf, j * leaf:(j + 1) * leaf] += 0.02 * rng.standard_normal((leaf, leaf))
    return A


@pytest.mark.parametrize("blocks", [[(0, 2)], [(3, 1)], [(0, 2), (3, 1), (1, 2), (2, 1), (3, 0)]])
def test_blr_dense_offdiagonal_tiles(built, blocks):
    """Tiles that do not compress to rank <= min(m,n)/2 stay dense (DenseTile,
    reference BLRMatrix.cpp:563-570) and take part in trsm / Schur update / solve."""
    sb = built
    n, leaf, tol = 1024, 256, 1e-6
    A = _with_noise_tiles(n, leaf, blocks)
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    B = sb.BLRMatrix.compress_and_factor(A, o)
    assert B.dense_tiles >= len(blocks)       # fill-in may densify more tiles later in the LU
    X = np.random.default_rng(0).standard_normal((n, 3))
    assert rel(B.solve(A @ X), X) <= 1e2 * tol
    assert 0 < B.rank < leaf // 2     # rank() only counts the low-rank tiles, as in the reference


def test_blr_all_tiles_dense_matches_lapack(built):
    sb = built
    n, leaf = 768, 128
    A = np.random.default_rng(7).standard_normal((n, n)) + n * np.eye(n) * 0.1
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=1e-8, abs_tol=1e-12, leaf_size=leaf)
    B = sb.BLRMatrix.compress_and_factor(A, o)
    assert B.dense_tiles == B.tiles * (B.tiles - 1)     # every off-diagonal tile
    X = np.random.default_rng(1).standard_normal((n, 2))
    Y = A @ X
    assert rel(B.solve(Y), np.linalg.solve(A, Y)) <= 1e-10    # plain blocked LU, no compression error
    C = sb.StructuredMatrix.from_dense(A, o)                   # compress only: mult through dense tiles
    assert rel(C.mult(X), Y) <= 1e-12


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_blr_dense_tiles_match_reference(built):
    from oracle import ref
    sb = built
    n, leaf, tol = 1024, 256, 1e-4
    A = _with_noise_tiles(n, leaf, [(0, 1), (2, 0)])
    R = ref.RefBLR(A, f"--blr_leaf_size {leaf} --blr_rel_tol {tol}")
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    B = sb.BLRMatrix.compress_and_factor(A, o)
    X = np.random.default_rng(1).standard_normal((n, 4))
    Y = A @ X
    xr, xs = R.solve(Y), B.solve(Y)
    assert rel(xs, X) <= 1e2 * tol and rel(xr, X) <= 1e2 * tol
    assert B.dense_tiles >= 2
    assert abs(B.rank - R.info()["rank"]) <= 3     # BLRMatrix::rank(): dense tiles count 0


# ---- partially factored fronts (BLRMatrix::construct_and_partial_factor) -------
def _front(n1, n2, shift=2.0):
    A = toeplitz(n1 + n2) + shift * np.eye(n1 + n2)
    return A, A[:n1, :n1], A[:n1, n1:], A[n1:, :n1], A[n1:, n1:]


@pytest.mark.parametrize("n1,n2,leaf,tol", [(1024, 512, 256, 1e-4), (700, 500, 128, 1e-6),
                                            (300, 900, 128, 1e-6), (512, 0, 128, 1e-6)])
def test_blr_partial_factor_schur_and_solves(built, n1, n2, leaf, tol):
    """The front [A11 A12; A21 A22] eliminated over A11 only (reference
    BLRMatrix.cpp:739-1037): the trailing block becomes the Schur complement and
    the two half solves (trsmLNU_gemm / gemm_trsmUNN, FrontBLR.cpp:525-568)
    around a dense Schur solve give the solution of the whole system, to the
    reference's own bound 1e2*tol (test_BLR_seq.cpp:192)."""
    sb = built
    A, A11, A12, A21, A22 = _front(n1, n2)
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    F, S = sb.BLRMatrix.construct_and_partial_factor(A11, A12, A21, A22, o)
    assert (F.rows, F.sep_rows) == (n1 + n2, n1)
    if n2:
        S_exact = A22 - A21 @ np.linalg.solve(A11, A12)
        assert rel(S, S_exact) <= 1e2 * tol
        assert 0 < F.rank < leaf
    X = np.random.default_rng(4).standard_normal((n1 + n2, 3))
    B = A @ X
    f = F.partial_forward_solve(B)
    y2 = np.linalg.solve(S, f[n1:]) if n2 else f[n1:]
    y = F.partial_backward_solve(np.vstack([f[:n1], y2]))
    assert rel(y, X) <= 1e2 * tol
    if n2:
        with pytest.raises(RuntimeError):
            F.solve(B)          # a partial factorization is not a solver for the whole front
    else:                       # no update block: the two halves ARE the solve
        o2 = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
        G = sb.BLRMatrix.compress_and_factor(A, o2)
        assert rel(y, G.solve(B)) <= 1e-12


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_blr_partial_factor_matches_reference(built):
    from oracle import ref
    sb = built
    n1, n2, leaf, tol = 1024, 768, 256, 1e-4
    A, A11, A12, A21, A22 = _front(n1, n2)
    R = ref.RefBLRFront(A11, A12, A21, A22, f"--blr_leaf_size {leaf} --blr_rel_tol {tol}")
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    F, S = sb.BLRMatrix.construct_and_partial_factor(A11, A12, A21, A22, o)
    S_exact = A22 - A21 @ np.linalg.solve(A11, A12)
    es, er = rel(S, S_exact), rel(R.S, S_exact)
    assert es <= 1e2 * tol and er <= 1e2 * tol
    assert rel(S, R.S) <= 10 * 1e2 * tol             # two approximations at the same tolerance
    assert es <= 10 * max(er, 1e-12)                 # and no less accurate than the reference's
    assert abs(F.rank - max(R.ranks())) <= 3
    b = np.random.default_rng(6).standard_normal((n1 + n2, 2))
    fs, fr = F.partial_forward_solve(b), R.partial_forward_solve(b)
    # b_upd - A21 A11^{-1} b_sep does not depend on how the LU pivots
    assert rel(fs[n1:], fr[n1:]) <= 10 * 1e2 * tol
    y2 = np.linalg.solve(S_exact, fs[n1:])
    ys = F.partial_backward_solve(np.vstack([fs[:n1], y2]))
    yr = R.partial_backward_solve(np.vstack([fr[:n1], np.linalg.solve(S_exact, fr[n1:])]))
    assert rel(ys, yr) <= 10 * 1e2 * tol
    assert rel(A @ ys, b) <= 1e2 * tol


def test_blr_transposed_mult_and_from_elements(built):
    """StructuredMatrix::mult(Trans::T/C) on a compressed BLR matrix (gemv with
    the transposed tiles, reference BLRMatrix.cpp:1742-1763) -- on a
    non-symmetric matrix with low-rank, dense and zero tiles -- and
    construct_from_elements for Type::BLR (StructuredMatrix.cpp:230-252)."""
    sb = built
    n, leaf, tol = 1024, 128, 1e-8
    i = np.arange(n)
    A = 1.0 / (1.0 + np.abs(i[:, None] - 1.7 * i[None, :]) / 3.0) + np.triu(toeplitz(n), 300)
    A[0:128, 256:384] += 0.1 * np.random.default_rng(9).standard_normal((128, 128))   # a dense tile
    A[640:768, 0:128] = 0.0                                                            # a zero tile
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-14, leaf_size=leaf)
    B = sb.StructuredMatrix.from_dense(A, o)
    x = np.random.default_rng(2).standard_normal((n, 3))
    assert rel(B.mult(x), A @ x) <= 1e2 * tol
    assert rel(B.mult(x, "T"), A.T @ x) <= 1e2 * tol
    assert rel(B.mult(x[:, :1], "C"), A.T @ x[:, :1]) <= 1e2 * tol

    m = 300
    M = toeplitz(m) + np.eye(m)
    E = sb.StructuredMatrix.from_elements(m, m, lambda r, c: float(M[r, c]),
                                          sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=1e-8, leaf_size=64))
    xm = x[:m]
    assert rel(E.mult(xm), M @ xm) <= 1e2 * 1e-8


def test_blr_left_looking_equals_right_looking(built):
    """BLRFactorAlgorithm::LL (reference BLRMatrix.cpp:186-212, 846-1013): block row /
    column i receives the updates of all earlier steps right before step i.  The
    engine applies the same tile updates in the same order as in the RL schedule,
    so the factors -- hence ranks and Schur complements -- are identical; the solves
    agree to roundoff (the block substitution accumulates with atomics)."""
    sb = built
    n, leaf, tol = 1536, 128, 1e-6
    A = _with_noise_tiles(n, leaf, [(0, 3), (5, 2)])        # low-rank, dense and updated tiles
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    X = np.random.default_rng(0).standard_normal((n, 3))
    Y = A @ X
    R = sb.BLRMatrix.compress_and_factor(A, o, factor_algorithm=sb.BLR_RL)
    L = sb.BLRMatrix.compress_and_factor(A, o, factor_algorithm=sb.BLR_LL)
    xr, xl = R.solve(Y), L.solve(Y)
    assert rel(xl, X) <= 1e2 * tol
    assert rel(xl, xr) <= 1e-12
    assert (R.rank, R.nonzeros, R.dense_tiles) == (L.rank, L.nonzeros, L.dense_tiles)
    for alg in (sb.BLR_COLWISE, sb.BLR_COMB, sb.BLR_STAR, 7):   # not built: rejected, never run as something else
        with pytest.raises(RuntimeError):
            sb.BLRMatrix.compress_and_factor(A, o, factor_algorithm=alg)
    # the front: A22 receives its update at the end in the LL schedule
    n1 = 896
    Fr, Sr = sb.BLRMatrix.construct_and_partial_factor(A[:n1, :n1], A[:n1, n1:], A[n1:, :n1], A[n1:, n1:], o,
                                                       factor_algorithm=sb.BLR_RL)
    Fl, Sl = sb.BLRMatrix.construct_and_partial_factor(A[:n1, :n1], A[:n1, n1:], A[n1:, :n1], A[n1:, n1:], o,
                                                       factor_algorithm=sb.BLR_LL)
    assert rel(Sl, Sr) <= 1e-13
    print("LL Schur complement bitwise equal to RL:", np.array_equal(Sr, Sl))
    S_exact = A[n1:, n1:] - A[n1:, :n1] @ np.linalg.solve(A[:n1, :n1], A[:n1, n1:])
    assert rel(Sl, S_exact) <= 1e2 * tol
    assert rel(Fl.partial_forward_solve(Y), Fr.partial_forward_solve(Y)) <= 1e-12


def test_blr_strong_admissibility(built):
    """compress_and_factor(A, admissible, opts) with a strong-admissibility mask
    (reference adm_t, BLRMatrix.hpp:78; FrontBLR.cpp:262-281): inadmissible tiles
    are DenseTiles (BLRMatrix.cpp:146-147) even though they would compress."""
    sb = built
    n, leaf, tol = 1024, 128, 1e-6
    A = toeplitz(n) + 2.0 * np.eye(n)
    nb = n // leaf
    t = np.arange(nb)
    adm = np.abs(t[:, None] - t[None, :]) > 1              # neighbours of the diagonal stay dense
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    W = sb.BLRMatrix.compress_and_factor(A, o)
    S = sb.BLRMatrix.compress_and_factor(A, o, admissible=adm)
    assert W.dense_tiles == 0
    assert S.dense_tiles == 2 * (nb - 1)
    assert S.nonzeros > W.nonzeros
    X = np.random.default_rng(0).standard_normal((n, 2))
    Y = A @ X
    es, ew = rel(S.solve(Y), X), rel(W.solve(Y), X)
    assert es <= 1e2 * tol and es <= 2 * ew + 1e-14        # fewer approximations, no less accurate
    L = sb.BLRMatrix.compress_and_factor(A, o, admissible=adm, factor_algorithm=sb.BLR_LL)
    assert rel(L.solve(Y), S.solve(Y)) <= 1e-12
    with pytest.raises(RuntimeError):
        sb.BLRMatrix.compress_and_factor(A, o, admissible=np.ones((3, 3)))     # wrong size


def test_blr_matches_reference_golden(built):
    """Against the committed outputs of the reference's BLRMatrix<double> on the
    same matrix (tests/golden/blr_toeplitz_1024.npz): full factorization + solve
    and the partially factored front.  Both are approximations at rel_tol: they
    agree to the compression tolerance (north star: 10 * eps_compress), and the
    engine's error against the exact answer is no worse than the reference's."""
    import os
    from conftest import GOLDEN
    sb = built
    g = np.load(os.path.join(GOLDEN, "blr_toeplitz_1024.npz"))
    rank, nnz, ntiles, n1, leaf = (int(v) for v in g["info"])
    tol = float(g["tol"][0])
    n = g["Y"].shape[0]
    A = toeplitz(n) + 2.0 * np.eye(n)
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    B = sb.BLRMatrix.compress_and_factor(A, o)
    assert B.tiles == ntiles and abs(B.rank - rank) <= 3
    assert abs(B.nonzeros - nnz) <= 0.1 * nnz
    xs = B.solve(g["Y"])
    x_exact = np.linalg.solve(A, g["Y"])
    assert rel(xs, g["X"]) <= 10 * tol
    assert rel(xs, x_exact) <= max(10 * rel(g["X"], x_exact), 1e-12)
    F, S = sb.BLRMatrix.construct_and_partial_factor(A[:n1, :n1], A[:n1, n1:], A[n1:, :n1], A[n1:, n1:], o)
    assert rel(S @ g["R"], g["SR"]) <= 10 * tol and rel(S.T @ g["R"], g["STR"]) <= 10 * tol
    f = F.partial_forward_solve(g["b"])
    assert rel(f[n1:], g["fwd"][n1:]) <= 10 * tol           # b_upd - A21 A11^{-1} b_sep: pivot independent
    mid = np.vstack([f[:n1], np.linalg.solve(S, f[n1:])])
    assert rel(F.partial_backward_solve(mid), g["bwd"]) <= 10 * tol


def test_blr_from_element_blocks_and_dense(built):
    """The extract_t forms (reference BLRMatrix.hpp:104-112, 223-232): compress,
    compress_and_factor and construct_and_partial_factor from a block callback,
    and BLRMatrix::dense() of a compressed matrix."""
    sb = built
    n, leaf, tol = 900, 128, 1e-6
    A = toeplitz(n) + 2.0 * np.eye(n)
    A[np.triu_indices(n, 200)] *= 0.5                      # non-symmetric
    ncalls = []

    def block(I, J):
        ncalls.append((len(I), len(J)))
        return A[np.ix_(I, J)]

    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=leaf)
    C_ = sb.BLRMatrix.from_element_blocks(n, block, o, factor=False)
    assert len(ncalls) == C_.tiles ** 2                     # one call per tile pair
    assert rel(C_.dense(), A) <= 1e2 * tol
    x = np.random.default_rng(0).standard_normal((n, 2))
    assert rel(C_.mult(x, "T"), A.T @ x) <= 1e2 * tol
    F = sb.BLRMatrix.from_element_blocks(n, block, o, factor=True)
    D = sb.BLRMatrix.compress_and_factor(A, o)
    y = A @ x
    assert rel(F.solve(y), x) <= 1e2 * tol
    assert rel(F.solve(y), D.solve(y)) <= 1e-12            # same matrix, same algorithm
    n1 = 500
    P, S = sb.BLRMatrix.construct_and_partial_factor_elements(n1, n - n1, block, o)
    P2, S2 = sb.BLRMatrix.construct_and_partial_factor(A[:n1, :n1], A[:n1, n1:], A[n1:, :n1], A[n1:, n1:], o)
    assert rel(S, S2) <= 1e-13
    assert rel(S, A[n1:, n1:] - A[n1:, :n1] @ np.linalg.solve(A[:n1, :n1], A[:n1, n1:])) <= 1e2 * tol


def test_blr_caller_given_tiles(built):
    """The `tiles` vectors of BLRMatrix(A, tiles, admissible, opts) and of
    construct_and_partial_factor(..., tiles1, tiles2, ...) (reference BLRMatrix.hpp:91-101,
    FrontBLR.cpp:422-433: the separator's own cluster tree leaves, ragged sizes)."""
    sb = built
    n1, n2, tol = 900, 500, 1e-6
    A, A11, A12, A21, A22 = _front(n1, n2)
    t1, t2 = [100, 250, 37, 313, 200], [300, 200]
    o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=tol, abs_tol=1e-12, leaf_size=128)
    F, S = sb.BLRMatrix.construct_and_partial_factor(A11, A12, A21, A22, o, tiles1=t1, tiles2=t2)
    assert F.tiles == len(t1) + len(t2)
    assert rel(S, A22 - A21 @ np.linalg.solve(A11, A12)) <= 1e2 * tol
    G = sb.BLRMatrix.compress_and_factor(A11, o, tiles=t1)
    assert G.tiles == len(t1)
    B = np.random.default_rng(5).standard_normal((n1, 2))
    assert rel(A11 @ G.solve(B), B) <= 1e2 * tol
    for bad in ([100, 250], [900, 0], [1000, -100]):    # wrong sum, empty tile, negative tile
        with pytest.raises((RuntimeError, ValueError)):
            sb.BLRMatrix.compress_and_factor(A11, o, tiles=bad)
