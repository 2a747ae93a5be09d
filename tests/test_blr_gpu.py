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
