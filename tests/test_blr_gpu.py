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
