"""The ULV leaf QR kernels on their own (C-ABI hook SB200_debug_qr_batch): blocked
Householder QR of the first k columns of an m x naug block, reflectors applied to
all naug columns -- what `W0.LQ` + the three GEMMs with `Q` compute in the
reference (src/HSS/HSSMatrix.factor.hpp:122-141).

Checked against the definition: with V_p (unit lower trapezoid) and T_p of every
16-column panel read back from the output,  prod_p (I - V_p T_p^T V_p^T) A  must
be [R | Q^T A_aug] exactly as stored, R upper triangular, column norms preserved
(Q orthogonal).  Tolerance: 1e-13 relative (fp64 Householder, m <= 256).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(256, 231, 281), (256, 256, 300), (256, 240, 240), (64, 50, 78), (128, 100, 141), (250, 221, 271),
          (255, 230, 270), (101, 88, 120), (40, 17, 30), (16, 9, 20), (9, 5, 9), (256, 16, 64), (256, 7, 40),
          (200, 33, 34)]


def check(A, k, out, T):
    m, naug = A.shape
    B = A.copy()
    for c0 in range(0, k, 16):
        jb = min(16, k - c0)
        V = np.tril(out[c0:, c0:c0 + jb], -1)
        V[np.arange(jb), np.arange(jb)] = 1.0
        Tp = np.triu(T[:jb, c0:c0 + jb])
        assert np.array_equal(Tp, T[:jb, c0:c0 + jb]), "T not upper triangular"
        B[c0:, :] -= V @ (Tp.T @ (V.T @ B[c0:, :]))
    scale = np.linalg.norm(A)
    R = np.triu(out[:, :k])
    e_r = np.linalg.norm(B[:, :k] - R) / scale
    e_aug = np.linalg.norm(B[:, k:] - out[:, k:]) / scale if naug > k else 0.0
    e_nrm = np.max(np.abs(np.linalg.norm(B, axis=0) - np.linalg.norm(A, axis=0))) / scale
    return e_r, e_aug, e_nrm


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("m,k,naug", SHAPES)
def test_leaf_qr_kernel(built, m, k, naug, variant):
    sb = built
    rng = np.random.default_rng(m * 1000 + k)
    A = np.asfortranarray(rng.standard_normal((m, naug)))
    out, T, _ = sb.debug_qr_batch(A, k, count=3, variant=variant)
    e_r, e_aug, e_nrm = check(A, k, out, T)
    assert e_r < 1e-13 and e_aug < 1e-13 and e_nrm < 1e-13, (e_r, e_aug, e_nrm)


def test_leaf_qr_kernels_agree_and_scale(built):
    """Both kernels on a batch larger than one wave: same R up to rounding, and a
    rank-deficient block (zero columns -> tau = 0 reflectors)."""
    sb = built
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.standard_normal((256, 281)))
    A[:, 40:44] = 0.0
    A[:, 100] = A[:, 3]
    o1, T1, _ = sb.debug_qr_batch(A, 231, count=700, variant=1)
    o0, T0, _ = sb.debug_qr_batch(A, 231, count=700, variant=0)
    assert max(check(A, 231, o1, T1)) < 1e-12
    R1, R0 = np.triu(o1[:, :231]), np.triu(o0[:, :231])
    # columns after the exactly dependent one have rounding-level pivots: compare the well-determined part
    assert np.linalg.norm(np.abs(R1[:100, :100]) - np.abs(R0[:100, :100])) / np.linalg.norm(R0[:100, :100]) < 1e-12
