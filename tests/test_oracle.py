"""CPU tests: pin the numpy oracle (oracle/hss_oracle.py).

(1) against the committed golden vectors (outputs of the reference itself,
    tests/golden/make_golden.py), (2) against the live reference library
    oracle/_ref when it is present, including the reference's own acceptance
    checks of test/test_HSS_seq.cpp (compression error <= 1e2*tol, :143-152;
    ULV residual <= 1e-12, :235-250)."""
import os

import numpy as np
import pytest

from conftest import CASES, GOLDEN, have_ref
from oracle import hss_file, hss_oracle as ho


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(case):
    nodes, ver = hss_file.read_hss(os.path.join(GOLDEN, case + ".hss"))
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    assert tuple(ver) == (8, 0, 0)
    assert nodes[0].rows == g["info"][0] and nodes[0].cols == g["info"][1]
    assert max(max(n.U_rank, n.V_rank) for n in nodes) == g["info"][2]
    # apply, both directions: bit-for-bit the same algorithm -> ~1e-15
    assert rel(ho.apply(nodes, g["x"]), g["y"]) < 1e-13
    assert rel(ho.apply(nodes, g["x"], trans=True), g["yt"]) < 1e-13
    # ULV factor + solve
    f = ho.factor(nodes)
    xs = ho.solve(nodes, f, g["y"])
    assert rel(xs, g["xs"]) < 1e-11
    # reference's own acceptance: ||B - H (H\B)|| / ||B|| <= 1e-12
    assert rel(ho.apply(nodes, xs), g["y"]) < 1e-12


def test_basis_identities():
    nodes, _ = hss_file.read_hss(os.path.join(GOLDEN, CASES[0] + ".hss"))
    nd = next(n for n in nodes if n.leaf)
    U = ho.basis_dense(nd.Pu, nd.Eu)
    b = np.random.default_rng(0).standard_normal((nd.U_rank, 2))
    assert np.allclose(ho.basis_apply(nd.Pu, nd.Eu, b), U @ b)
    c = np.random.default_rng(1).standard_normal((nd.U_rows, 2))
    assert np.allclose(ho.basis_applyC(nd.Pu, nd.Eu, c), U.T @ c)
    # interpolative: U contains the identity on the skeleton rows
    g = hss_file.ipiv_to_gather(nd.Pu)
    assert np.allclose(U[g[:nd.U_rank], :], np.eye(nd.U_rank))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_oracle_matches_live_reference(tmp_path):
    from oracle import ref
    ref.set_num_threads(2)
    n = 768
    H = ref.RefHSS.toeplitz(n, "T", "--hss_leaf_size 48 --hss_rel_tol 1e-5")
    p = tmp_path / "h.hss"
    H.write(p)
    nodes, _ = hss_file.read_hss(p)
    x = np.random.default_rng(5).standard_normal((n, 2))
    assert rel(ho.apply(nodes, x), H.mult(x)) < 1e-13
    H.factor()
    y = H.mult(x)
    f = ho.factor(nodes)
    assert rel(ho.solve(nodes, f, y), H.solve(y)) < 1e-11
    # test_HSS_seq.cpp:143-152 compression check on the dense matrix
    i = np.arange(n)
    A = 1.0 / (1.0 + np.abs(i[:, None] - i[None, :]))
    assert rel(ho.to_dense(nodes), A) < 1e2 * 1e-5


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_reference_selftests_pass():
    """The oracle build of the reference passes the reference's own tests."""
    import subprocess
    from oracle import ref
    d = os.path.dirname(ref._SO)
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="2")
    out = subprocess.run([os.path.join(d, "test_HSS_seq"), "T", "200",
                          "--hss_leaf_size", "16", "--hss_rel_tol", "1e-5"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-500:]
    out = subprocess.run([os.path.join(d, "test_BLR_seq"), "512"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-500:]


@pytest.mark.parametrize("case", CASES)
def test_schur_restatement_matches_reference_golden(case):
    """partial_factor / Schur_update / Schur_product_direct: the
    numpy restatement reproduces what the reference computed through the call
    sequence of FrontHSS (tests/golden/make_golden_schur.py).  Compared on the
    quantities that do not depend on the ULV's orthogonal basis."""
    def rel(a, b):   # the upper-triangular case has Theta = 0 exactly
        return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    nodes, _ = hss_file.read_hss(os.path.join(GOLDEN, case + ".hss"))
    g = np.load(os.path.join(GOLDEN, case + "_schur.npz"))
    f = ho.partial_factor(nodes)
    Theta, DUB01, Phi = ho.schur_update(nodes, f)
    assert [list(a.shape) for a in (Theta, DUB01, Phi, f.Vhat)] == g["sizes"].tolist()
    assert rel(Theta, g["Theta"]) < 1e-13
    assert rel(f.Vhat.T @ DUB01, g["VtD"]) < 1e-12
    assert rel(f.Vhat.T @ Phi.T, g["VtPhiT"]) < 1e-12
    Sr, Sc = ho.schur_product_direct(nodes, f, Theta, DUB01, Phi, g["R"])
    assert rel(Sr, g["Sr"]) < 1e-13 and rel(Sc, g["Sc"]) < 1e-13
    # and the identity the fronts rely on: S = H11 - H10 H00^{-1} H01
    A = ho.to_dense(nodes)
    n0 = A.shape[0] - Theta.shape[0]
    S = A[n0:, n0:] - A[n0:, :n0] @ np.linalg.solve(A[:n0, :n0], A[:n0, n0:])
    assert rel(A[n0:, n0:] - Theta @ f.Vhat.T @ Phi.T, S) < 1e-12
    assert rel(g["Sr"], S @ g["R"]) < 1e-12 and rel(g["Sc"], S.T @ g["R"]) < 1e-12
    # partial forward / backward solves of child(0), with the update in between
    st, red = ho.partial_forward_solve(nodes, f, g["b0"])
    assert rel(red, g["red"]) < 1e-12
    assert rel(ho.partial_backward_solve(nodes, st), g["x0"]) < 1e-12
    st, _ = ho.partial_forward_solve(nodes, f, g["b0"])
    st.xs[st.root] = st.xs[st.root] - Phi.T @ g["yupd"]
    assert rel(ho.partial_backward_solve(nodes, st), g["x0u"]) < 1e-12
    assert rel(g["Theta"] @ g["red"], A[n0:, :n0] @ g["x0"]) < 1e-12
    assert rel(g["x0"], np.linalg.solve(A[:n0, :n0], g["b0"])) < 1e-12


def _blr_matrix(n):
    i = np.arange(n)
    return 1.0 / (1.0 + np.abs(i[:, None] - i[None, :])) + 2.0 * np.eye(n)


def test_blr_restatement_matches_reference_golden():
    """oracle/blr_oracle.py (RRQR tiles, right-looking tile LU, block substitution,
    partial factorization of a front) against outputs of the reference's
    BLRMatrix<double> (tests/golden/make_golden_blr.py).  Same algorithm, same
    LAPACK primitives: agreement to roundoff although both are 1e-6 approximations."""
    from oracle import blr_oracle as bo
    g = np.load(os.path.join(GOLDEN, "blr_toeplitz_1024.npz"))
    rank, nnz, ntiles, n1, leaf = (int(v) for v in g["info"])
    tol = float(g["tol"][0])
    A = _blr_matrix(g["Y"].shape[0])
    F = bo.compress_and_factor(A, leaf, tol)
    assert F.nb == ntiles and F.rank() == rank
    assert rel(F.solve(g["Y"]), g["X"]) < 1e-12
    assert rel(A @ g["X"], g["Y"]) <= 1e2 * tol              # the reference's own bound (test_BLR_seq.cpp:192)
    P = bo.construct_and_partial_factor(A[:n1, :n1], A[:n1, n1:], A[n1:, :n1], A[n1:, n1:], leaf, tol)
    S = P.schur()
    assert rel(S @ g["R"], g["SR"]) < 1e-12 and rel(S.T @ g["R"], g["STR"]) < 1e-12
    assert P.rank() == int(max(g["ranks"]))
    f = P.forward(g["b"])
    assert rel(f, g["fwd"]) < 1e-12
    mid = np.vstack([f[:n1], np.linalg.solve(S, f[n1:])])
    assert rel(P.backward(mid), g["bwd"]) < 1e-12
    S_exact = A[n1:, n1:] - A[n1:, :n1] @ np.linalg.solve(A[:n1, :n1], A[:n1, n1:])
    assert rel(g["SR"], S_exact @ g["R"]) <= 1e2 * tol


@pytest.mark.skipif(not have_ref(), reason="reference library not built")
def test_blr_restatement_matches_live_reference():
    from oracle import blr_oracle as bo, ref
    n, leaf, tol = 600, 64, 1e-4
    A = _blr_matrix(n)
    A[64:128, 320:384] += 0.05 * np.random.default_rng(3).standard_normal((64, 64))   # a tile that stays dense
    R = ref.RefBLR(A, f"--blr_leaf_size {leaf} --blr_rel_tol {tol}")
    F = bo.compress_and_factor(A, leaf, tol)
    assert list(R.tiles) == list(np.diff(F.off))
    assert F.rank() == R.info()["rank"]
    assert sum(t.nonzeros() for t in F.t.values()) + sum(l[0].size for l in F.lu.values()) == R.info()["nonzeros"]
    Y = np.random.default_rng(4).standard_normal((n, 2))
    assert rel(F.solve(Y), R.solve(Y)) < 1e-11


@pytest.mark.skipif(not have_ref(), reason="reference library not built")
def test_reference_factor_algorithms_agree():
    """What the engine's treatment of BLRFactorAlgorithm rests on:
    in the reference itself LL produces exactly the RL factors (the same tile
    updates in another order), and COMB / STAR (LUAR accumulation) differ from RL
    far below the compression tolerance, with the same ranks."""
    from oracle import ref
    n, leaf, tol = 1024, 128, 1e-6
    A = _blr_matrix(n)
    Y = np.random.default_rng(0).standard_normal((n, 2))
    res = {}
    for alg in ("RL", "LL", "Comb", "Star"):
        B = ref.RefBLR(A, f"--blr_leaf_size {leaf} --blr_rel_tol {tol} --blr_factor_algorithm {alg}")
        res[alg] = (B.solve(Y), B.info())
    assert np.array_equal(res["LL"][0], res["RL"][0])
    for alg in ("Comb", "Star"):
        assert rel(res[alg][0], res["RL"][0]) < 1e-2 * tol
        assert res[alg][1]["rank"] == res["RL"][1]["rank"]
