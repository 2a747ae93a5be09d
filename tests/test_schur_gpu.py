"""GPU parity tests (through the C ABI) of the Schur-complement path the
reference's HSS fronts use: partial_factor, Schur_update,
Schur_product_direct / _indirect and the partial forward / backward solves
(reference HSSMatrix.factor.hpp:44-50, HSSMatrix.Schur.hpp:40-215,
HSSMatrix.solve.hpp:133-152; caller src/sparse/fronts/FrontHSS.cpp:385-495).

Compared with the golden vectors the reference itself produced
(tests/golden/make_golden_schur.py), with the live reference library when it
is present, and with the dense Schur complement.  fp64; the reference has no
test of this path, the bounds are those of its ULV test (1e-12 residual,
test/test_HSS_seq.cpp:247-250) with one digit of slack for the extra products."""
import os

import numpy as np
import pytest

from conftest import CASES, GOLDEN, have_ref
from oracle import hss_file, hss_oracle as ho

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def sb(built):
    import torch
    assert torch.cuda.is_available()
    return built


def _dense_schur(case):
    nodes, _ = hss_file.read_hss(os.path.join(GOLDEN, case + ".hss"))
    A = ho.to_dense(nodes)
    n0 = nodes[nodes[0].ch[0]].rows
    S = A[n0:, n0:] - A[n0:, :n0] @ np.linalg.solve(A[:n0, :n0], A[:n0, n0:])
    return A, n0, S


@pytest.mark.parametrize("case", CASES)
def test_schur_update_matches_reference(sb, case):
    g = np.load(os.path.join(GOLDEN, case + "_schur.npz"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    H.partial_factor()
    Theta, DUB01, Phi = H.schur_update()
    Vhat = H.vhat()
    assert [list(a.shape) for a in (Theta, DUB01, Phi, Vhat)] == g["sizes"].tolist()
    assert rel(Theta, g["Theta"]) < 1e-13
    assert rel(Vhat.T @ DUB01, g["VtD"]) < 1e-11
    assert rel(Vhat.T @ Phi.T, g["VtPhiT"]) < 1e-11
    A, n0, S = _dense_schur(case)
    assert rel(A[n0:, n0:] - Theta @ Vhat.T @ Phi.T, S) < 1e-11


@pytest.mark.parametrize("case", CASES)
def test_schur_product_direct_matches_reference(sb, case):
    g = np.load(os.path.join(GOLDEN, case + "_schur.npz"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    H.partial_factor()
    Theta, DUB01, Phi = H.schur_update()
    Sr, Sc = H.schur_product_direct(g["R"])                      # device copies of Theta/DUB01/Phi
    assert rel(Sr, g["Sr"]) < 1e-11 and rel(Sc, g["Sc"]) < 1e-11
    Sr2, Sc2 = H.schur_product_direct(g["R"][:, :1], Theta, DUB01, Phi)   # host copies, 1 column
    assert rel(Sr2, g["Sr"][:, :1]) < 1e-11 and rel(Sc2, g["Sc"][:, :1]) < 1e-11
    # many columns: the GEMM-shaped sweeps
    _, _, S = _dense_schur(case)
    R = np.random.default_rng(7).standard_normal((S.shape[0], 20))
    Sr3, Sc3 = H.schur_product_direct(R)
    assert rel(Sr3, S @ R) < 1e-11 and rel(Sc3, S.T @ R) < 1e-11


@pytest.mark.parametrize("case", CASES)
def test_schur_product_indirect(sb, case):
    """Sr = Sr1 - H10 R0 - H10 H00^{-1} H01 R1 with Sr1 = (H [R0; R1])_1 gives
    S R1 (HSSMatrix.Schur.hpp:139-147); the reference keeps this path behind
    indirect_sampling (off by default), so the check is the identity itself."""
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    A, n0, S = _dense_schur(case)
    H.partial_factor()
    H.schur_update()
    rng = np.random.default_rng(11)
    for c in (1, 6):
        R0 = rng.standard_normal((n0, c))
        R1 = rng.standard_normal((A.shape[0] - n0, c))
        R = np.vstack([R0, R1])
        Sr1 = (A @ R)[n0:]
        Sc1 = (A.T @ R)[n0:]
        Sr, Sc = H.schur_product_indirect(R0, R1, Sr1, Sc1)
        assert rel(Sr, S @ R1) < 1e-10
        assert rel(Sc, S.T @ R1) < 1e-10


@pytest.mark.parametrize("case", CASES)
def test_partial_solves_match_reference(sb, case):
    g = np.load(os.path.join(GOLDEN, case + "_schur.npz"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    H.partial_factor()
    Theta, DUB01, Phi = H.schur_update()
    red = H.partial_forward_solve(g["b0"])
    assert rel(red, g["red"]) < 1e-11
    x0 = H.partial_backward_solve()
    assert rel(x0, g["x0"]) < 1e-11
    # FrontHSS::bwd_solve_node: x -= Phi^H y_upd between the two sweeps
    H.partial_forward_solve(g["b0"])
    H.partial_x(H.partial_x() - Phi.T @ g["yupd"])
    assert rel(H.partial_backward_solve(), g["x0u"]) < 1e-11
    # single right-hand side
    red1 = H.partial_forward_solve(g["b0"][:, :1])
    assert rel(red1, g["red"][:, :1]) < 1e-11
    assert rel(H.partial_backward_solve(), g["x0"][:, :1]) < 1e-11


def test_front_elimination_through_schur(sb):
    """The way FrontHSS uses the pieces (FrontHSS.cpp:385-410,440-500): eliminate
    block 0 of [F11 F12; F21 F22] x = b and solve the Schur system densely."""
    case = "gauss2d_1024_leaf64"
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    A, n0, S = _dense_schur(case)
    H.partial_factor()
    Theta, DUB01, Phi = H.schur_update()
    Vhat = H.vhat()
    rng = np.random.default_rng(5)
    b = rng.standard_normal((A.shape[0], 2))
    red = H.partial_forward_solve(b[:n0])
    bupd = b[n0:] - Theta @ red                               # fwd_solve_node
    Sd = A[n0:, n0:] - Theta @ (Vhat.T @ Phi.T)               # what the parent front receives
    yupd = np.linalg.solve(Sd, bupd)
    H.partial_x(H.partial_x() - Phi.T @ yupd)                 # bwd_solve_node
    x = np.vstack([H.partial_backward_solve(), yupd])
    assert rel(A @ x, b) < 1e-10


def test_full_factor_after_partial_and_errors(sb, capfd):
    case = CASES[0]
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    with pytest.raises(RuntimeError):
        H.schur_update()                                      # before partial_factor
    assert "Operation failed" in capfd.readouterr().err
    H.partial_factor()
    with pytest.raises(RuntimeError):
        H.solve(g["y"])                                       # a partial factorization is not a full one
    H.factor()
    assert rel(H.solve(g["y"]), g["xs"]) < 1e-10
    with pytest.raises(RuntimeError):
        H.schur_update()                                      # factor() replaced the partial factors
    H.partial_factor()
    gs = np.load(os.path.join(GOLDEN, case + "_schur.npz"))
    assert rel(H.schur_update()[0], gs["Theta"]) < 1e-13


@pytest.mark.skipif(not have_ref(), reason="reference library not built")
def test_schur_live_reference_two_level(sb):
    """A matrix whose root children are leaves (one-level tree): the sub-roots
    are leaves, D0 is the leaf block itself."""
    from oracle import ref
    n = 200
    Href = ref.RefHSS.toeplitz(n, "T", "--hss_leaf_size 128 --hss_rel_tol 1e-8")
    path = "/tmp/sb200_two_level.hss"
    Href.write(path)
    d = Href.partial_factor()
    H = sb.HSSMatrix.read(path)
    assert H.levels == 2
    H.partial_factor()
    Theta, DUB01, Phi = H.schur_update()
    Vhat = H.vhat()
    assert rel(Theta, d["Theta"]) < 1e-13
    assert rel(Vhat.T @ DUB01, d["Vhat"].T @ d["DUB01"]) < 1e-11
    assert rel(Vhat.T @ Phi.T, d["Vhat"].T @ d["Phi"].T) < 1e-11
    R = np.random.default_rng(2).standard_normal((Theta.shape[0], 4))
    Sr, Sc = H.schur_product_direct(R)
    Sr_ref, Sc_ref = Href.schur_product_direct(R)
    assert rel(Sr, Sr_ref) < 1e-11 and rel(Sc, Sc_ref) < 1e-11
    b0 = np.random.default_rng(3).standard_normal((n - Theta.shape[0], 2))
    red = H.partial_forward_solve(b0)
    assert rel(red, Href.partial_forward_solve(b0, Vhat.shape[1])) < 1e-11
    assert rel(H.partial_backward_solve(), Href.partial_backward_solve()) < 1e-11


def test_schur_device_resident(sb):
    """The device-pointer forms (SB200_d_hss_*_device) on torch's current stream."""
    import torch
    case = CASES[2]
    g = np.load(os.path.join(GOLDEN, case + "_schur.npz"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    RT = torch.tensor(np.ascontiguousarray(g["R"].T), device="cuda")
    SrT, ScT = H.schur_device(RT)
    torch.cuda.synchronize()
    assert rel(SrT.cpu().numpy().T, g["Sr"]) < 1e-11
    assert rel(ScT.cpu().numpy().T, g["Sc"]) < 1e-11
