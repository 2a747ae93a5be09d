"""GPU parity tests (through the C ABI): HSS apply / ULV factor / ULV solve
vs the golden vectors produced by the reference and vs the CPU oracle.

Tolerances: the reference's own test demands ||B - H(H\\B)||/||B|| <= 1e-12
(test/test_HSS_seq.cpp:39,247-250) in fp64; apply is the same arithmetic in a
different summation order -> 1e-13."""
import os

import numpy as np
import pytest

from conftest import CASES, GOLDEN, have_ref
from oracle import hss_file, hss_oracle as ho

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def sb(built):
    import torch
    assert torch.cuda.is_available()
    return built


@pytest.mark.parametrize("case", CASES)
def test_apply_matches_reference(sb, case):
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    assert (H.rows, H.cols, H.rank, H.levels) == tuple(g["info"][:4])
    assert rel(H.mult(g["x"]), g["y"]) < 1e-13
    assert rel(H.mult(g["x"], "T"), g["yt"]) < 1e-13
    assert rel(H.mult(g["x"][:, 0]), g["y"][:, :1]) < 1e-13   # single rhs
    assert H.launches > 0


@pytest.mark.parametrize("case", CASES)
def test_ulv_matches_reference(sb, case):
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    H.factor()
    xs = H.solve(g["y"])
    assert rel(xs, g["xs"]) < 1e-10
    assert rel(H.mult(xs), g["y"]) < 1e-12      # reference's acceptance bound
    assert H.flops("factor") == g["flops"][0]


def test_solve_before_factor_fails(sb, capfd):
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, CASES[0] + ".hss"))
    with pytest.raises(RuntimeError):
        H.solve(np.ones(H.rows))
    assert "Operation failed" in capfd.readouterr().err


def test_from_generators_and_dense(sb):
    nodes, _ = hss_file.read_hss(os.path.join(GOLDEN, CASES[1] + ".hss"))
    H = sb.HSSMatrix.from_generators(nodes)
    A = H.dense()
    assert rel(A, ho.to_dense(nodes)) < 1e-13


def test_shift_and_write(sb, tmp_path):
    path = os.path.join(GOLDEN, CASES[0] + ".hss")
    nodes, _ = hss_file.read_hss(path)
    H = sb.HSSMatrix.read(path)
    x = np.random.default_rng(3).standard_normal((H.rows, 2))
    y0 = H.mult(x)
    H.shift(2.5)
    assert rel(H.mult(x), y0 + 2.5 * x) < 1e-13
    H.factor()
    y = H.mult(x)
    assert rel(H.solve(y), x) < 1e-10
    out = tmp_path / "shifted.hss"
    H.write(out)
    n2, _ = hss_file.read_hss(out)
    leaf = next(i for i, n in enumerate(nodes) if n.leaf)
    assert np.allclose(n2[leaf].D, nodes[leaf].D + 2.5 * np.eye(nodes[leaf].rows))


def test_device_resident_path(sb):
    import torch
    path = os.path.join(GOLDEN, CASES[2] + ".hss")
    g = np.load(os.path.join(GOLDEN, CASES[2] + ".npz"))
    H = sb.HSSMatrix.read(path)
    xT = torch.tensor(g["x"].T.copy(), device="cuda", dtype=torch.float64)
    yT = torch.empty_like(xT)
    H.mult_device(xT, yT)
    H.factor_device()
    bT = yT.clone()
    H.solve_device(bT)
    torch.cuda.synchronize()
    assert rel(yT.cpu().numpy().T, g["y"]) < 1e-13
    assert rel(bT.cpu().numpy().T, g["xs"]) < 1e-10


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,leaf,tol,kind", [(4096, 128, 1e-4, "T"),
                                              (2000, 100, 1e-8, "U"),
                                              (1000, 16, 1e-2, "T"),
                                              (97, 128, 1e-4, "T")])
def test_live_reference_toeplitz(sb, tmp_path, n, leaf, tol, kind):
    """BASELINE configs[0] (4096 Toeplitz, leaf 128, tol 1e-4) and the ragged /
    tiny-leaf / single-node edge cases of test/CMakeLists.txt:57-159."""
    from oracle import ref
    ref.set_num_threads(8)
    R = ref.RefHSS.toeplitz(n, kind, f"--hss_leaf_size {leaf} --hss_rel_tol {tol}")
    p = tmp_path / "h.hss"
    R.write(p)
    H = sb.HSSMatrix.read(p)
    x = np.random.default_rng(n).standard_normal((n, 2))
    y_ref = R.mult(x)
    assert rel(H.mult(x), y_ref) < 1e-13
    assert rel(H.mult(x, "T"), R.mult(x, trans=True)) < 1e-13
    ref.flops_reset()
    R.factor()
    assert H.flops("factor") == ref.flops()["ulv_factor"]
    H.factor()
    xs = H.solve(y_ref)
    assert rel(xs, R.solve(y_ref)) < 1e-9
    assert rel(H.mult(xs), y_ref) < 1e-12


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_live_reference_gauss_ragged(sb, tmp_path):
    """Gaussian kernel, 2-means clustering -> ragged leaves."""
    from oracle import ref
    ref.set_num_threads(8)
    pts = np.random.default_rng(42).random((2, 8192))
    R = ref.RefHSS.gauss(pts, 0.1, 1.0, "--hss_leaf_size 256 --hss_rel_tol 1e-4")
    p = tmp_path / "h.hss"
    R.write(p)
    H = sb.HSSMatrix.read(p)
    x = np.random.default_rng(0).standard_normal((8192, 3))
    y_ref = R.mult(x)
    assert rel(H.mult(x), y_ref) < 1e-13
    R.factor()
    H.factor()
    xs = H.solve(y_ref)
    assert rel(xs, R.solve(y_ref)) < 1e-9
    assert rel(H.mult(xs), y_ref) < 1e-12


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("nrhs", [4, 16, 37])
def test_apply_many_rhs_gemm_path(sb, case, nrhs):
    """>= 4 right-hand sides run the GEMM-shaped kernels (fp64 tensor pipe, 16
    vectors per CTA): same result as the oracle's apply_fwd/apply_bwd restatement
    and as the one-column kernels (reference multi-rhs case, SURVEY 8d C2 'N x 64')."""
    nodes, _ = hss_file.read_hss(os.path.join(GOLDEN, case + ".hss"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    x = np.random.default_rng(nrhs).standard_normal((H.rows, nrhs))
    y = H.mult(x)
    assert rel(y, ho.apply(nodes, x)) < 1e-13
    yt = H.mult(x, "T")
    assert rel(yt, ho.apply(nodes, x, trans=True)) < 1e-13
    # column by column through the single-rhs kernels
    y1 = np.hstack([H.mult(x[:, j]) for j in range(0, nrhs, 5)])
    assert rel(y[:, ::5], y1) < 1e-13


@pytest.mark.parametrize("nrhs", [1, 3, 20])
def test_apply_hss_with_beta_and_split_solve(sb, nrhs):
    """apply_HSS(op, A, B, beta, C) (HSSMatrix.cpp:419-435) and the
    forward_solve / backward_solve pair (HSSMatrix.solve.hpp:52-66)."""
    case = CASES[2]
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    nodes, _ = hss_file.read_hss(os.path.join(GOLDEN, case + ".hss"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    rng = np.random.default_rng(nrhs)
    x = rng.standard_normal((H.rows, nrhs))
    c = rng.standard_normal((H.rows, nrhs))
    for tr in ("N", "T"):
        ref = ho.apply(nodes, x, trans=(tr == "T")) - 0.75 * c
        assert rel(H.apply(x, -0.75, c, tr), ref) < 1e-13
    assert rel(H.apply(x), ho.apply(nodes, x)) < 1e-13
    H.factor()
    y = H.mult(x)
    H.forward_solve(y)
    xs = H.backward_solve()
    assert rel(xs, H.solve(y)) < 1e-14       # same kernels, same order
    assert rel(xs, x) < 1e-10


@pytest.mark.parametrize("case", CASES)
def test_extract_sub_blocks(sb, case):
    """extract / extract_add / get against dense(H) from the oracle, as
    test/test_HSS_seq.cpp:204-233 does (random index sets, sorted and unsorted,
    with repetitions, crossing every level of the tree)."""
    nodes, _ = hss_file.read_hss(os.path.join(GOLDEN, case + ".hss"))
    A = ho.to_dense(nodes)
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    n = H.rows
    rng = np.random.default_rng(11)
    for nI, nJ in ((1, 1), (7, 5), (64, 80), (n, 3)):
        I = rng.integers(0, n, nI) if nI < n else np.arange(n)
        J = rng.integers(0, n, nJ)
        B = H.extract(I, J)
        ref = A[np.ix_(I, J)]
        assert np.abs(B - ref).max() <= 1e-13 * max(1.0, np.abs(A).max())
        C0 = rng.standard_normal((nI, nJ))
        assert np.abs(H.extract(I, J, add_to=C0) - (C0 + ref)).max() <= 1e-12
    assert abs(H.get(3, n - 2) - A[3, n - 2]) <= 1e-13
    assert abs(H.get(5, 5) - A[5, 5]) <= 1e-13
    with pytest.raises(RuntimeError):
        H.extract([n], [0])


def test_ulv_factor_export(sb):
    """HSSMatrix::ULV() accessor (reference HSSMatrix.hpp:497/511): the factor arena
    comes back to the host with the advertised size; before factor() it is an error."""
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, CASES[0] + ".hss"))
    with pytest.raises(RuntimeError):
        H.ulv_data()
    H.factor()
    f, t = H.ulv_data()
    assert f.size + t.size == H.factor_nonzeros and t.size > 0
    assert np.all(np.isfinite(f)) and np.all(np.isfinite(t))
    assert np.count_nonzero(f) > 0.25 * f.size and np.count_nonzero(t) > 0


def _random_hss(rng, leaf_rows, ru_leaf, rv_leaf, ru_in, rv_in):
    """A random 2-level HSS (root, two inner nodes, four leaves) given by its
    generators, with independent U and V ranks (nonsymmetric)."""
    N = hss_file.Node

    def ipiv(n):
        return np.array([rng.integers(i, n) + 1 for i in range(n)], dtype=np.int32)

    def E(rows, rank):
        return np.asfortranarray(rng.standard_normal((rows - rank, rank)) / np.sqrt(rows))

    z = np.zeros((0, 0), order="F")
    zi = np.zeros(0, dtype=np.int32)
    nodes = [None] * 7
    m = leaf_rows
    nodes[0] = N(0, -1, 4 * m, 4 * m, 0, 0, 0, 0, zi, z, zi, z, z,
                 rng.standard_normal((ru_in, rv_in)), rng.standard_normal((ru_in, rv_in)), [1, 4])
    for a, base in ((1, 0), (4, 2)):
        nodes[a] = N(a, 0, 2 * m, 2 * m, ru_in, 2 * ru_leaf, rv_in, 2 * rv_leaf,
                     ipiv(2 * ru_leaf), E(2 * ru_leaf, ru_in), ipiv(2 * rv_leaf), E(2 * rv_leaf, rv_in), z,
                     rng.standard_normal((ru_leaf, rv_leaf)), rng.standard_normal((ru_leaf, rv_leaf)),
                     [a + 1, a + 2], base * m, base * m)
        for q in (1, 2):
            nodes[a + q] = N(a + q, a, m, m, ru_leaf, m, rv_leaf, m, ipiv(m), E(m, ru_leaf), ipiv(m),
                             E(m, rv_leaf), np.asfortranarray(rng.standard_normal((m, m)) + m * np.eye(m)), z, z,
                             [], (base + q - 1) * m, (base + q - 1) * m)
    return nodes


@pytest.mark.parametrize("ru_in,rv_in", [(5, 29), (29, 5), (12, 12)])
def test_nonsymmetric_ranks_transposed_multi_rhs(sb, ru_in, rv_in):
    """V ranks much larger than U ranks (and the reverse) at the root's children
    and below: shared-memory staging of the apply kernels is sized by the larger
    side (round-1 advisor finding), N and T products, 1 and many right-hand sides."""
    rng = np.random.default_rng(5)
    nodes = _random_hss(rng, 48, ru_leaf=7 if ru_in < rv_in else 20, rv_leaf=20 if ru_in < rv_in else 7,
                        ru_in=ru_in, rv_in=rv_in)
    H = sb.HSSMatrix.from_generators(nodes)
    A = ho.to_dense(nodes)
    for s in (1, 3, 4, 9, 37):
        x = rng.standard_normal((H.rows, s))
        assert rel(H.mult(x), A @ x) < 1e-13
        assert rel(H.mult(x, "T"), A.T @ x) < 1e-13
    assert rel(H.dense(), A) < 1e-13
    H.factor()
    b = rng.standard_normal((H.rows, 5))
    assert rel(A @ H.solve(b), b) < 1e-10


def test_two_handles_from_two_threads(sb):
    """One handle = one stream + one lock (re-entrant per object,
    thread-safe across objects): two host threads drive two matrices through
    the host-pointer C ABI at the same time, and two threads share ONE matrix;
    every result equals the single-threaded one."""
    import threading
    # arrays materialised up front: an NpzFile reads lazily from one zip handle, which threads must not share
    g = dict(np.load(os.path.join(GOLDEN, CASES[0] + ".npz")))
    H1 = sb.HSSMatrix.read(os.path.join(GOLDEN, CASES[0] + ".hss"))
    H2 = sb.HSSMatrix.read(os.path.join(GOLDEN, CASES[2] + ".hss"))
    g2 = dict(np.load(os.path.join(GOLDEN, CASES[2] + ".npz")))
    H1.factor()
    H2.factor()
    errs = []

    def work(H, gg, reps):
        try:
            for _ in range(reps):
                assert rel(H.mult(gg["x"]), gg["y"]) < 1e-13
                assert rel(H.solve(gg["y"]), gg["xs"]) < 1e-10
        except Exception as e:      # surfaced below (a thread's assertion is otherwise lost)
            errs.append(e)

    ts = [threading.Thread(target=work, args=a) for a in ((H1, g, 40), (H2, g2, 40), (H1, g, 40))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[0]
