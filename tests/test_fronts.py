"""CPU check of the synthetic BLR front generator (BASELINE configs[3] input):
the separable sine-transform formula must reproduce the dense Schur complement
of the 7-point Laplacian onto the middle plane."""
import numpy as np

from strumpack_b200.fronts import laplacian_root_front, plane_bisection_order


def _laplacian_schur(k):
    n = k ** 3
    idx = lambda x, y, z: (x * k + y) * k + z
    A = np.zeros((n, n))
    for x in range(k):
        for y in range(k):
            for z in range(k):
                i = idx(x, y, z)
                A[i, i] = 6
                for d in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
                    X, Y, Z = x + d[0], y + d[1], z + d[2]
                    if 0 <= X < k and 0 <= Y < k and 0 <= Z < k:
                        A[i, idx(X, Y, Z)] = -1
    sep = [idx(x, y, k // 2) for x in range(k) for y in range(k)]
    rest = sorted(set(range(n)) - set(sep))
    return A[np.ix_(sep, sep)] - A[np.ix_(sep, rest)] @ np.linalg.solve(
        A[np.ix_(rest, rest)], A[np.ix_(rest, sep)])


def test_front_formula_matches_dense_schur_complement():
    for k in (5, 7):
        S = _laplacian_schur(k)
        F, perm = laplacian_root_front(k, leaf=8)
        assert sorted(perm.tolist()) == list(range(k * k))
        assert np.abs(F - S[np.ix_(perm, perm)]).max() < 1e-13
        Ft, _ = laplacian_root_front(k, leaf=8, device="cpu")      # the torch path used at full size
        assert np.abs(Ft.numpy() - F).max() < 1e-13


def test_bisection_clusters_are_compact():
    k, leaf = 33, 64
    perm = plane_bisection_order(k, leaf)
    pts = np.stack(np.divmod(perm, k), 1)
    # consecutive chunks of the final clusters have a small bounding box
    sizes = []
    i = 0
    from strumpack_b200.fronts import plane_bisection_order as _  # noqa: F401
    while i < len(perm):
        j = min(i + leaf // 2, len(perm))
        box = pts[i:j].max(0) - pts[i:j].min(0) + 1
        sizes.append(box.prod())
        i = j
    assert np.median(sizes) <= 4 * (leaf // 2)
