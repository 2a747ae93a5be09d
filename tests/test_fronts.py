"""CPU check of the synthetic BLR front generator (BASELINE configs[3] input):
the separable sine-transform formula must reproduce the dense Schur complement
of the 7-point Laplacian onto the middle plane."""
import numpy as np
import pytest

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


@pytest.mark.gpu
def test_device_resident_extend_add(built):
    """F(I[y], I[x]) += CB(y, x) for the left and the right child of several parent
    fronts in one call, everything resident on the device (reference
    extend_add_kernel, src/sparse/fronts/FrontCUDA.cu:111-148); integer-valued data,
    so the result is exact whatever the order of the additions."""
    import ctypes as C
    import torch
    sb = built
    rng = np.random.default_rng(11)
    fronts, keep, expect = [], [], []
    for d1, d2, n1, n2 in [(40, 70, 55, 31), (33, 0, 20, 0), (17, 90, 0, 64), (128, 200, 300, 257)]:
        n = d1 + d2
        F = rng.integers(-5, 6, size=(n, n)).astype(np.float64)
        blocks = [np.asfortranarray(F[:d1, :d1]), np.asfortranarray(F[:d1, d1:]),
                  np.asfortranarray(F[d1:, :d1]), np.asfortranarray(F[d1:, d1:])]
        dev = [torch.tensor(b.T.copy(), device="cuda") for b in blocks]      # row-major of the transpose = column-major
        ref = F.copy()
        desc = sb.SB200FrontAssemble()
        desc.F11, desc.F12, desc.F21, desc.F22 = [t.data_ptr() for t in dev]
        desc.d1, desc.d2 = d1, d2
        for side, nc in ((1, n1), (2, n2)):
            if nc == 0:
                continue
            I = np.sort(rng.choice(n, size=nc, replace=False)).astype(np.int32)
            CB = rng.integers(-9, 10, size=(nc, nc)).astype(np.float64)
            ref[np.ix_(I, I)] += CB
            dI = torch.tensor(I, device="cuda")
            dCB = torch.tensor(CB.T.copy(), device="cuda")
            keep += [dI, dCB]
            setattr(desc, f"CB{side}", dCB.data_ptr())
            setattr(desc, f"I{side}", dI.data_ptr())
            setattr(desc, f"dCB{side}", nc)
        fronts.append(desc)
        keep += dev
        expect.append((dev, ref, d1))
    arr = (sb.SB200FrontAssemble * len(fronts))(*fronts)
    darr = torch.tensor(np.frombuffer(bytes(arr), dtype=np.uint8).copy(), device="cuda")
    rc = sb.lib().SB200_d_front_extend_add_device(len(fronts), C.c_void_p(darr.data_ptr()), 300,
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    for dev, ref, d1 in expect:
        got = np.block([[dev[0].cpu().numpy().T, dev[1].cpu().numpy().T],
                        [dev[2].cpu().numpy().T, dev[3].cpu().numpy().T]])
        assert np.array_equal(got, ref)
