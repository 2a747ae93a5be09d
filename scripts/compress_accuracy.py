"""Accuracy of the GPU HSS construction vs the exact kernel matrix, next to the
reference's own HSSMatrix(kernel) construction (oracle/_ref) on the same points."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.setrecursionlimit(100000)
import torch
import strumpack_b200 as sb

def exact(p, x, h, lam):
    n = p.shape[1]
    P = torch.tensor(p.T.copy(), device="cuda"); X = torch.tensor(x, device="cuda")
    Y = torch.empty_like(X)
    for r0 in range(0, n, 4096):
        d2 = ((P[r0:r0 + 4096, None, :] - P[None, :, :]) ** 2).sum(-1) if n <= 32768 else torch.cdist(P[r0:r0 + 4096], P).pow(2)
        Y[r0:r0 + 4096] = torch.exp(-d2 / (2 * h * h)) @ X
    return (Y + lam * X).cpu().numpy()

rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
h, lam, tol = 0.1, 1.0, 1e-4
do_ref = "--ref" in sys.argv
for n in [int(a) for a in sys.argv[1:] if a.isdigit()]:
    pts = np.random.default_rng(42).random((2, n))
    x = np.random.default_rng(1).standard_normal((n, 2))
    o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=tol, abs_tol=1e-10, leaf_size=256)
    t0 = time.perf_counter()
    H, perm, p = sb.HSSMatrix.from_kernel(pts, sb.KERNEL_GAUSS, h, lam, o)
    tc = time.perf_counter() - t0
    e = rel(H.mult(x), exact(p, x, h, lam))
    print(f"N={n} ours: err {e:.3e} rank {H.rank} nnz {H.nonzeros/1e6:.2f}M levels {H.levels} compress {tc:.2f}s", flush=True)
    if do_ref and n <= 131072:
        from oracle import ref
        ref.set_num_threads(min(32, os.cpu_count()))
        t0 = time.perf_counter()
        R = ref.RefHSS.gauss(pts, h, lam, f"--hss_leaf_size 256 --hss_rel_tol {tol}")
        tr = time.perf_counter() - t0
        pr = R.pts       # points in the reference's permuted (HSS) ordering
        info = R.info()
        if pr is not None:
            er = rel(R.mult(x), exact(pr, x, h, lam))
            print(f"N={n} reference: err {er:.3e} info {info} compress {tr:.2f}s", flush=True)
        else:
            print(f"N={n} reference: info {info} compress {tr:.2f}s (no permuted points accessor)", flush=True)
