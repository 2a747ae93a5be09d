"""Compression accuracy at the BASELINE sizes with the tolerance-scaled sample sizes."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import strumpack_b200 as sb
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))

def gauss_rows_err(H, p, x, h, lam, nrows=4096, seed=0):
    """|| (K x - H x)[rows] || / || (K x)[rows] || on a random row sample, K exact (fp64, on the GPU)"""
    n = p.shape[1]
    rows = np.sort(np.random.default_rng(seed).choice(n, size=min(nrows, n), replace=False))
    P = torch.tensor(p.T.copy(), device="cuda")
    X = torch.tensor(x, device="cuda")
    Y = torch.zeros(len(rows), x.shape[1], dtype=torch.float64, device="cuda")
    R = P[torch.tensor(rows, device="cuda")]
    for c0 in range(0, n, 65536):
        d2 = torch.cdist(R, P[c0:c0 + 65536]).pow(2)
        Y += torch.exp(-d2 / (2 * h * h)) @ X[c0:c0 + 65536]
    Y += lam * X[torch.tensor(rows, device="cuda")]
    y = H.mult(x)
    return rel(y[rows], Y.cpu().numpy())

if __name__ == "__main__":
  for n in (65536, 1 << 20):
      pts = np.random.default_rng(42).random((2, n))
      o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=1e-4, abs_tol=1e-10, leaf_size=256)
      t0 = time.perf_counter()
      H, perm, p = sb.HSSMatrix.from_kernel(pts, sb.KERNEL_GAUSS, 0.1, 1.0, o)
      t = time.perf_counter() - t0
      x = np.random.default_rng(1).standard_normal((n, 2))
      print(json.dumps({"case": f"gauss 2-D h=0.1 N={n} tol 1e-4", "err_rows": gauss_rows_err(H, p, x, 0.1, 1.0),
                        "rank": H.rank, "nnz": H.nonzeros, "compress_s": t}), flush=True)
      H.close()
