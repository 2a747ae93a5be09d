#!/usr/bin/env python
"""BLR LU timing (BASELINE.json configs[3]-like): compress_and_factor + solve of a
dense N x N matrix, tile 256, tol 1e-4, through the C ABI (host matrix in, H2D
inside the timing), next to the reference's CPU path on the same matrix when
oracle/_ref is present.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import strumpack_b200 as sb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
with_ref = len(sys.argv) > 2 and sys.argv[2] == "ref"
i = np.arange(n)
A = np.asfortranarray(1.0 / (1.0 + np.abs(i[:, None] - i[None, :])))
o = sb.default_options(type=sb.SP_TYPE_BLR, rel_tol=1e-4, abs_tol=1e-12, leaf_size=256)
B = sb.BLRMatrix.compress_and_factor(A, o)      # warm-up (allocations, module load)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    B = sb.BLRMatrix.compress_and_factor(A, o)
    ts.append(time.perf_counter() - t0)
import torch
dA = torch.tensor(A.T.copy(), device="cuda")   # symmetric here; column-major = row-major transpose
torch.cuda.synchronize()
td = []
for _ in range(3):
    t0 = time.perf_counter()
    Bd = sb.BLRMatrix.compress_and_factor_device(dA, o)
    torch.cuda.synchronize()
    td.append(time.perf_counter() - t0)
del Bd, dA
X = np.random.default_rng(0).standard_normal((n, 10))
Y = A @ X
t0 = time.perf_counter(); Xs = B.solve(Y); t_solve = time.perf_counter() - t0
err = float(np.linalg.norm(Xs - X) / np.linalg.norm(X))
out = {"workload": f"BLR compress_and_factor (RL, weak admissibility) + solve(10 rhs), Toeplitz N={n}, tile 256, tol 1e-4",
       "factor_s": min(ts), "factor_device_resident_s": min(td), "solve_s": t_solve, "rank": B.rank, "tiles": B.tiles,
       "nonzeros_frac": B.nonzeros / (n * n), "rel_err": err, "launches": B.launches}
if with_ref:
    from oracle import ref
    ref.set_num_threads(min(os.cpu_count(), 32))
    t0 = time.perf_counter(); R = ref.RefBLR(A, "--blr_leaf_size 256 --blr_rel_tol 1e-4"); tr = time.perf_counter() - t0
    xr = R.solve(Y)
    out["reference_cpu"] = {"factor_s": tr, "threads": min(os.cpu_count(), 32), "rank": R.info()["rank"],
                            "rel_err": float(np.linalg.norm(xr - X) / np.linalg.norm(X))}
print(json.dumps(out))
