import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import strumpack_b200 as sb
from r2_accuracy import gauss_rows_err
n = 1 << 20
pts = np.random.default_rng(42).random((2, n))
x = np.random.default_rng(1).standard_normal((n, 2))
for near, far in [(192, 256), (384, 512), (768, 256), (192, 1024)]:
    os.environ["SB200_SAMPLE_NEAR"], os.environ["SB200_SAMPLE_FAR"] = str(near), str(far)
    o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=1e-4, abs_tol=1e-10, leaf_size=256)
    t0 = time.perf_counter()
    H, perm, p = sb.HSSMatrix.from_kernel(pts.copy(), sb.KERNEL_GAUSS, 0.1, 1.0, o)
    t = time.perf_counter() - t0
    print(json.dumps({"case": f"gauss N={n}", "near": near, "far": far, "err_rows": gauss_rows_err(H, p, x, 0.1, 1.0),
                      "rank": H.rank, "nnz": H.nonzeros, "compress_s": t}), flush=True)
    H.close()
