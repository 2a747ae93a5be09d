"""Accuracy of the sampled-ID HSS construction vs sample sizes (configs[2]: Toeplitz 262144, tol 1e-6;
configs[1]: Gauss 65536, tol 1e-4)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import strumpack_b200 as sb

def toeplitz_fft_product(x):
    n = x.shape[0]
    c = 1.0 / (1.0 + np.arange(n))
    col = np.concatenate([c, [0.0], c[:0:-1]])
    return np.fft.irfft(np.fft.rfft(col) * np.fft.rfft(np.concatenate([x, np.zeros(n)])))[:n]

rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
n = 262144
x = np.random.default_rng(2).standard_normal(n)
y = toeplitz_fft_product(x)
for near, far in [(96, 128), (192, 256), (384, 512), (96, 512), (384, 128)]:
    os.environ["SB200_SAMPLE_NEAR"], os.environ["SB200_SAMPLE_FAR"] = str(near), str(far)
    o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=1e-6, abs_tol=1e-12, leaf_size=256)
    t0 = time.perf_counter()
    H, _, _ = sb.HSSMatrix.from_kernel(np.zeros((1, n)), sb.KERNEL_TOEPLITZ_INVDIST, 1.0, 0.0, o)
    t = time.perf_counter() - t0
    print(json.dumps({"case": "toeplitz 262144 tol 1e-6", "near": near, "far": far, "err": rel(H.mult(x)[:, 0], y),
                      "rank": H.rank, "nnz": H.nonzeros, "compress_s": t}), flush=True)
    H.close()
