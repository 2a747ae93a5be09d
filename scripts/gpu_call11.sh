#!/bin/bash
mkdir -p gpurun_out
for w in 0 1; do echo "WEIGHTED=$w"; SB200_COMPRESS_WEIGHTED=$w timeout 900 python scripts/compress_accuracy.py 16384 65536 262144 2>&1 | tail -3; done
python - <<'PY'
import os, sys, numpy as np, time
sys.path.insert(0, os.getcwd())
import strumpack_b200 as sb
def fftprod(x):
    n = x.shape[0]; c = 1.0/(1.0+np.arange(n)); col = np.concatenate([c,[0.0],c[:0:-1]])
    return np.fft.irfft(np.fft.rfft(col)*np.fft.rfft(np.concatenate([x,np.zeros(n)])))[:n]
for w in (0, 1):
    os.environ["SB200_COMPRESS_WEIGHTED"] = str(w)
    for n, tol in ((32768, 1e-6), (262144, 1e-6), (262144, 1e-4)):
        o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=tol, abs_tol=1e-12, leaf_size=256)
        t0 = time.perf_counter()
        H, _, _ = sb.HSSMatrix.from_kernel(np.zeros((1, n)), sb.KERNEL_TOEPLITZ_INVDIST, 1.0, 0.0, o)
        tc = time.perf_counter() - t0
        x = np.random.default_rng(2).standard_normal(n); y = fftprod(x)
        e = np.linalg.norm(H.mult(x)[:,0]-y)/np.linalg.norm(y)
        print(f"toeplitz weighted={w} N={n} tol={tol}: err {e:.3e} rank {H.rank} nnz {H.nonzeros/1e6:.1f}M compress {tc:.2f}s", flush=True)
PY
