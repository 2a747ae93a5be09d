"""Development check: left-looking leaf QR (SB200_QR_LL) vs the right-looking kernel.
Compares the ULV factor arenas entry by entry and the solves."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strumpack_b200 as sb

def build(mode, maker):
    os.environ["SB200_QR_LL"] = str(mode)
    H = maker()
    H.factor()
    return H

def compare(name, maker, n, nrhs=2):
    b = np.random.default_rng(0).standard_normal((n, nrhs))
    H0 = build(0, maker)
    f0, t0 = H0.ulv_data()
    x0 = H0.solve(b)
    for mode in (1, 2):
        H = build(mode, maker)
        f, t = H.ulv_data()
        x = H.solve(b)
        df = np.abs(f - f0); dt = np.abs(t - t0)
        sf = np.abs(f0).max(); st = max(np.abs(t0).max(), 1e-300)
        bad = int(np.argmax(df)); badt = int(np.argmax(dt))
        nanf = int(np.isnan(f).sum()); nant = int(np.isnan(t).sum())
        res = np.linalg.norm(H.mult(x) - b) / np.linalg.norm(b)
        print(f"{name} LL={mode}: fact maxdiff {df.max()/sf:.2e} (at {bad}/{f.size}, nan {nanf})  "
              f"T maxdiff {dt.max()/st:.2e} (at {badt}/{t.size}, nan {nant})  "
              f"x diff {np.linalg.norm(x-x0)/np.linalg.norm(x0):.2e}  resid {res:.2e}", flush=True)

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for g, n in (("toeplitz_512_leaf64", 512), ("utoeplitz_300_leaf32", 300), ("gauss2d_1024_leaf64", 1024)):
    p = os.path.join(root, "tests", "golden", g + ".hss")
    compare(g, lambda p=p: sb.HSSMatrix.read(p), n)
for (d, h, n, leaf) in ((2, 0.1, 16384, 256), (3, 0.2, 4096, 200), (2, 0.1, 5000, 100)):
    pts = np.random.default_rng(42).random((d, n))
    o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=1e-4, abs_tol=1e-10, leaf_size=leaf)
    compare(f"gauss{d}d_{n}_leaf{leaf}", lambda: sb.HSSMatrix.from_kernel(pts, sb.KERNEL_GAUSS, h, 1.0, o)[0], n)
