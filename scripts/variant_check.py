"""Development check: QR kernel variants (env switches) vs the default kernel.
Compares the solves and residuals (the T arenas differ in layout between panel widths).
usage: variant_check.py NAME=VAL[,NAME=VAL] ..."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strumpack_b200 as sb

variants = [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:]]
ALL = sorted({k for v in variants for k in v})

def build(env, maker):
    for k in ALL:
        os.environ.pop(k, None)
    os.environ.update(env)
    H = maker()
    H.factor()
    return H

def compare(name, maker, n, nrhs=2):
    b = np.random.default_rng(0).standard_normal((n, nrhs))
    H0 = build({}, maker)
    x0 = H0.solve(b)
    f0, _ = H0.ulv_data()
    for env in variants:
        H = build(env, maker)
        x = H.solve(b)
        f, _ = H.ulv_data()
        res = np.linalg.norm(H.mult(x) - b) / np.linalg.norm(b)
        print(f"{name} {env}: fact maxdiff {np.abs(f - f0).max() / np.abs(f0).max():.2e} nan {int(np.isnan(f).sum())} "
              f"x diff {np.linalg.norm(x - x0) / np.linalg.norm(x0):.2e}  resid {res:.2e}", flush=True)

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for g, n in (("toeplitz_512_leaf64", 512), ("utoeplitz_300_leaf32", 300), ("gauss2d_1024_leaf64", 1024)):
    p = os.path.join(root, "tests", "golden", g + ".hss")
    compare(g, lambda p=p: sb.HSSMatrix.read(p), n)
for (d, h, n, leaf) in ((2, 0.1, 16384, 256), (3, 0.2, 4096, 200), (2, 0.1, 5000, 100)):
    pts = np.random.default_rng(42).random((d, n))
    o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=1e-4, abs_tol=1e-10, leaf_size=leaf)
    compare(f"gauss{d}d_{n}_leaf{leaf}", lambda: sb.HSSMatrix.from_kernel(pts, sb.KERNEL_GAUSS, h, 1.0, o)[0], n)
