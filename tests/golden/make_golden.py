"""Generates the golden fixtures in this directory by running the UNMODIFIED
reference (pghysels/STRUMPACK @ cfba574, built by oracle/Makefile from
/root/reference) in the build container.  Run once:  python tests/golden/make_golden.py

Each case writes
  <name>.hss   the reference's own HSSMatrix<double>::write dump (generators)
  <name>.npz   inputs and the reference's outputs on them:
               x (N x 3), y = H x, yt = H^T x, xs = H \\ y   (HSSMatrix::mult /
               factor / solve), info (rows, cols, rank, levels, nonzeros) and
               the reference's flop counters for factor and a 3-rhs solve.
Problem definitions follow test/test_HSS_seq.cpp:69-91 (Toeplitz 'T'/'U') and
examples/dense/KernelRegression (Gauss kernel, 2-means clustering).
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402


def emit(name, H, seed):
    inf = H.info()
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((inf["cols"], 3))
    H.write(os.path.join(HERE, name + ".hss"))
    y = H.mult(x)
    yt = H.mult(x, trans=True)
    ref.flops_reset()
    H.factor()
    ff = ref.flops()["ulv_factor"]
    xs = H.solve(y)
    fs = ref.flops()["hss_solve"]
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), x=x, y=y, yt=yt, xs=xs,
        info=np.array([inf[k] for k in ("rows", "cols", "rank", "levels",
                                        "nonzeros")], dtype=np.int64),
        flops=np.array([ff, fs], dtype=np.int64),
        pts=(H.pts if H.pts is not None else np.zeros((0, 0))))
    print(name, inf, "factor flops", ff, "solve flops (3 rhs)", fs,
          "resid", np.linalg.norm(xs - x) / np.linalg.norm(x))


if __name__ == "__main__":
    ref.set_num_threads(4)
    emit("toeplitz_512_leaf64",
         ref.RefHSS.toeplitz(512, "T", "--hss_leaf_size 64 --hss_rel_tol 1e-6"), 1)
    emit("utoeplitz_300_leaf32",
         ref.RefHSS.toeplitz(300, "U", "--hss_leaf_size 32 --hss_rel_tol 1e-8"), 2)
    pts = np.random.default_rng(42).random((2, 1024))
    emit("gauss2d_1024_leaf64",
         ref.RefHSS.gauss(pts, 0.1, 1.0, "--hss_leaf_size 64 --hss_rel_tol 1e-4"), 3)
