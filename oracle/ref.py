"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product.

ctypes view of ``oracle/_ref/libsb200_ref.so``: the UNMODIFIED reference
(pghysels/STRUMPACK @ cfba574) HSS/BLR sources compiled by ``oracle/Makefile``
plus the C-ABI driver ``oracle/ref_driver.cpp``.  Every numeric result that
comes out of this module was computed by the reference's own code
(``HSSMatrix<double>`` compress / mult / factor / solve; ``BLRMatrix<double>``
compress_and_factor / solve).

Allowed importers: ``tests/``, ``__graft_entry__.smoke()``, and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.
"""
import ctypes as C
import glob
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsb200_ref.so")
_OBLAS_DIR = ("/opt/prime-rl/.venv/lib/python3.12/site-packages/"
              "opencv_python_headless.libs")

_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError(
            f"{_SO} missing: run `make -C oracle` where /root/reference exists")
    # BLAS threads off: the reference parallelises over the tree with OpenMP
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    # OpenBLAS from the image's opencv wheel + its private Fortran runtime;
    # libsb200_ref.so carries a DT_RPATH to that directory, the preload is a
    # belt-and-braces for loaders that ignore it.
    for pat in ("libquadmath*.so*", "libgfortran*.so*", "libopenblasp*.so"):
        for f in sorted(glob.glob(os.path.join(_OBLAS_DIR, pat))):
            try:
                C.CDLL(f, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
    L = C.CDLL(_SO)
    vp, i, d, cp = C.c_void_p, C.c_int, C.c_double, C.c_char_p
    dp = np.ctypeslib.ndpointer(np.float64, flags="F_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(np.int32)
    lp = np.ctypeslib.ndpointer(np.int64)
    sig = {
        "ref_set_num_threads": (None, [i]),
        "ref_get_max_threads": (i, []),
        "ref_flops_reset": (None, []),
        "ref_flops_total": (C.c_longlong, []),
        "ref_flops_ulv_factor": (C.c_longlong, []),
        "ref_flops_hss_solve": (C.c_longlong, []),
        "ref_hss_toeplitz": (vp, [i, i, cp]),
        "ref_hss_dense": (vp, [i, i, dp, i, cp]),
        "ref_hss_gauss": (vp, [i, i, dp, d, d, cp, ip]),
        "ref_hss_read": (vp, [cp]),
        "ref_hss_write": (i, [vp, cp]),
        "ref_hss_info": (None, [vp, lp]),
        "ref_hss_print_info": (None, [vp]),
        "ref_hss_mult": (None, [vp, i, i, dp, i, dp, i]),
        "ref_hss_factor": (None, [vp]),
        "ref_hss_solve": (None, [vp, i, dp, i]),
        "ref_hss_shift": (None, [vp, d]),
        "ref_hss_dense_out": (None, [vp, dp, i]),
        "ref_hss_get": (d, [vp, i, i]),
        "ref_hss_destroy": (None, [vp]),
        "ref_hss_partial_factor": (None, [vp]),
        "ref_hss_schur_sizes": (None, [vp, lp]),
        "ref_hss_schur_get": (None, [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "ref_hss_schur_product_direct": (None, [vp, i, dp, i, dp, i, dp, i]),
        "ref_hss_partial_forward": (None, [vp, i, dp, i, dp]),
        "ref_hss_partial_x_rows": (i, [vp]),
        "ref_hss_partial_x": (None, [vp, dp, i]),
        "ref_hss_partial_backward": (None, [vp, i, dp, i]),
        "ref_blr_factor_dense": (vp, [i, dp, i, cp, C.POINTER(i), C.c_void_p]),
        "ref_blr_solve": (None, [vp, i, dp, i]),
        "ref_blr_info": (None, [vp, lp]),
        "ref_blr_destroy": (None, [vp]),
        "ref_blr_partial_factor": (vp, [i, i, dp, i, dp, i, dp, i, dp, i, cp]),
        "ref_blr_partial_info": (None, [vp, lp]),
        "ref_blr_partial_forward": (None, [vp, i, dp, i]),
        "ref_blr_partial_backward": (None, [vp, i, dp, i]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _f(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


class RefHSS:
    """A reference ``strumpack::HSS::HSSMatrix<double>``."""

    def __init__(self, handle, perm1=None, pts=None):
        self._h = handle
        self.perm1 = perm1   # reference's 1-based clustering permutation
        self.pts = pts       # points in the permuted (HSS) ordering, d x n

    # -- constructors ------------------------------------------------------
    @classmethod
    def toeplitz(cls, n, kind="T", args=""):
        return cls(lib().ref_hss_toeplitz(n, ord(kind), args.encode()))

    @classmethod
    def dense(cls, A, args=""):
        A = _f(A)
        return cls(lib().ref_hss_dense(A.shape[0], A.shape[1], A, A.shape[0],
                                       args.encode()))

    @classmethod
    def gauss(cls, pts, h, lam, args=""):
        """pts: d x n (one point per column).  Returns the matrix in the
        reference's permuted ordering; ``.pts`` holds the permuted points."""
        pts = np.asfortranarray(np.array(pts, dtype=np.float64))
        d, n = pts.shape
        perm1 = np.zeros(n, dtype=np.int32)
        h_ = lib().ref_hss_gauss(n, d, pts, h, lam, args.encode(), perm1)
        return cls(h_, perm1, pts)

    @classmethod
    def read(cls, path):
        return cls(lib().ref_hss_read(str(path).encode()))

    # -- queries -----------------------------------------------------------
    def info(self):
        out = np.zeros(8, dtype=np.int64)
        lib().ref_hss_info(self._h, out)
        keys = ("rows", "cols", "rank", "levels", "nonzeros",
                "factor_nonzeros", "memory", "is_compressed")
        return dict(zip(keys, (int(v) for v in out)))

    def write(self, path):
        lib().ref_hss_write(self._h, str(path).encode())

    def to_dense(self):
        inf = self.info()
        A = np.zeros((inf["rows"], inf["cols"]), order="F")
        lib().ref_hss_dense_out(self._h, A, A.shape[0])
        return A

    # -- the hot path, as the reference computes it -------------------------
    def mult(self, x, trans=False):
        x = _f(x)
        inf = self.info()
        ny = inf["cols"] if trans else inf["rows"]
        y = np.zeros((ny, x.shape[1]), order="F")
        lib().ref_hss_mult(self._h, int(trans), x.shape[1], x, x.shape[0],
                           y, ny)
        return y

    def factor(self):
        lib().ref_hss_factor(self._h)

    def solve(self, b):
        x = _f(b).copy(order="F")
        lib().ref_hss_solve(self._h, x.shape[1], x, x.shape[0])
        return x

    def shift(self, sigma):
        lib().ref_hss_shift(self._h, float(sigma))

    # -- Schur complement of the (0,0) block, as FrontHSS drives it ---------
    def partial_factor(self):
        """partial_factor + Schur_update (+ ThetaVhatC_or_VhatCPhiC); returns
        dict(Theta, DUB01, Phi, Vhat)."""
        L = lib()
        L.ref_hss_partial_factor(self._h)
        z = np.zeros(8, dtype=np.int64)
        L.ref_hss_schur_sizes(self._h, z)
        out = {k: np.zeros((int(z[2 * q]), int(z[2 * q + 1])), order="F")
               for q, k in enumerate(("Theta", "DUB01", "Phi", "Vhat"))}
        L.ref_hss_schur_get(self._h, *(out[k].ctypes.data for k in ("Theta", "DUB01", "Phi", "Vhat")))
        return out

    def schur_product_direct(self, R):
        R = _f(R)
        n1, c = R.shape
        Sr = np.zeros((n1, c), order="F")
        Sc = np.zeros((n1, c), order="F")
        lib().ref_hss_schur_product_direct(self._h, c, R, n1, Sr, n1, Sc, n1)
        return Sr, Sc

    def partial_forward_solve(self, b0, rv0):
        b0 = _f(b0)
        red = np.zeros((rv0, b0.shape[1]), order="F")
        lib().ref_hss_partial_forward(self._h, b0.shape[1], b0, b0.shape[0], red)
        self._ps = b0.shape
        return red

    def partial_x(self, new=None):
        m0 = lib().ref_hss_partial_x_rows(self._h)
        if new is not None:
            x = _f(new).copy(order="F")
            lib().ref_hss_partial_x(self._h, x, 1)
            return x
        x = np.zeros((m0, self._ps[1]), order="F")
        lib().ref_hss_partial_x(self._h, x, 0)
        return x

    def partial_backward_solve(self):
        n0, s = self._ps
        x0 = np.zeros((n0, s), order="F")
        lib().ref_hss_partial_backward(self._h, s, x0, n0)
        return x0

    def close(self):
        if self._h:
            lib().ref_hss_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefBLR:
    """Reference ``BLRMatrix<double>::compress_and_factor`` (weak
    admissibility, tiles from ``ClusterTree(n).refine(leaf)``)."""

    def __init__(self, A, args=""):
        A = _f(A)
        n = A.shape[0]
        nt = C.c_int(0)
        tiles = np.zeros(n, dtype=np.int32)
        self._h = lib().ref_blr_factor_dense(
            n, A, n, args.encode(), C.byref(nt), tiles.ctypes.data)
        self.tiles = tiles[:nt.value].copy()
        self.n = n

    def info(self):
        out = np.zeros(4, dtype=np.int64)
        lib().ref_blr_info(self._h, out)
        return dict(zip(("rows", "cols", "rank", "nonzeros"),
                        (int(v) for v in out)))

    def solve(self, b):
        x = _f(b).copy(order="F")
        lib().ref_blr_solve(self._h, x.shape[1], x, x.shape[0])
        return x

    def close(self):
        if self._h:
            lib().ref_blr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefBLRFront:
    """Reference ``BLRMatrix<double>::construct_and_partial_factor`` on a front
    [A11 A12; A21 A22] (weak admissibility, ClusterTree tiles); ``.S`` is the
    Schur complement the reference leaves in A22."""

    def __init__(self, A11, A12, A21, A22, args=""):
        A11, A12, A21 = _f(A11), _f(A12), _f(A21)
        self.S = _f(A22).copy(order="F")
        self.n1, self.n2 = A11.shape[0], self.S.shape[0]
        self._h = lib().ref_blr_partial_factor(
            self.n1, self.n2, A11, self.n1, A12, self.n1, A21, self.n2,
            self.S, self.n2, args.encode())

    def ranks(self):
        out = np.zeros(3, dtype=np.int64)
        lib().ref_blr_partial_info(self._h, out)
        return tuple(int(v) for v in out)

    def partial_forward_solve(self, b):
        x = _f(b).copy(order="F")
        lib().ref_blr_partial_forward(self._h, x.shape[1], x, x.shape[0])
        return x

    def partial_backward_solve(self, y):
        x = _f(y).copy(order="F")
        lib().ref_blr_partial_backward(self._h, x.shape[1], x, x.shape[0])
        return x

    def close(self):
        if self._h:
            lib().ref_blr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def flops_reset():
    lib().ref_flops_reset()


def flops():
    L = lib()
    return {"total": L.ref_flops_total(),
            "ulv_factor": L.ref_flops_ulv_factor(),
            "hss_solve": L.ref_flops_hss_solve()}


def set_num_threads(t):
    lib().ref_set_num_threads(int(t))


def max_threads():
    return lib().ref_get_max_threads()
