"""strumpack_b200 -- host-side Python mirror of the engine's C ABI.

The product is ``libstrumpack_b200.so`` (C ABI in ``include/sb200_structured.h``,
the drop-in for the reference's ``src/structured/StructuredMatrix.h``).  This
module is a thin ctypes view of it whose names follow the reference's
``structured::StructuredMatrix`` / ``HSS::HSSMatrix`` interface
(reference src/structured/StructuredMatrix.hpp:209-418,
src/HSS/HSSMatrix.hpp:95-511): ``rows, cols, rank, memory, nonzeros, levels,
mult, factor, solve, shift``.

There is no CPU fallback: if the shared library is missing, or no GPU is
visible when a compute entry point is called, the call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SB200_LIB: development override (A/B builds of the engine); default = the in-tree library
_SO = os.environ.get("SB200_LIB") or os.path.join(_HERE, "libstrumpack_b200.so")
_lib = None

SP_TYPE_HSS, SP_TYPE_BLR = 0, 1
# reference BLRFactorAlgorithm (src/BLR/BLROptions.hpp:65)
BLR_COLWISE, BLR_RL, BLR_LL, BLR_COMB, BLR_STAR = 0, 1, 2, 3, 4
KERNEL_GAUSS, KERNEL_LAPLACE, KERNEL_TOEPLITZ_INVDIST = 0, 1, 2
NODE_FIELDS = 16


class CSPOptions(C.Structure):
    """reference StructuredMatrix.h:68-75"""
    _fields_ = [("type", C.c_int), ("rel_tol", C.c_double),
                ("abs_tol", C.c_double), ("leaf_size", C.c_int),
                ("max_rank", C.c_int), ("verbose", C.c_int)]


class SB200FrontAssemble(C.Structure):
    """One parent front + the contribution blocks of its two children (device pointers)."""
    _fields_ = [("F11", C.c_void_p), ("F12", C.c_void_p), ("F21", C.c_void_p), ("F22", C.c_void_p),
                ("d1", C.c_int), ("d2", C.c_int),
                ("CB1", C.c_void_p), ("I1", C.c_void_p), ("dCB1", C.c_int),
                ("CB2", C.c_void_p), ("I2", C.c_void_p), ("dCB2", C.c_int)]


class SB200BLRParams(C.Structure):
    """include/sb200_structured.h: BLROptions members CSPOptions does not carry"""
    _fields_ = [("pivot_threshold", C.c_double), ("factor_algorithm", C.c_int),
                ("admissible", C.c_void_p), ("n_admissible", C.c_int),
                ("tiles1", C.c_void_p), ("n_tiles1", C.c_int),
                ("tiles2", C.c_void_p), ("n_tiles2", C.c_int)]


def _blr_params(pivot_threshold, factor_algorithm, admissible, tiles1=None, tiles2=None):
    p = SB200BLRParams(float(pivot_threshold), int(factor_algorithm), None, 0, None, 0, None, 0)
    keep = []
    if admissible is not None:
        adm = np.asfortranarray(np.asarray(admissible) != 0, dtype=np.int32)
        p.admissible, p.n_admissible = adm.ctypes.data, adm.shape[0]
        keep.append(adm)
    for name, t in (("tiles1", tiles1), ("tiles2", tiles2)):
        if t is not None:
            t = np.ascontiguousarray(t, dtype=np.int32)
            setattr(p, name, t.ctypes.data)
            setattr(p, "n_" + name, t.size)
            keep.append(t)
    return p, keep


# every symbol include/sb200_structured.h declares: name -> (restype, argtypes)
# HSS::ClusteringAlgorithm of the reference (HSSOptions.hpp)
CLUSTER_NATURAL, CLUSTER_TWO_MEANS, CLUSTER_KD_TREE, CLUSTER_PCA, CLUSTER_COBBLE = range(5)

_vp, _i, _d, _ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
_pvp = C.POINTER(C.c_void_p)
_po = C.POINTER(CSPOptions)
SYMBOLS = {
    "SB200_version": (C.c_char_p, []),
    "SB200_fp64_dmma_peak_tflops": (_d, []),
    "SB200_nccl_unique_id": (_i, [_vp]),
    "SB200_d_hss_dist_init": (_i, [_vp, _i, _i, _vp]),
    "SB200_d_hss_dist_mult": (_i, [_vp, C.c_char, _i, _vp, _i, _vp, _i, _vp]),
    "SB200_d_hss_dist_factor": (_i, [_vp, _vp]),
    "SB200_d_hss_dist_solve": (_i, [_vp, _i, _vp, _i, _vp]),
    "SB200_d_front_extend_add_device": (_i, [_i, _vp, _i, _vp]),
    "SB200_debug_qr_batch": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "SP_d_struct_default_options": (None, [_po]),
    "SP_d_struct_destroy": (None, [_pvp]),
    "SP_d_struct_rows": (_i, [_vp]),
    "SP_d_struct_cols": (_i, [_vp]),
    "SP_d_struct_memory": (_ll, [_vp]),
    "SP_d_struct_nonzeros": (_ll, [_vp]),
    "SP_d_struct_rank": (_i, [_vp]),
    "SP_d_struct_from_dense": (_i, [_pvp, _i, _i, _vp, _i, _po]),
    "SP_d_struct_from_elements": (_i, [_pvp, _i, _i, _vp, _po]),
    "SP_d_struct_mult": (_i, [_vp, C.c_char, _i, _vp, _i, _vp, _i]),
    "SP_d_struct_factor": (_i, [_vp]),
    "SP_d_struct_solve": (_i, [_vp, _i, _vp, _i]),
    "SP_d_struct_shift": (_i, [_vp, _d]),
    "SP_s_struct_default_options": (None, [_po]),
    "SP_s_struct_destroy": (None, [_pvp]),
    "SP_s_struct_rows": (_i, [_vp]),
    "SP_s_struct_cols": (_i, [_vp]),
    "SP_s_struct_memory": (_ll, [_vp]),
    "SP_s_struct_nonzeros": (_ll, [_vp]),
    "SP_s_struct_rank": (_i, [_vp]),
    "SP_s_struct_from_dense": (_i, [_pvp, _i, _i, _vp, _i, _po]),
    "SP_s_struct_from_elements": (_i, [_pvp, _i, _i, _vp, _po]),
    "SP_s_struct_mult": (_i, [_vp, C.c_char, _i, _vp, _i, _vp, _i]),
    "SP_s_struct_factor": (_i, [_vp]),
    "SP_s_struct_solve": (_i, [_vp, _i, _vp, _i]),
    "SP_s_struct_shift": (_i, [_vp, C.c_float]),
    "SB200_d_hss_from_element_blocks": (_i, [_pvp, _i, _vp, _vp, _po]),
    "SB200_d_hss_from_element_blocks_ex": (_i, [_pvp, _i, _vp, _vp, _po, _i, _vp, _vp, _i, _vp]),
    "SB200_d_hss_from_dense_tree": (_i, [_pvp, _i, _vp, _i, _po, _i, _vp, _vp]),
    "SB200_d_hss_node_table": (_i, [_vp, _vp]),
    "SB200_d_hss_from_kernel_ex": (_i, [_pvp, _i, _i, _vp, _i, _d, _d, _po, _vp, _i]),
    "SB200_d_hss_from_kernel": (_i, [_pvp, _i, _i, _vp, _i, _d, _d, _po, _vp]),
    "SB200_d_blr_compress_and_factor": (_i, [_pvp, _i, _vp, _i, _po, _d]),
    "SB200_d_blr_compress_and_factor_device": (_i, [_pvp, _i, _vp, _i, _po, _d]),
    "SB200_d_blr_partial_factor": (_i, [_pvp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _po, _d]),
    "SB200_d_blr_partial_factor_device": (_i, [_pvp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _po, _d]),
    "SB200_d_blr_compress_and_factor_ex": (_i, [_pvp, _i, _vp, _i, _po, _vp]),
    "SB200_d_blr_partial_factor_ex": (_i, [_pvp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _po, _vp]),
    "SB200_d_blr_from_element_blocks": (_i, [_pvp, _i, _vp, _vp, _po, _vp, _i]),
    "SB200_d_blr_partial_factor_element_blocks": (_i, [_pvp, _i, _i, _vp, _vp, _vp, _i, _po, _vp]),
    "SB200_d_blr_sep_rows": (_i, [_vp]),
    "SB200_d_blr_partial_forward_solve": (_i, [_vp, _i, _vp, _i]),
    "SB200_d_blr_partial_backward_solve": (_i, [_vp, _i, _vp, _i]),
    "SB200_d_blr_tiles": (_i, [_vp]),
    "SB200_d_blr_dense_tiles": (_i, [_vp]),
    "SB200_d_hss_read": (_i, [_pvp, C.c_char_p]),
    "SB200_d_hss_write": (_i, [_vp, C.c_char_p]),
    "SB200_d_hss_from_generators": (_i, [_pvp, _i, _vp, _vp, C.c_int64, _vp,
                                        C.c_int64]),
    "SB200_d_struct_mult_device": (_i, [_vp, C.c_char, _i, _vp, _i, _vp, _i, _vp]),
    "SB200_d_struct_factor_device": (_i, [_vp, _vp]),
    "SB200_d_struct_solve_device": (_i, [_vp, _i, _vp, _i, _vp]),
    "SB200_d_hss_apply": (_i, [_vp, C.c_char, _i, _vp, _i, _d, _vp, _i]),
    "SB200_d_hss_apply_device": (_i, [_vp, C.c_char, _i, _vp, _i, _d, _vp, _i, _vp]),
    "SB200_d_hss_extract": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i]),
    "SB200_d_hss_forward_solve": (_i, [_vp, _i, _vp, _i]),
    "SB200_d_hss_backward_solve": (_i, [_vp, _i, _vp, _i]),
    "SB200_d_hss_forward_solve_device": (_i, [_vp, _i, _vp, _i, _vp]),
    "SB200_d_hss_backward_solve_device": (_i, [_vp, _i, _vp, _i, _vp]),
    "SB200_d_hss_partial_factor": (_i, [_vp]),
    "SB200_d_hss_schur_sizes": (_i, [_vp, _vp]),
    "SB200_d_hss_schur_update": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i]),
    "SB200_d_hss_vhat": (_i, [_vp, _vp, _i]),
    "SB200_d_hss_schur_product_direct": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _i,
                                             _vp, _i, _vp, _i]),
    "SB200_d_hss_schur_product_indirect": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _i,
                                               _vp, _i, _vp, _i, _vp, _i]),
    "SB200_d_hss_partial_factor_device": (_i, [_vp, _vp]),
    "SB200_d_hss_schur_update_device": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp]),
    "SB200_d_hss_schur_product_direct_device": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _i,
                                                    _vp, _i, _vp, _i, _vp]),
    "SB200_d_hss_partial_forward_solve": (_i, [_vp, _i, _vp, _i, _vp, _i]),
    "SB200_d_hss_partial_x": (_i, [_vp, _i, _vp, _i, _i]),
    "SB200_d_hss_partial_backward_solve": (_i, [_vp, _i, _vp, _i]),
    "SB200_d_hss_set_partition": (_i, [_vp, _i, _i]),
    "SB200_d_hss_owned_range": (_i, [_vp, _vp, _vp]),
    "SB200_d_hss_dist_sizes": (_i, [_vp, _i, _vp]),
    "SB200_d_hss_dist_mult_begin": (_i, [_vp, C.c_char, _i, _vp, _i, _vp, _vp]),
    "SB200_d_hss_dist_mult_end": (_i, [_vp, C.c_char, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "SB200_d_hss_dist_factor_begin": (_i, [_vp, _vp, _vp]),
    "SB200_d_hss_dist_factor_end": (_i, [_vp, _vp, _vp]),
    "SB200_d_hss_dist_solve_begin": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "SB200_d_hss_dist_solve_end": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "SB200_d_hss_file_info": (_i, [C.c_char_p, _vp]),
    "SB200_d_hss_file_copy": (_i, [C.c_char_p, C.c_char_p]),
    "SB200_d_struct_levels": (_i, [_vp]),
    "SB200_d_struct_factor_nonzeros": (_ll, [_vp]),
    "SB200_d_struct_ulv_data": (_i, [_vp, _vp, _vp, _vp]),
    "SB200_d_struct_flops": (_ll, [_vp, _i]),
    "SB200_d_struct_set_profile": (_i, [_vp, _i]),
    "SB200_d_struct_kernel_ms": (_d, [_vp, _i]),
    "SB200_d_struct_launches": (_ll, [_vp]),
    "SB200_d_struct_print_info": (_i, [_vp]),
    "SB200_d_struct_dense": (_i, [_vp, _vp, _i]),
}


def lib():
    """dlopen the engine; raise (never fall back) if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"{_SO} not built: run `python __graft_entry__.py` "
                "(strumpack_b200 has no CPU fallback)")
        L = C.CDLL(_SO)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)      # AttributeError if a symbol is missing
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def default_options(type=SP_TYPE_HSS, **kw):
    o = CSPOptions()
    lib().SP_d_struct_default_options(C.byref(o))
    o.type = type
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"strumpack_b200: {what} failed (see stderr)")


def _fortran(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


def pack_generators(nodes):
    """Pack a pre-order list of node records (attributes: parent, ch (list of
    child indices), rows, cols, U_rows, U_rank, V_rows, V_rank, Pu/Pv (LAPACK
    ipiv, 1-based), Eu, Ev, D, B01, B10) into the flat arrays of
    ``SB200_d_hss_from_generators``."""
    tab = np.full((len(nodes), NODE_FIELDS), -1, dtype=np.int64)
    vals, perms = [], []
    nv = npm = 0

    def put(a):
        nonlocal nv
        a = np.asarray(a, dtype=np.float64)
        if a.size == 0:
            return -1
        off = nv
        vals.append(np.asfortranarray(a).ravel(order="F"))
        nv += a.size
        return off

    def putp(ipiv):
        nonlocal npm
        ipiv = np.asarray(ipiv)
        if ipiv.size == 0:
            return -1
        g = np.arange(ipiv.size, dtype=np.int32)
        for i, p in enumerate(ipiv):
            p = int(p) - 1
            if p != i:
                g[i], g[p] = g[p], g[i]
        off = npm
        perms.append(g)
        npm += g.size
        return off

    for i, n in enumerate(nodes):
        t = tab[i]
        t[0] = n.parent
        t[1], t[2] = (n.ch[0], n.ch[1]) if n.ch else (-1, -1)
        t[3], t[4] = n.rows, n.cols
        has_u, has_v = len(n.Pu) > 0, len(n.Pv) > 0
        t[5], t[6] = (len(n.Pu), n.Eu.shape[1]) if has_u else (0, 0)
        t[7], t[8] = (len(n.Pv), n.Ev.shape[1]) if has_v else (0, 0)
        t[9] = put(n.D)
        t[10], t[11] = put(n.Eu), put(n.Ev)
        t[12], t[13] = put(n.B01), put(n.B10)
        t[14], t[15] = putp(n.Pu), putp(n.Pv)
    v = np.concatenate(vals) if vals else np.zeros(0)
    p = np.concatenate(perms) if perms else np.zeros(0, dtype=np.int32)
    return tab, v, p


def fp64_dmma_peak_tflops():
    """fp64 tensor-pipe peak of the current GPU, measured now (TFLOP/s)."""
    return float(lib().SB200_fp64_dmma_peak_tflops())


def debug_qr_batch(A, k, count=1, variant=1, reps=1):
    """Leaf-QR kernel hook (tests / microbenchmark): returns (out, T, ms)."""
    A = np.asfortranarray(np.asarray(A, dtype=np.float64))
    m, naug = A.shape
    out = np.zeros_like(A, order="F")
    T = np.zeros((16, max(k, 1)), order="F")
    ms = C.c_float(0.0)
    _check(lib().SB200_debug_qr_batch(m, k, naug, count, A.ctypes.data, out.ctypes.data, T.ctypes.data,
                                      variant, reps, C.addressof(ms)), "debug_qr_batch")
    return out, T, float(ms.value)


def hss_file_info(path):
    """Host-only: statistics and flop counts of a reference HSS dump."""
    out = np.zeros(10, dtype=np.int64)
    _check(lib().SB200_d_hss_file_info(str(path).encode(), out.ctypes.data),
           "hss_file_info")
    keys = ("rows", "cols", "nodes", "levels", "rank", "nonzeros",
            "apply_flops", "factor_flops", "solve_flops", "factor_flops_exec")
    return dict(zip(keys, (int(v) for v in out)))


def hss_file_copy(src, dst):
    _check(lib().SB200_d_hss_file_copy(str(src).encode(), str(dst).encode()),
           "hss_file_copy")


def _block_callback(fn):
    """ctypes trampoline for SB200ElemBlockFn around ``fn(I, J) -> A[I, J]``."""
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)

    def tramp(nI, I, nJ, J, B, ldB, user):
        Ia = np.ctypeslib.as_array(I, shape=(nI,))
        Ja = np.ctypeslib.as_array(J, shape=(nJ,))
        out = np.ctypeslib.as_array(B, shape=(nJ, ldB))      # column-major nI x nJ with ld ldB
        out[:, :nI] = np.asarray(fn(Ia, Ja), dtype=np.float64).T
    return C.CFUNCTYPE(None, C.c_int, ip, C.c_int, ip, dp, C.c_int, C.c_void_p)(tramp)


class StructuredMatrix:
    """Mirror of ``structured::StructuredMatrix<double>``
    (reference src/structured/StructuredMatrix.hpp:209-418)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    # -- factories (reference StructuredMatrix.cpp:53-127, 193-312) ---------
    @classmethod
    def from_dense(cls, A, opts=None):
        A = _fortran(A)
        opts = opts or default_options()
        h = C.c_void_p()
        _check(lib().SP_d_struct_from_dense(
            C.byref(h), A.shape[0], A.shape[1], A.ctypes.data, A.shape[0],
            C.byref(opts)), "construct_from_dense")
        return cls(h.value)

    @classmethod
    def from_elements(cls, rows, cols, fn, opts=None):
        opts = opts or default_options()
        cb = C.CFUNCTYPE(C.c_double, C.c_int, C.c_int)(fn)
        h = C.c_void_p()
        _check(lib().SP_d_struct_from_elements(
            C.byref(h), rows, cols, C.cast(cb, C.c_void_p), C.byref(opts)),
            "construct_from_elements")
        return cls(h.value)

    # -- queries --------------------------------------------------------------
    @property
    def rows(self):
        return lib().SP_d_struct_rows(self._h)

    @property
    def cols(self):
        return lib().SP_d_struct_cols(self._h)

    @property
    def rank(self):
        return lib().SP_d_struct_rank(self._h)

    @property
    def memory(self):
        return lib().SP_d_struct_memory(self._h)

    @property
    def nonzeros(self):
        return lib().SP_d_struct_nonzeros(self._h)

    @property
    def levels(self):
        return lib().SB200_d_struct_levels(self._h)

    @property
    def factor_nonzeros(self):
        return lib().SB200_d_struct_factor_nonzeros(self._h)

    def ulv_data(self):
        """(factor arena, T arena) of the ULV factorization as numpy arrays
        (reference accessor HSSMatrix::ULV(), HSSMatrix.hpp:497)."""
        sz = np.zeros(2, dtype=np.int64)
        _check(lib().SB200_d_struct_ulv_data(self._h, None, None, sz.ctypes.data), "ulv_data")
        f = np.empty(int(sz[0])); t = np.empty(int(sz[1]))
        _check(lib().SB200_d_struct_ulv_data(self._h, f.ctypes.data, t.ctypes.data, sz.ctypes.data), "ulv_data")
        return f, t

    @property
    def launches(self):
        return lib().SB200_d_struct_launches(self._h)

    def flops(self, which):
        """'apply' | 'factor' | 'solve' (reference accounting) | 'factor_exec'"""
        k = {"apply": 0, "factor": 1, "solve": 2, "factor_exec": 3,
             "qr_leaf": 4, "qr_leaf_exec": 5}[which]
        return lib().SB200_d_struct_flops(self._h, k)

    def set_profile(self, on=True):
        _check(lib().SB200_d_struct_set_profile(self._h, int(on)), "set_profile")

    def kernel_ms(self, which=0):
        return lib().SB200_d_struct_kernel_ms(self._h, which)

    def print_info(self):
        _check(lib().SB200_d_struct_print_info(self._h), "print_info")

    def dense(self):
        A = np.zeros((self.rows, self.cols), order="F")
        _check(lib().SB200_d_struct_dense(self._h, A.ctypes.data, A.shape[0]),
               "dense")
        return A

    # -- the hot path, host operands (the reference-facing call) --------------
    def mult(self, x, trans="N", out=None):
        """y = op(S) x  (StructuredMatrix::mult, StructuredMatrix.hpp:280-300).
        ``out``: caller-allocated result (column-major ny x s, e.g. a view of
        pinned memory), as in the C interface where C is the caller's buffer."""
        x = _fortran(x)
        t = trans.upper() != "N"
        ny = self.cols if t else self.rows
        if out is None:
            y = np.zeros((ny, x.shape[1]), order="F")
        else:
            y = out
            if (y.dtype != np.float64 or y.shape != (ny, x.shape[1])
                    or not y.flags.f_contiguous):
                raise ValueError("mult: out must be a column-major float64 array of the result shape")
        _check(lib().SP_d_struct_mult(self._h, trans.encode()[:1], x.shape[1],
                                      x.ctypes.data, x.shape[0],
                                      y.ctypes.data, ny), "mult")
        return y

    def factor(self):
        _check(lib().SP_d_struct_factor(self._h), "factor")

    def solve(self, b, overwrite=False):
        """x = S^{-1} b (StructuredMatrix::solve, StructuredMatrix.hpp:340-360;
        the C call overwrites its argument, this wrapper returns a copy unless
        ``overwrite`` is set)."""
        if overwrite:      # the C semantics: b <- S^{-1} b in the caller's buffer
            x = b
            if x.dtype != np.float64 or x.ndim != 2 or not x.flags.f_contiguous:
                raise ValueError("solve(overwrite=True): b must be a column-major float64 matrix")
        else:
            x = _fortran(b).copy(order="F")
        _check(lib().SP_d_struct_solve(self._h, x.shape[1], x.ctypes.data,
                                       x.shape[0]), "solve")
        return x

    def shift(self, sigma):
        _check(lib().SP_d_struct_shift(self._h, float(sigma)), "shift")

    # -- device operands (torch CUDA tensors, column-major = .t() of a
    #    contiguous (s, n) tensor); queued on torch's current stream ----------
    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def mult_device(self, xT, yT, trans="N"):
        """xT, yT: torch float64 CUDA tensors of shape (s, n) contiguous, i.e.
        column-major n x s with ld = n."""
        s, n = xT.shape
        _check(lib().SB200_d_struct_mult_device(
            self._h, trans.encode()[:1], s, C.c_void_p(xT.data_ptr()), n,
            C.c_void_p(yT.data_ptr()), yT.shape[1], self._stream()),
            "mult_device")

    def factor_device(self):
        _check(lib().SB200_d_struct_factor_device(self._h, self._stream()),
               "factor_device")

    def solve_device(self, bT):
        s, n = bT.shape
        _check(lib().SB200_d_struct_solve_device(
            self._h, s, C.c_void_p(bT.data_ptr()), n, self._stream()),
            "solve_device")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().SP_d_struct_destroy(C.byref(self._h))
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BLRMatrix(StructuredMatrix):
    """Mirror of ``BLR::BLRMatrix<double>`` (reference src/BLR/BLRMatrix.hpp:68-291)."""

    @classmethod
    def compress_and_factor(cls, A, opts=None, pivot_threshold=-1.0, factor_algorithm=BLR_RL,
                            admissible=None, tiles=None):
        """BLRMatrix::compress_and_factor(A, admissible, opts) with tiles from
        ClusterTree(n).refine(leaf) (reference BLRMatrix.cpp:113-241,
        test/test_BLR_seq.cpp:136-156); factor_algorithm: BLR_RL / BLR_LL / ...;
        admissible: nb x nb matrix (None: weak admissibility); tiles: the caller's
        tile sizes (the `tiles` vector of BLRMatrix.hpp:91-101) instead of refine(leaf)."""
        A = _fortran(A)
        opts = opts or default_options(type=SP_TYPE_BLR, leaf_size=256)
        p, keep = _blr_params(pivot_threshold, factor_algorithm, admissible, tiles)
        h = C.c_void_p()
        _check(lib().SB200_d_blr_compress_and_factor_ex(
            C.byref(h), A.shape[0], A.ctypes.data, A.shape[0], C.byref(opts),
            C.addressof(p)), "compress_and_factor")
        return cls(h.value)

    @classmethod
    def compress_and_factor_device(cls, dA, opts=None, pivot_threshold=-1.0):
        """dA: torch float64 CUDA tensor holding the column-major n x n matrix
        (i.e. the transpose of a contiguous (n, n) row-major tensor)."""
        n = dA.shape[0]
        opts = opts or default_options(type=SP_TYPE_BLR, leaf_size=256)
        h = C.c_void_p()
        _check(lib().SB200_d_blr_compress_and_factor_device(
            C.byref(h), n, C.c_void_p(dA.data_ptr()), n, C.byref(opts),
            float(pivot_threshold)), "compress_and_factor_device")
        return cls(h.value)

    @classmethod
    def construct_and_partial_factor(cls, A11, A12, A21, A22, opts=None, pivot_threshold=-1.0,
                                     factor_algorithm=BLR_RL, admissible=None, tiles1=None, tiles2=None):
        """BLRMatrix::construct_and_partial_factor (reference BLRMatrix.cpp:739-1037,
        RL, weak admissibility, tiles from ClusterTree(n1/n2).refine(leaf)).
        Returns (F, S): F holds F11 = LU(A11), F12, F21 in BLR form, S is the dense
        Schur complement A22 - A21 A11^{-1} A12 (the reference updates A22 in place)."""
        A11, A12, A21 = _fortran(A11), _fortran(A12), _fortran(A21)
        S = _fortran(A22).copy(order="F")
        n1, n2 = A11.shape[0], S.shape[0]
        opts = opts or default_options(type=SP_TYPE_BLR, leaf_size=256)
        h = C.c_void_p()
        p, keep = _blr_params(pivot_threshold, factor_algorithm, admissible, tiles1, tiles2)
        _check(lib().SB200_d_blr_partial_factor_ex(
            C.byref(h), n1, n2, A11.ctypes.data, n1, A12.ctypes.data, max(n1, 1),
            A21.ctypes.data, max(n2, 1), S.ctypes.data, max(n2, 1), C.byref(opts),
            C.addressof(p)), "construct_and_partial_factor")
        return cls(h.value), S

    @classmethod
    def from_element_blocks(cls, n, fn, opts=None, factor=True, pivot_threshold=-1.0,
                            factor_algorithm=BLR_RL, admissible=None):
        """BLRMatrix::compress / compress_and_factor(const extract_t& Aelem, admissible, opts)
        (reference BLRMatrix.hpp:104-112): ``fn(I, J)`` returns A[I, J]."""
        opts = opts or default_options(type=SP_TYPE_BLR, leaf_size=256)
        p, keep = _blr_params(pivot_threshold, factor_algorithm, admissible)
        cb = _block_callback(fn)
        h = C.c_void_p()
        _check(lib().SB200_d_blr_from_element_blocks(C.byref(h), int(n), C.cast(cb, C.c_void_p), None,
                                                     C.byref(opts), C.addressof(p), int(bool(factor))),
               "from_element_blocks")
        return cls(h.value)

    @classmethod
    def construct_and_partial_factor_elements(cls, n1, n2, fn, opts=None, pivot_threshold=-1.0,
                                              factor_algorithm=BLR_RL):
        """construct_and_partial_factor(n1, n2, extractors, ...) (reference BLRMatrix.hpp:223-232):
        ``fn(I, J)`` defines the whole front.  Returns (F, Schur complement)."""
        opts = opts or default_options(type=SP_TYPE_BLR, leaf_size=256)
        p, keep = _blr_params(pivot_threshold, factor_algorithm, None)
        cb = _block_callback(fn)
        S = np.zeros((n2, n2), order="F")
        h = C.c_void_p()
        _check(lib().SB200_d_blr_partial_factor_element_blocks(
            C.byref(h), int(n1), int(n2), C.cast(cb, C.c_void_p), None, S.ctypes.data, max(n2, 1),
            C.byref(opts), C.addressof(p)), "construct_and_partial_factor_elements")
        return cls(h.value), S

    @property
    def sep_rows(self):
        return lib().SB200_d_blr_sep_rows(self._h)

    def partial_forward_solve(self, b):
        """[b_sep; b_upd] -> [L11^{-1} P b_sep; b_upd - F21 b_sep] (laswp + trsmLNU_gemm,
        reference BLRMatrix.cpp:1552-1608, FrontBLR.cpp:529-531)."""
        x = _fortran(b).copy(order="F")
        _check(lib().SB200_d_blr_partial_forward_solve(self._h, x.shape[1], x.ctypes.data, x.shape[0]),
               "partial_forward_solve")
        return x

    def partial_backward_solve(self, y):
        """[y_sep; y_upd] -> [U11^{-1} (y_sep - F12 y_upd); y_upd] (gemm_trsmUNN,
        reference BLRMatrix.cpp:1610-1665, FrontBLR.cpp:555)."""
        x = _fortran(y).copy(order="F")
        _check(lib().SB200_d_blr_partial_backward_solve(self._h, x.shape[1], x.ctypes.data, x.shape[0]),
               "partial_backward_solve")
        return x

    @property
    def tiles(self):
        return lib().SB200_d_blr_tiles(self._h)

    @property
    def dense_tiles(self):
        """off-diagonal tiles kept dense (incompressible at the tolerance)"""
        return lib().SB200_d_blr_dense_tiles(self._h)


class HSSMatrix(StructuredMatrix):
    """Mirror of ``HSS::HSSMatrix<double>`` (reference src/HSS/HSSMatrix.hpp)."""

    @classmethod
    def read(cls, path):
        """HSSMatrix::read (reference HSSMatrix.cpp:488-510)."""
        h = C.c_void_p()
        _check(lib().SB200_d_hss_read(C.byref(h), str(path).encode()), "read")
        return cls(h.value)

    def write(self, path):
        _check(lib().SB200_d_hss_write(self._h, str(path).encode()), "write")

    def apply(self, b, beta=0.0, c=None, trans="N"):
        """apply_HSS(op, H, B, beta, C): returns op(H) b + beta c
        (reference HSSMatrix.cpp:419-435)."""
        b = _fortran(b)
        t = trans.upper() != "N"
        ny = self.cols if t else self.rows
        out = np.zeros((ny, b.shape[1]), order="F") if c is None else _fortran(c).copy(order="F")
        _check(lib().SB200_d_hss_apply(self._h, trans.encode()[:1], b.shape[1], b.ctypes.data,
                                       b.shape[0], float(beta), out.ctypes.data, ny), "apply")
        return out

    def extract(self, I, J, add_to=None):
        """HSSMatrix::extract(I, J) (or extract_add when ``add_to`` is given):
        the dense sub-block H(I, J) (reference HSSMatrix.extract.hpp:8-188)."""
        I = np.ascontiguousarray(I, dtype=np.int32)
        J = np.ascontiguousarray(J, dtype=np.int32)
        B = (np.zeros((I.size, J.size), order="F") if add_to is None
             else _fortran(add_to).copy(order="F"))
        if I.size and J.size:
            _check(lib().SB200_d_hss_extract(self._h, I.size, I.ctypes.data, J.size, J.ctypes.data,
                                             B.ctypes.data, max(I.size, 1), int(add_to is not None)),
                   "extract")
        return B

    def get(self, i, j):
        """HSSMatrix::get(i, j) (reference HSSMatrix.extract.hpp:8-16)."""
        return float(self.extract([i], [j])[0, 0])

    def forward_solve(self, b):
        """HSSMatrix::forward_solve (HSSMatrix.solve.hpp:52-58); state stays in the object."""
        b = _fortran(b)
        _check(lib().SB200_d_hss_forward_solve(self._h, b.shape[1], b.ctypes.data, b.shape[0]),
               "forward_solve")
        self._fwd_shape = b.shape

    def backward_solve(self):
        """HSSMatrix::backward_solve (HSSMatrix.solve.hpp:60-66): returns x."""
        n, s = self._fwd_shape
        x = np.zeros((n, s), order="F")
        _check(lib().SB200_d_hss_backward_solve(self._h, s, x.ctypes.data, n), "backward_solve")
        return x

    # -- Schur complement of the (0,0) block (what FrontHSS uses) ------------
    def partial_factor(self):
        """HSSMatrix::partial_factor (reference HSSMatrix.factor.hpp:44-50)."""
        _check(lib().SB200_d_hss_partial_factor(self._h), "partial_factor")

    def schur_sizes(self):
        out = np.zeros(7, dtype=np.int32)
        _check(lib().SB200_d_hss_schur_sizes(self._h, out.ctypes.data), "schur_sizes")
        keys = ("rows1", "cols1", "rv0", "m0", "rv1", "ru1", "rows0")
        return dict(zip(keys, (int(v) for v in out)))

    def schur_update(self):
        """HSSMatrix::Schur_update (reference HSSMatrix.Schur.hpp:40-59):
        returns (Theta, DUB01, Phi); S = H11 - Theta Vhat^H Phi^H."""
        z = self.schur_sizes()
        Theta = np.zeros((z["rows1"], z["rv0"]), order="F")
        DUB01 = np.zeros((z["m0"], z["rv1"]), order="F")
        Phi = np.zeros((z["cols1"], z["m0"]), order="F")
        _check(lib().SB200_d_hss_schur_update(
            self._h, Theta.ctypes.data, max(z["rows1"], 1), DUB01.ctypes.data, max(z["m0"], 1),
            Phi.ctypes.data, max(z["cols1"], 1)), "Schur_update")
        return Theta, DUB01, Phi

    def schur_device(self, RT):
        """Device-resident path: partial_factor + Schur_update + Schur_product_direct with
        torch CUDA tensors on the current stream.  RT: (c, rows1) contiguous float64
        tensor (= column-major rows1 x c).  Returns (SrT, ScT) of the same shape."""
        import torch
        z = self.schur_sizes()
        dev = RT.device
        c = RT.shape[0]
        mk = lambda r, q: torch.empty((max(q, 1), max(r, 1)), dtype=torch.float64, device=dev)
        Th, Du, Ph = mk(z["rows1"], z["rv0"]), mk(z["m0"], z["rv1"]), mk(z["cols1"], z["m0"])
        SrT, ScT = torch.empty_like(RT), torch.empty_like(RT)
        st = self._stream()
        p = lambda t: C.c_void_p(t.data_ptr())
        _check(lib().SB200_d_hss_partial_factor_device(self._h, st), "partial_factor_device")
        _check(lib().SB200_d_hss_schur_update_device(
            self._h, p(Th), max(z["rows1"], 1), p(Du), max(z["m0"], 1), p(Ph), max(z["cols1"], 1), st),
            "schur_update_device")
        _check(lib().SB200_d_hss_schur_product_direct_device(
            self._h, p(Th), max(z["rows1"], 1), p(Du), max(z["m0"], 1), p(Ph), max(z["cols1"], 1), c,
            p(RT), z["rows1"], p(SrT), z["rows1"], p(ScT), z["cols1"], st), "schur_product_direct_device")
        return SrT, ScT

    def vhat(self):
        """child(0)->ULV().Vhat() (reference HSSExtra.hpp:197-212)."""
        z = self.schur_sizes()
        V = np.zeros((z["m0"], z["rv0"]), order="F")
        _check(lib().SB200_d_hss_vhat(self._h, V.ctypes.data, max(z["m0"], 1)), "Vhat")
        return V

    def schur_product_direct(self, R, Theta=None, DUB01=None, Phi=None):
        """HSSMatrix::Schur_product_direct (reference HSSMatrix.Schur.hpp:73-137):
        returns (Sr, Sc) = (S R, S^H R).  Theta/DUB01/Phi default to the ones of
        the last schur_update (kept on the device)."""
        R = _fortran(R)
        z = self.schur_sizes()
        Sr = np.zeros((z["rows1"], R.shape[1]), order="F")
        Sc = np.zeros((z["cols1"], R.shape[1]), order="F")

        def arg(a):
            if a is None:
                return None, 1
            a = _fortran(a)
            keep.append(a)
            return a.ctypes.data, max(a.shape[0], 1)
        keep = []
        (t, ldt), (d, ldd), (p_, ldp) = arg(Theta), arg(DUB01), arg(Phi)
        _check(lib().SB200_d_hss_schur_product_direct(
            self._h, t, ldt, d, ldd, p_, ldp, R.shape[1], R.ctypes.data, R.shape[0],
            Sr.ctypes.data, max(z["rows1"], 1), Sc.ctypes.data, max(z["cols1"], 1)),
            "Schur_product_direct")
        return Sr, Sc

    def schur_product_indirect(self, R0, R1, Sr1, Sc1, DUB01=None):
        """HSSMatrix::Schur_product_indirect (reference HSSMatrix.Schur.hpp:139-215)."""
        R0, R1, Sr1, Sc1 = (_fortran(a) for a in (R0, R1, Sr1, Sc1))
        z = self.schur_sizes()
        c = R1.shape[1]
        Sr = np.zeros((z["rows1"], c), order="F")
        Sc = np.zeros((z["cols1"], c), order="F")
        d = None if DUB01 is None else _fortran(DUB01)
        _check(lib().SB200_d_hss_schur_product_indirect(
            self._h, None if d is None else d.ctypes.data, max(z["m0"], 1), c,
            R0.ctypes.data, R0.shape[0], R1.ctypes.data, R1.shape[0],
            Sr1.ctypes.data, Sr1.shape[0], Sc1.ctypes.data, Sc1.shape[0],
            Sr.ctypes.data, max(z["rows1"], 1), Sc.ctypes.data, max(z["cols1"], 1)),
            "Schur_product_indirect")
        return Sr, Sc

    def partial_forward_solve(self, b0):
        """child(0)->forward_solve(w, b0, partial=True) (reference
        HSSMatrix.solve.hpp:133-152): returns reduced_rhs (rv0 x s)."""
        b0 = _fortran(b0)
        z = self.schur_sizes()
        red = np.zeros((z["rv0"], b0.shape[1]), order="F")
        _check(lib().SB200_d_hss_partial_forward_solve(
            self._h, b0.shape[1], b0.ctypes.data, b0.shape[0], red.ctypes.data, max(z["rv0"], 1)),
            "partial forward_solve")
        self._pfwd_s = b0.shape[1]
        return red

    def partial_x(self, new=None):
        """The reduced solution x = D0^{-1} f (m0 x s) between the partial
        forward and backward solves; ``new`` overwrites it."""
        z = self.schur_sizes()
        s = self._pfwd_s
        if new is not None:
            x = _fortran(new)
            _check(lib().SB200_d_hss_partial_x(self._h, s, x.ctypes.data, max(z["m0"], 1), 1), "partial_x")
            return x
        x = np.zeros((z["m0"], s), order="F")
        _check(lib().SB200_d_hss_partial_x(self._h, s, x.ctypes.data, max(z["m0"], 1), 0), "partial_x")
        return x

    def partial_backward_solve(self):
        """child(0)->backward_solve(w, x0) (reference HSSMatrix.solve.hpp:62-66)."""
        z = self.schur_sizes()
        x0 = np.zeros((z["rows0"], self._pfwd_s), order="F")
        _check(lib().SB200_d_hss_partial_backward_solve(
            self._h, self._pfwd_s, x0.ctypes.data, max(z["rows0"], 1)), "partial backward_solve")
        return x0

    @classmethod
    def from_generators(cls, nodes):
        tab, vals, perms = pack_generators(nodes)
        h = C.c_void_p()
        _check(lib().SB200_d_hss_from_generators(
            C.byref(h), len(nodes), tab.ctypes.data, vals.ctypes.data,
            vals.size, perms.ctypes.data, perms.size), "from_generators")
        return cls(h.value)

    @classmethod
    def from_element_blocks(cls, n, fn, opts=None):
        """HSSMatrix::compress(Amult, Aelem, opts) with the reference's block
        extraction callback (HSSMatrix.hpp:68-70): ``fn(I, J)`` returns the dense
        sub-block A[I, J] for integer index arrays."""
        opts = opts or default_options()
        cb = _block_callback(fn)
        h = C.c_void_p()
        _check(lib().SB200_d_hss_from_element_blocks(C.byref(h), int(n), C.cast(cb, C.c_void_p), None,
                                                     C.byref(opts)), "from_element_blocks")
        return cls(h.value)

    @classmethod
    def from_kernel(cls, pts, kernel=KERNEL_GAUSS, h=1.0, lam=0.0, opts=None, clustering=CLUSTER_KD_TREE):
        """HSSMatrix(kernel::Kernel&, opts) (reference HSSMatrix.cpp:88-106).
        pts: d x n.  clustering: CLUSTER_NATURAL / CLUSTER_TWO_MEANS (the reference's
        default) / CLUSTER_KD_TREE.  Returns (H, perm, pts_permuted)."""
        pts = np.asfortranarray(np.array(pts, dtype=np.float64))
        d, n = pts.shape
        opts = opts or default_options()
        perm = np.zeros(n, dtype=np.int32)
        hd = C.c_void_p()
        _check(lib().SB200_d_hss_from_kernel_ex(
            C.byref(hd), n, d, pts.ctypes.data, kernel, h, lam, C.byref(opts),
            perm.ctypes.data, int(clustering)), "from_kernel")
        return cls(hd.value), perm, pts
