"""The single-precision C interface SP_s_struct_* (reference
src/structured/StructuredMatrix.h:103-569): float operands at the boundary,
the engine's fp64 kernels inside.  Tolerance: float rounding of the inputs and
outputs (eps_f = 6e-8) times a modest growth factor, plus the compression
tolerance -- the bound the reference's test_structure_reuse-style checks use
for float is 1e2*tol as for double."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("stype", ["HSS", "BLR"])
def test_float_interface(built, stype):
    sb = built
    L = sb.lib()
    n, tol = 700, 1e-4
    i = np.arange(n)
    A = (1.0 / (1.0 + np.abs(i[:, None] - i[None, :])) + 2.0 * np.eye(n)).astype(np.float32, order="F")
    o = sb.CSPOptions()
    L.SP_s_struct_default_options(C.byref(o))
    # StructuredOptions<float>: rel_tol 1e-2, abs_tol 1e-5 (reference StructuredOptions.hpp:49-54)
    assert o.type == sb.SP_TYPE_BLR and o.rel_tol == 1e-2 and o.abs_tol == 1e-5 and o.leaf_size == 128
    o.abs_tol = 1e-10
    o.type = sb.SP_TYPE_HSS if stype == "HSS" else sb.SP_TYPE_BLR
    o.rel_tol, o.leaf_size = tol, 64
    h = C.c_void_p()
    assert L.SP_s_struct_from_dense(C.byref(h), n, n, A.ctypes.data, n, C.byref(o)) == 0
    assert (L.SP_s_struct_rows(h), L.SP_s_struct_cols(h)) == (n, n)
    assert 0 < L.SP_s_struct_rank(h) < 64
    assert 0 < L.SP_s_struct_nonzeros(h) < n * n and L.SP_s_struct_memory(h) > 0
    x = np.asfortranarray(np.random.default_rng(0).standard_normal((n, 3)).astype(np.float32))
    y = np.zeros((n, 3), dtype=np.float32, order="F")
    assert L.SP_s_struct_mult(h, b"N", 3, x.ctypes.data, n, y.ctypes.data, n) == 0
    Ad, xd = A.astype(np.float64), x.astype(np.float64)
    assert rel(y, Ad @ xd) <= 1e2 * tol
    assert y.dtype == np.float32
    if stype == "HSS":
        yt = np.zeros_like(y)
        assert L.SP_s_struct_mult(h, b"T", 3, x.ctypes.data, n, yt.ctypes.data, n) == 0
        assert rel(yt, Ad.T @ xd) <= 1e2 * tol
        assert L.SP_s_struct_shift(h, C.c_float(0.5)) == 0
        assert L.SP_s_struct_factor(h) == 0
        b = np.asfortranarray(((Ad + 0.5 * np.eye(n)) @ xd).astype(np.float32))
        assert L.SP_s_struct_solve(h, 3, b.ctypes.data, n) == 0
        assert rel(b, xd) <= 1e2 * tol
    else:
        assert L.SP_s_struct_factor(h) == 1          # as in the reference: no factor() on a compressed BLR matrix
    L.SP_s_struct_destroy(C.byref(h))
    assert h.value is None


def test_float_from_elements(built):
    sb = built
    L = sb.lib()
    n = 256
    cb = C.CFUNCTYPE(C.c_float, C.c_int, C.c_int)(lambda r, c: 2.0 if r == c else 1.0 / (1 + abs(r - c)))
    o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=1e-6, leaf_size=32)
    h = C.c_void_p()
    assert L.SP_s_struct_from_elements(C.byref(h), n, n, C.cast(cb, C.c_void_p), C.byref(o)) == 0
    i = np.arange(n)
    A = 1.0 / (1.0 + np.abs(i[:, None] - i[None, :])) + np.eye(n)
    x = np.asfortranarray(np.random.default_rng(1).standard_normal((n, 1)).astype(np.float32))
    y = np.zeros_like(x)
    assert L.SP_s_struct_mult(h, b"N", 1, x.ctypes.data, n, y.ctypes.data, n) == 0
    assert rel(y, A @ x.astype(np.float64)) <= 1e-5
    L.SP_s_struct_destroy(C.byref(h))
