"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product.

Reader for the reference's HSS binary dump (``HSSMatrix<double>::write``,
reference src/HSS/HSSMatrix.cpp:438-486; ``DenseMatrix`` record
src/dense/DenseMatrix.cpp:881-889; ``HSSBasisID`` record
src/HSS/HSSBasisID.hpp:93-101), x86-64 / gcc ABI:

  file   := int32 version[3] , node
  node   := u64 rows, u64 cols, char U_state, char V_state, int32 omp_depth,
            u8 active, int32 U_rank, U_rows, V_rank, V_rows,
            dense Asub, basis U, basis V, dense D, dense B01, dense B10,
            int32 nchild, node*nchild                       (pre-order)
  dense  := int32 version[3], 40 raw bytes of the C++ object
            {vptr? no: data_*, rows_, cols_, ld_ + vptr = 5 x 8 bytes},
            rows*cols float64 column-major
  basis  := u64 Psize, int32 P[Psize] (LAPACK ipiv, 1-based), dense E

The result is a list of ``Node`` records in pre-order, i.e. the generators
(D, P_u, E_u, P_v, E_v, B01, B10) of every node -- what the GPU engine eats.
"""
import struct
from dataclasses import dataclass, field
import numpy as np


@dataclass
class Node:
    idx: int
    parent: int
    rows: int
    cols: int
    U_rank: int
    U_rows: int
    V_rank: int
    V_rows: int
    Pu: np.ndarray        # LAPACK ipiv (1-based sequential swaps), len U_rows
    Eu: np.ndarray        # (U_rows-U_rank) x U_rank
    Pv: np.ndarray
    Ev: np.ndarray
    D: np.ndarray         # leaf: rows x cols ; else 0x0
    B01: np.ndarray       # U_rank(c0) x V_rank(c1)
    B10: np.ndarray       # U_rank(c1) x V_rank(c0)
    ch: list = field(default_factory=list)
    row_off: int = 0
    col_off: int = 0

    @property
    def leaf(self):
        return not self.ch


class _Reader:
    def __init__(self, buf):
        self.b = memoryview(buf)
        self.p = 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def dense(self):
        self.take("3i")
        # DenseMatrix<double> object image: vptr, data_, rows_, cols_, ld_
        _vptr, _data, rows, cols, _ld = self.take("5Q")
        n = rows * cols
        a = np.frombuffer(self.b, dtype="<f8", count=n, offset=self.p)
        self.p += 8 * n
        return a.reshape((rows, cols), order="F").copy(order="F")

    def basis(self):
        psize = self.take("Q")
        P = np.frombuffer(self.b, dtype="<i4", count=psize, offset=self.p).copy()
        self.p += 4 * psize
        E = self.dense()
        return P, E


def ipiv_to_gather(ipiv):
    """LAPACK ipiv (1-based, sequential row swaps, ``laswp`` forward) ->
    gather index g with (P^T b)[i] = b[g[i]]   (HSSBasisID.hpp:45-46,
    DenseMatrix::laswp fwd)."""
    g = np.arange(len(ipiv), dtype=np.int64)
    for i, p in enumerate(ipiv):
        p = int(p) - 1
        if p != i:
            g[i], g[p] = g[p], g[i]
    return g


def read_hss(path):
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(buf)
    version = r.take("3i")
    nodes = []

    def rec(parent, row_off, col_off):
        rows, cols = r.take("QQ")
        r.take("cc")           # U_state, V_state
        r.take("i")            # openmp_task_depth_
        r.take("B")            # active_
        U_rank, U_rows, V_rank, V_rows = r.take("4i")
        r.dense()              # Asub_ (empty outside MPI redistribution)
        Pu, Eu = r.basis()
        Pv, Ev = r.basis()
        D = r.dense()
        B01 = r.dense()
        B10 = r.dense()
        nc = r.take("i")
        nd = Node(len(nodes), parent, rows, cols, U_rank, U_rows, V_rank,
                  V_rows, Pu, Eu, Pv, Ev, D, B01, B10, [], row_off, col_off)
        nodes.append(nd)
        ro, co = row_off, col_off
        for _ in range(nc):
            c = rec(nd.idx, ro, co)
            nd.ch.append(c)
            ro += nodes[c].rows
            co += nodes[c].cols
        return nd.idx

    rec(-1, 0, 0)
    assert r.p == len(buf), (r.p, len(buf))
    return nodes, version
