"""Subtree-sharded HSS over the GPUs of one node.

Rank g owns the subtree of the g-th node at depth log2(world); the world-1 nodes
above the cut are replicated.  Every operation (apply, ULV factor, ULV solve) is
`begin` (rank-local sweep, fills a small send buffer) -> ONE all-gather over
NCCL -> `end` (replicated top of the tree, then the rank-local sweep back down).
The exchanged payload is O(rank^2) doubles per GPU (t1 | [Vt1, Dt] | [z; ft1] of
the cut node): latency-, not bandwidth-bound.  Replaces the reference's
MPI/BLACS subtree mapping (src/HSS/HSSMatrixMPI.cpp:317-345).

`ShardedHSS` drives any engine exposing the begin/end protocol; the GPU engine
is `GpuShardEngine` (C ABI SB200_d_hss_dist_*), tests drive a CPU engine through
the same code with the gloo backend.
"""
import ctypes as C

import numpy as np


class GpuShardEngine:
    """begin/end protocol on top of the C ABI; buffers are torch CUDA tensors."""

    def __init__(self, H, world, rank):
        import torch
        from . import lib, _check
        self.H, self.world, self.rank = H, world, rank
        self._lib, self._check, self.torch = lib(), _check, torch
        _check(self._lib.SB200_d_hss_set_partition(H._h, world, rank), "set_partition")
        lo, hi = C.c_int(), C.c_int()
        _check(self._lib.SB200_d_hss_owned_range(H._h, C.byref(lo), C.byref(hi)), "owned_range")
        self.lo, self.hi = lo.value, hi.value

    def sizes(self, s):
        out = np.zeros(3, dtype=np.int64)
        self._check(self._lib.SB200_d_hss_dist_sizes(self.H._h, s, out.ctypes.data), "dist_sizes")
        return [int(v) for v in out]

    def new_buffer(self, n):
        return self.torch.zeros(n, dtype=self.torch.float64, device="cuda")

    def _st(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def mult_begin(self, xT, send, trans="N"):
        s, n = xT.shape
        self._check(self._lib.SB200_d_hss_dist_mult_begin(
            self.H._h, trans.encode()[:1], s, self._p(xT), n, self._p(send), self._st()), "mult_begin")

    def mult_end(self, xT, yT, recv, trans="N"):
        s, n = xT.shape
        self._check(self._lib.SB200_d_hss_dist_mult_end(
            self.H._h, trans.encode()[:1], s, self._p(xT), n, self._p(yT), yT.shape[1],
            self._p(recv), self._st()), "mult_end")

    def factor_begin(self, send):
        self._check(self._lib.SB200_d_hss_dist_factor_begin(self.H._h, self._p(send), self._st()),
                    "factor_begin")

    def factor_end(self, recv):
        self._check(self._lib.SB200_d_hss_dist_factor_end(self.H._h, self._p(recv), self._st()),
                    "factor_end")

    def solve_begin(self, bT, send):
        s, n = bT.shape
        self._check(self._lib.SB200_d_hss_dist_solve_begin(
            self.H._h, s, self._p(bT), n, self._p(send), self._st()), "solve_begin")

    def solve_end(self, bT, recv):
        s, n = bT.shape
        self._check(self._lib.SB200_d_hss_dist_solve_end(
            self.H._h, s, self._p(bT), n, self._p(recv), self._st()), "solve_end")


class ShardedHSS:
    """y = H x, H = ULV, x = H^{-1} b on a subtree-sharded HSS matrix.
    Vectors are full length on every rank; only rows [lo, hi) are read/written."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.e, self.dist, self.group = engine, dist, group
        self.world = dist.get_world_size(group)
        self._bufs = {}

    @property
    def owned(self):
        return self.e.lo, self.e.hi

    def _exchange(self, key, n, fill):
        if (key, n) not in self._bufs:
            self._bufs[(key, n)] = (self.e.new_buffer(n), self.e.new_buffer(n * self.world))
        send, recv = self._bufs[(key, n)]
        fill(send)
        self.dist.all_gather_into_tensor(recv, send, group=self.group)
        return recv

    def mult(self, xT, yT, trans="N"):
        n = self.e.sizes(xT.shape[0])[0]
        recv = self._exchange("a", n, lambda send: self.e.mult_begin(xT, send, trans))
        self.e.mult_end(xT, yT, recv, trans)

    def factor(self):
        n = self.e.sizes(1)[1]
        recv = self._exchange("f", n, lambda send: self.e.factor_begin(send))
        self.e.factor_end(recv)

    def solve(self, bT):
        n = self.e.sizes(bT.shape[0])[2]
        recv = self._exchange("s", n, lambda send: self.e.solve_begin(bT, send))
        self.e.solve_end(bT, recv)

    def gather_rows(self, vT):
        """Assemble the full vector from the owned slices (all ranks)."""
        import torch
        lo, hi = self.owned
        parts = [None] * self.world
        self.dist.all_gather_object(parts, (lo, hi, vT[:, lo:hi].cpu()), group=self.group)
        out = torch.zeros_like(vT, device="cpu")
        for l, h, t in parts:
            out[:, l:h] = t
        return out


class NcclShardedHSS:
    """The sharded matrix with the exchange inside the engine (C ABI SB200_d_hss_dist_{init,mult,factor,solve}):
    one C call per operation queues the local sweep, ONE ncclAllGather and the replicated top on the caller's
    stream; from the second call on the whole sequence is one CUDA graph.  `torch.distributed` is only used once,
    to hand rank 0's NCCL unique id to the other ranks."""

    def __init__(self, H, world, rank, group=None):
        import torch
        import torch.distributed as dist
        from . import lib, _check
        self.H, self.world, self.rank, self.torch = H, world, rank, torch
        self._lib, self._check = lib(), _check
        uid = C.create_string_buffer(128)
        if rank == 0:
            _check(self._lib.SB200_nccl_unique_id(uid), "nccl_unique_id")
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        _check(self._lib.SB200_d_hss_dist_init(H._h, world, rank, box[0]), "dist_init")
        lo, hi = C.c_int(), C.c_int()
        _check(self._lib.SB200_d_hss_owned_range(H._h, C.byref(lo), C.byref(hi)), "owned_range")
        self.lo, self.hi = lo.value, hi.value
        self._dist, self._group = dist, group

    @property
    def owned(self):
        return self.lo, self.hi

    def _st(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def mult(self, xT, yT, trans="N"):
        s, n = xT.shape
        self._check(self._lib.SB200_d_hss_dist_mult(self.H._h, trans.encode()[:1], s, C.c_void_p(xT.data_ptr()), n,
                                                    C.c_void_p(yT.data_ptr()), yT.shape[1], self._st()), "dist_mult")

    def factor(self):
        self._check(self._lib.SB200_d_hss_dist_factor(self.H._h, self._st()), "dist_factor")

    def solve(self, bT):
        s, n = bT.shape
        self._check(self._lib.SB200_d_hss_dist_solve(self.H._h, s, C.c_void_p(bT.data_ptr()), n, self._st()),
                    "dist_solve")

    def gather_rows(self, vT):
        import torch
        parts = [None] * self.world
        self._dist.all_gather_object(parts, (self.lo, self.hi, vT[:, self.lo:self.hi].cpu()), group=self._group)
        out = torch.zeros_like(vT, device="cpu")
        for l, h, t in parts:
            out[:, l:h] = t
        return out
