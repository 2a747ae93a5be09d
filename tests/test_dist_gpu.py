"""GPU tests of the subtree-sharded engine (C ABI SB200_d_hss_dist_*).

On a 1-GPU box the `world` ranks are emulated in one process: one handle per
rank on the same device, the all-gather replaced by a concatenation -- this
exercises exactly the rank-local kernels, pack/unpack and the replicated top
sweeps.  With >= 2 GPUs visible the real NCCL path runs under mp.spawn."""
import os
import sys

import numpy as np
import pytest

from conftest import CASES, GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("world,case", [(2, CASES[0]), (4, CASES[2]), (2, CASES[1]), (8, CASES[2])])
def test_sharded_engine_single_gpu_emulation(built, world, case):
    import torch
    from strumpack_b200.dist import GpuShardEngine
    sb = built
    path = os.path.join(GOLDEN, case + ".hss")
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    Hs = [sb.HSSMatrix.read(path) for _ in range(world)]
    engs = [GpuShardEngine(H, world, r) for r, H in enumerate(Hs)]
    # owned ranges partition the rows
    ranges = sorted((e.lo, e.hi) for e in engs)
    assert ranges[0][0] == 0 and ranges[-1][1] == Hs[0].rows
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    s = g["x"].shape[1]
    xT = torch.tensor(g["x"].T.copy(), device="cuda")

    def exchange(kind, begin, end):
        n = engs[0].sizes(s if kind != 1 else 1)[kind]
        sends = [e.new_buffer(n) for e in engs]
        for e, b in zip(engs, sends):
            begin(e, b)
        recv = torch.cat(sends)
        for e in engs:
            end(e, recv)

    yTs = [torch.zeros_like(xT) for _ in engs]
    exchange(0, lambda e, b: e.mult_begin(xT, b), lambda e, r: e.mult_end(xT, yTs[engs.index(e)], r))
    y = torch.zeros_like(xT)
    for e, yT in zip(engs, yTs):
        y[:, e.lo:e.hi] = yT[:, e.lo:e.hi]
    torch.cuda.synchronize()
    assert rel(y.cpu().numpy().T, g["y"]) < 1e-13
    exchange(1, lambda e, b: e.factor_begin(b), lambda e, r: e.factor_end(r))
    bTs = [torch.tensor(g["y"].T.copy(), device="cuda") for _ in engs]
    exchange(2, lambda e, b: e.solve_begin(bTs[engs.index(e)], b),
             lambda e, r: e.solve_end(bTs[engs.index(e)], r))
    xs = torch.zeros_like(xT)
    for e, bT in zip(engs, bTs):
        xs[:, e.lo:e.hi] = bT[:, e.lo:e.hi]
    torch.cuda.synchronize()
    assert rel(xs.cpu().numpy().T, g["xs"]) < 1e-10


def test_unsharded_calls_refuse_sharded_matrix(built, capfd):
    sb = built
    from strumpack_b200.dist import GpuShardEngine
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, CASES[0] + ".hss"))
    GpuShardEngine(H, 2, 0)
    with pytest.raises(RuntimeError):
        H.mult(np.ones(H.rows))
    assert "sharded" in capfd.readouterr().err


def _nccl_worker(rank, world, port, case, out, in_engine=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import strumpack_b200 as sb
    from strumpack_b200.dist import GpuShardEngine, ShardedHSS, NcclShardedHSS
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    H = sb.HSSMatrix.read(os.path.join(GOLDEN, case + ".hss"))
    S = NcclShardedHSS(H, world, rank) if in_engine else ShardedHSS(GpuShardEngine(H, world, rank))
    stream = torch.cuda.Stream()          # not the legacy default stream: the engine replays CUDA graphs
    torch.cuda.set_stream(stream)
    xT = torch.tensor(g["x"].T.copy(), device="cuda")
    yT = torch.zeros_like(xT)
    S.mult(xT, yT)
    y = S.gather_rows(yT)
    S.factor()
    bT = torch.tensor(g["y"].T.copy(), device="cuda")
    S.solve(bT)
    xs = S.gather_rows(bT)
    for _ in range(3):                    # again: from the second call on the sequences are graph replays
        yT.zero_()
        S.mult(xT, yT)
        S.factor()
        bT.copy_(torch.tensor(g["y"].T.copy(), device="cuda"))
        S.solve(bT)
    y2, xs2 = S.gather_rows(yT), S.gather_rows(bT)
    if rank == 0:
        torch.save({"y": y, "xs": xs, "y2": y2, "xs2": xs2}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("in_engine", [False, True])
def test_sharded_engine_nccl(built, tmp_path, in_engine):
    """Real NCCL (needs >= 2 GPUs): the Python-driven begin / all_gather / end protocol and the in-engine
    exchange (SB200_d_hss_dist_*: ncclAllGather issued by the engine, CUDA-graph replays) against the
    reference's golden vectors."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world, case = 2, CASES[2]
    out = str(tmp_path / "res.pt")
    mp.spawn(_nccl_worker, args=(world, 29611 + int(in_engine), case, out, in_engine), nprocs=world, join=True)
    res = torch.load(out)
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    for ky, kx in (("y", "xs"), ("y2", "xs2")):
        assert rel(res[ky].numpy().T, g["y"]) < 1e-13
        assert rel(res[kx].numpy().T, g["xs"]) < 1e-10
