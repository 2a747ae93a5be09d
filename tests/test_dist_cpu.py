"""CPU tests of the multi-GPU protocol (world_size 2 and 4, gloo backend).

The driver under test is the product's `strumpack_b200.dist.ShardedHSS`
(begin -> all_gather -> end); the engine plugged into it here is the oracle's
CPU `ShardOracle`, which speaks the same payload layout as the C ABI
`SB200_d_hss_dist_*`.  Result must equal the unsharded oracle (= the reference)
on the same generators."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _worker(rank, world, port, case, out):
    sys.path.insert(0, ROOT)
    sys.setrecursionlimit(20000)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import hss_file, hss_oracle as ho
    from strumpack_b200.dist import ShardedHSS
    nodes, _ = hss_file.read_hss(os.path.join(GOLDEN, case + ".hss"))
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    eng = ho.ShardOracle(nodes, world, rank)
    S = ShardedHSS(eng)
    xT = torch.from_numpy(g["x"].T.copy())
    yT = torch.zeros_like(xT)
    S.mult(xT, yT)
    y = S.gather_rows(yT)
    S.factor()
    bT = torch.from_numpy(g["y"].T.copy())
    S.solve(bT)
    xs = S.gather_rows(bT)
    if rank == 0:
        torch.save({"y": y, "xs": xs, "owned": S.owned}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, "toeplitz_512_leaf64"), (4, "gauss2d_1024_leaf64"),
                                        (2, "utoeplitz_300_leaf32")])
def test_sharded_protocol_matches_reference(tmp_path, world, case):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, case, out), nprocs=world, join=True)
    res = torch.load(out)
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(res["y"].numpy().T, g["y"]) < 1e-13
    assert rel(res["xs"].numpy().T, g["xs"]) < 1e-10
