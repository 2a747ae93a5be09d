#!/bin/bash
# 2-GPU box: real NCCL path of the sharded HSS (tests + bench), and the BLR golden test
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_dist_gpu.py tests/test_blr_gpu.py -q -m gpu -x -k "dist or nccl or golden or sharded") > gpurun_out/n2_pytest.log 2>&1
tail -n 3 gpurun_out/n2_pytest.log | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1d_bench_n2.json 2> gpurun_out/r1d_bench_n2.err
tail -c 1500 gpurun_out/r1d_bench_n2.json
tail -n 3 gpurun_out/r1d_bench_n2.err
