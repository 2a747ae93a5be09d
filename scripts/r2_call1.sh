#!/bin/bash
# round 2, call 1: full GPU suite (incl. same-generator parity at size), bench both arms
mkdir -p gpurun_out
T=r2a
nproc > gpurun_out/${T}_nproc.txt
(time timeout 1200 python -m pytest tests -q -m gpu -x --durations=8) > gpurun_out/${T}_pytest.log 2>&1
tail -n 25 gpurun_out/${T}_pytest.log | cut -c1-220
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 2500 gpurun_out/${T}_bench_n1.json; tail -n 5 gpurun_out/${T}_bench_n1.err
(time OMP_NUM_THREADS=1 timeout 900 python bench.py --impl reference --steps 5 --warmup 3) > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -c 1800 gpurun_out/${T}_bench_reference.json; tail -n 5 gpurun_out/${T}_bench_reference.err
