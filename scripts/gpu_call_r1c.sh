#!/bin/bash
# round-1 (session c) GPU call: new tests, bench N=1, launch list, one ncu --set full capture
mkdir -p gpurun_out
T=r1c
(time timeout 300 python -m pytest tests/test_cpp_mirror.py tests/test_float_gpu.py tests/test_blr_gpu.py -q -m gpu -x -k "cpp or float or transposed") > gpurun_out/${T}_pytest_new.log 2>&1
tail -5 gpurun_out/${T}_pytest_new.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 2500 gpurun_out/${T}_bench_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
wc -l gpurun_out/${T}_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ulv_qr|ulv_eliminate" -s 0 -c 2 -o gpurun_out/${T}_ncu -f python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_ncu.ncu-rep
