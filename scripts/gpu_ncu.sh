#!/bin/bash
# one ncu --set full capture of the leaf-class QR launch (variant from env), N = 262144
mkdir -p gpurun_out
T=${1:-ncu}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ulv_qr -s 0 -c 1 -o gpurun_out/${T} -f python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}.log 2>&1
tail -3 gpurun_out/${T}.log
ls -la gpurun_out/${T}.ncu-rep
