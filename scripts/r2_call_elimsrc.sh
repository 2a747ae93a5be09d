#!/bin/bash
# source-level ncu capture of the leaf-class eliminate kernel (bench at N = 2^18: 1024 leaves)
mkdir -p gpurun_out
T=${1:-r3c}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ulv_eliminate_kernel -c 1 -o gpurun_out/${T}_elim -f python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
tail -n 2 gpurun_out/${T}_ncu.log | cut -c1-200
ncu -i gpurun_out/${T}_elim.ncu-rep --page raw --csv > gpurun_out/${T}_elim_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_elim.ncu-rep --page source --csv > gpurun_out/${T}_elim_source.csv 2>/dev/null
ls -la gpurun_out/${T}_elim* | awk '{print $5, $9}'
rm -f gpurun_out/${T}_elim.ncu-rep
