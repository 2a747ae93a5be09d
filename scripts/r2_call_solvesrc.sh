#!/bin/bash
# source-level ncu capture of the leaf-class forward / backward sweep kernels (bench at N = 2^18: 1024 leaves, 10 non-root classes)
mkdir -p gpurun_out
T=${1:-r3f}
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/${T}_$1 -f python bench.py --n 262144 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${T}_$1.log 2>&1
  tail -n 1 gpurun_out/${T}_$1.log | cut -c1-200
  ncu -i gpurun_out/${T}_$1.ncu-rep --page raw --csv > gpurun_out/${T}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${T}_$1.ncu-rep --page source --csv > gpurun_out/${T}_$1_source.csv 2>/dev/null
  rm -f gpurun_out/${T}_$1.ncu-rep
  grep -o "launch__grid_size[^,]*,[^,]*,[^,]*" gpurun_out/${T}_$1_raw.csv | head -2
}
cap fwd ulv_fwd_pipe_kernel 0
cap bwd ulv_bwd_pipe_kernel 9
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
