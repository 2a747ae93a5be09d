#!/bin/bash
# round 2 evidence run: bench lines (HSS, BLR), launch lists, ncu --set full per kernel family.
# The .ncu-rep files are reduced to their raw-page CSV on the box (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
T=${1:-r2i}
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/${T}_ncu_${name} -f "$@" > gpurun_out/${T}_ncu_${name}.log 2>&1
  ncu -i gpurun_out/${T}_ncu_${name}.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_${name}_raw.csv 2>/dev/null
  ls -la gpurun_out/${T}_ncu_${name}.ncu-rep | awk '{print $5, $9}'
  rm -f gpurun_out/${T}_ncu_${name}.ncu-rep
  tail -n 2 gpurun_out/${T}_ncu_${name}.log | cut -c1-200
}
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 600 gpurun_out/${T}_bench_n1.json; tail -n 3 gpurun_out/${T}_bench_n1.err
timeout 900 python bench.py --workload blr --steps 3 --warmup 1 > gpurun_out/${T}_bench_blr.json 2> gpurun_out/${T}_bench_blr.err
tail -c 600 gpurun_out/${T}_bench_blr.json; tail -n 3 gpurun_out/${T}_bench_blr.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
wc -l gpurun_out/${T}_launches.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_blr.csv python bench.py --workload blr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_launches_blr.log 2>&1
wc -l gpurun_out/${T}_launches_blr.csv
cap factor "ulv_qr_kernel|ulv_eliminate" 0 2 python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline
cap apply "hss_up_kernel|hss_leaf_kernel|hss_down_kernel" 0 26 python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline
cap solve "ulv_fwd_pipe|ulv_bwd_pipe" 0 14 python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline
SB200_QR3=1 cap qr3 "ulv_qr3" 0 1 python bench.py --n 262144 --steps 1 --warmup 1 --no-cpu-baseline
cap blr "blr_schur_kernel|blr_getrf_kernel|id_cpqr_kernel|blr_trsm" 120 8 python bench.py --workload blr --blr-k 127 --steps 1 --warmup 1 --no-cpu-baseline
du -sh gpurun_out
