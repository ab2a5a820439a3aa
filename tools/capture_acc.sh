#!/bin/bash
# ncu evidence for the current bucket-accumulation kernel (run after the same commands exited 0 without ncu):
#   1. launch list of the bench command (per-launch gpu__time_duration, cold-cache and serialised: shares, not absolutes)
#   2. one `--set full` capture of msm_accumulate_kernel at 2^20 (DRAM traffic for roofline.traffic, pipe utilisation)
# Usage on the GPU box: bash tools/capture_acc.sh <tag>     (outputs under gpurun_out/<tag>_*)
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-prove --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel --launch-skip 3 -c 1 \
    -o $OUT/${TAG}_msm_acc -f python tools/msm_time.py --sizes 20 --reps 1 > $OUT/${TAG}_ncu_msm_acc.log 2>&1
echo "full capture rc=$?"
[ -f $OUT/${TAG}_msm_acc.ncu-rep ] && ncu -i $OUT/${TAG}_msm_acc.ncu-rep --page raw --csv > $OUT/${TAG}_msm_acc_raw.csv 2>/dev/null
ls -la $OUT | grep $TAG
