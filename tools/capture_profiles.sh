#!/bin/bash
# Runs on the GPU box (gpurun): ncu --set full captures of the dominant kernels, exported as raw CSV next to the reports.
# Usage: bash tools/capture_profiles.sh <tag>      (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-prof}
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# MSM 2^20 (fixed-base tables): the bucket accumulation and the first two levels of the window reduction
$NCU -k regex:msm_accumulate_kernel --launch-skip 3 -c 1 -o $OUT/${TAG}_msm_acc -f python tools/msm_time.py --sizes 20 --reps 1 > $OUT/${TAG}_ncu_msm_acc.log 2>&1
$NCU -k regex:msm_wsum_level_kernel --launch-skip 51 -c 2 -o $OUT/${TAG}_msm_wsum -f python tools/msm_time.py --sizes 20 --reps 1 > $OUT/${TAG}_ncu_msm_wsum.log 2>&1
# coset NTT 2^22: both passes of one transform
$NCU -k regex:ntt_pass_kernel --launch-skip 4 -c 2 -o $OUT/${TAG}_ntt -f python tools/ntt_one.py > $OUT/${TAG}_ncu_ntt.log 2>&1
# the fused quotient kernel of a 2^20-gate proof
$NCU -k regex:quotient_kernel -c 1 -o $OUT/${TAG}_quotient -f python tools/prove_native_once.py 20 > $OUT/${TAG}_ncu_quotient.log 2>&1
for f in msm_acc msm_wsum ntt quotient; do
  if [ -f $OUT/${TAG}_$f.ncu-rep ]; then ncu -i $OUT/${TAG}_$f.ncu-rep --page raw --csv > $OUT/${TAG}_${f}_raw.csv 2>/dev/null; fi
done
ls -la $OUT | grep $TAG
