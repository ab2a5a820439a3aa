#!/bin/bash
# ncu evidence for round 2's kernels (run only after the same commands exited 0 without ncu in this call):
#   1. launch list of the bench command (per-launch gpu__time_duration; cold-cache and serialised: compare shares)
#   2. `--set full` of msm_accumulate_kernel at 2^20 (roofline.traffic, pipe utilisation)
#   3. `--set full` of the radix-4 NTT pass kernel on a 2^22 coset transform (both passes)
# Usage on the GPU box: bash tools/capture_r02.sh <tag>     (outputs under gpurun_out/<tag>_*)
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 2 --warmup 3 --no-prove --no-cpu-baseline --no-sweep > $OUT/${TAG}_plain_bench.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-prove --no-cpu-baseline --no-sweep > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/msm_time.py --sizes 20 --reps 1 > $OUT/${TAG}_plain_msm.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel --launch-skip 3 -c 1 \
    -o $OUT/${TAG}_msm_acc -f python tools/msm_time.py --sizes 20 --reps 1 > $OUT/${TAG}_ncu_msm_acc.log 2>&1
echo "accumulate capture rc=$?"
python tools/ntt_one.py > $OUT/${TAG}_plain_ntt.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel_r4 --launch-skip 4 -c 2 \
    -o $OUT/${TAG}_ntt_r4 -f python tools/ntt_one.py > $OUT/${TAG}_ncu_ntt.log 2>&1
echo "ntt capture rc=$?"
for k in msm_acc ntt_r4; do
    [ -f $OUT/${TAG}_$k.ncu-rep ] && ncu -i $OUT/${TAG}_$k.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
done
ls -la $OUT | grep $TAG
