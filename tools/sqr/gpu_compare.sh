#!/bin/bash
# Default library vs the A/B variant build (make -C zkt_plonk_b200/csrc variant VARIANT_FLAGS=...) on one B200:
# parity tests of the variant first, then MSM timings of both.   gpurun --timeout 400 -- 'bash tools/sqr/gpu_compare.sh r01f'
tag=${1:-sqr}
mkdir -p gpurun_out
SQR=$PWD/zkt_plonk_b200/libzkb200_variant.so
ZKB200_LIB=$SQR timeout 200 python -m pytest tests/test_gpu_field.py tests/test_gpu_msm.py tests/test_gpu_prover.py tests/test_gpu_poly.py -q -x \
    > gpurun_out/${tag}_sqr_tests.log 2>&1
echo "variant tests rc=$?" | tee -a gpurun_out/${tag}_sqr_tests.log
for lib in default sqr; do   # 'sqr' = the variant (file names kept from the first experiment)
    if [ $lib = sqr ]; then export ZKB200_LIB=$SQR; else unset ZKB200_LIB; fi
    timeout 120 python tools/msm_time.py --sizes 18 20 22 --reps 7 --peaks > gpurun_out/${tag}_msm_${lib}.jsonl 2>&1
done
tail -4 gpurun_out/${tag}_sqr_tests.log
cat gpurun_out/${tag}_msm_default.jsonl gpurun_out/${tag}_msm_sqr.jsonl | cut -c1-400
