#!/bin/bash
# A/B of the quotient kernel: default build vs `make variant VARIANT_FLAGS=-DZKB_QUOT_FUSED=1`.  Parity of the variant
# (polynomial + prover suites against the oracle), then the kernel's device time and an output checksum from both.
tag=${1:-quot}
mkdir -p gpurun_out
VAR=$PWD/zkt_plonk_b200/libzkb200_variant.so
ZKB200_LIB=$VAR timeout 150 python -m pytest tests/test_gpu_poly.py tests/test_gpu_prover.py -q -x > gpurun_out/${tag}_variant_tests.log 2>&1
echo "variant tests rc=$?" | tee -a gpurun_out/${tag}_variant_tests.log
timeout 60 python tools/sqr/time_quotient.py 20 > gpurun_out/${tag}_quotient_default.json 2>&1
ZKB200_LIB=$VAR timeout 60 python tools/sqr/time_quotient.py 20 > gpurun_out/${tag}_quotient_variant.json 2>&1
tail -n 3 gpurun_out/${tag}_variant_tests.log; cat gpurun_out/${tag}_quotient_default.json gpurun_out/${tag}_quotient_variant.json
