#!/bin/bash
# One GPU call that re-validates the tree on a fresh B200 box: the tests added last first (so that a time limit cuts
# the least informative part), then the whole GPU suite, smoke() and the default bench line.  Logs go to gpurun_out/.
#   gpurun --timeout 600 -- 'bash tools/round_check.sh r01e'
tag=${1:-check}
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_msm.py tests/test_gpu_sharded.py -q -x -m gpu -k "pair_rounds or replicated or push_finish" > gpurun_out/${tag}_new_tests.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/${tag}_new_tests.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/${tag}_gpu_tests.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
timeout 200 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
tail -n 3 gpurun_out/${tag}_new_tests.log gpurun_out/${tag}_gpu_tests.log gpurun_out/${tag}_smoke.log
