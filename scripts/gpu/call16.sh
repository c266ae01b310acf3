#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_gpu_parity_r2.py tests/test_optim_mix.py -m gpu -q -k "graphed or fused_search or arch_grads or genotype" > $O/r2q_tests.log 2>&1; echo "tests_rc=$?"
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2q_bench_wf1.json 2> $O/r2q_bench_wf1.err
SENAS_WAVEFRONT=0 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2q_bench_wf0.json 2> $O/r2q_bench_wf0.err
echo done
