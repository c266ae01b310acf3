#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_optim_mix.py -m gpu -q -k "fused_search or arch_grads" > $O/r2e_tests_new.log 2>&1; echo "new_rc=$?"
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "graphed" > $O/r2e_tests_graphed.log 2>&1; echo "graphed_rc=$?"
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line --arch-grads-only > $O/r2e_bench_archonly.json 2> $O/r2e_bench_archonly.err; echo "bench_rc=$?"
echo done
