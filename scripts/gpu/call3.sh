#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_optim_mix.py -m gpu -q > $O/r2d_tests_new.log 2>&1; echo "new_rc=$?"
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "graphed or genotype" > $O/r2d_tests_graphed.log 2>&1; echo "graphed_rc=$?"
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2d_bench_new.json 2> $O/r2d_bench_new.err; echo "bench_rc=$?"
SENAS_NO_CONVBN=1 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2d_bench_noconvbn.json 2> $O/r2d_bench_noconvbn.err
SENAS_NO_MIX=1 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2d_bench_nomix.json 2> $O/r2d_bench_nomix.err
echo done
