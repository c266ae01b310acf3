#!/bin/bash
# round 2, call 2: new rows (f1 convbn, f3 mix, f4 fused optim), reduce kernels: tests + bench A/B
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_optim_mix.py -m gpu -x -q > $O/r2c_tests_new.log 2>&1; echo "new_rc=$?"
python -m pytest tests -m gpu -q > $O/r2c_tests_all.log 2>&1; echo "all_rc=$?"
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2c_bench_new.json 2> $O/r2c_bench_new.err; echo "bench_rc=$?"
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line --no-fused-optim > $O/r2c_bench_nofopt.json 2> $O/r2c_bench_nofopt.err
SENAS_NO_CONVBN=1 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2c_bench_noconvbn.json 2> $O/r2c_bench_noconvbn.err
SENAS_NO_MIX=1 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2c_bench_nomix.json 2> $O/r2c_bench_nomix.err
echo done
