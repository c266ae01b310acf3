#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2n_tests_all.log 2>&1; echo "all_rc=$?"
python scripts/profile_cell.py bf16 16 3 256 > $O/r2n_cell256.log 2>&1
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2n_bench_pwmma.json 2> $O/r2n_bench_pwmma.err
SENAS_PW_MMA=0 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2n_bench_nopwmma.json 2> $O/r2n_bench_nopwmma.err
echo done
