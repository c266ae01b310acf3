#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
SENAS_STEM_CONVBN=1 python -m pytest tests -m gpu -q > $O/r2t_tests_all_stem.log 2>&1; echo "all_rc=$?"
SENAS_STEM_CONVBN=1 timeout 200 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2t_bench_stem1.json 2> $O/r2t_bench_stem1.err
timeout 200 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2t_bench_stem0.json 2> $O/r2t_bench_stem0.err
echo done
