#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2i_tests_all.log 2>&1; echo "all_rc=$?"
python scripts/profile_cell.py bf16 16 3 256 > $O/r2i_cell256.log 2>&1
SENAS_WGRAD_MMA=0 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2i_bench_nowm.json 2> $O/r2i_bench_nowm.err
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2i_bench_wm.json 2> $O/r2i_bench_wm.err
SENAS_GATHER_MMA=1 python -m pytest tests -m gpu -q > $O/r2i_tests_all_gmma.log 2>&1; echo "gmma_rc=$?"
echo done
