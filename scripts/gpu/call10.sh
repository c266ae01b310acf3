#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2k_tests_all.log 2>&1; echo "all_rc=$?"
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2k_bench_dgmma.json 2> $O/r2k_bench_dgmma.err
SENAS_DGRAD_MMA=0 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2k_bench_nodgmma.json 2> $O/r2k_bench_nodgmma.err
SENAS_DW_PF=5 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2k_bench_pf5.json 2> $O/r2k_bench_pf5.err
SENAS_DW_PF=15 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2k_bench_pf15.json 2> $O/r2k_bench_pf15.err
SENAS_DW_PF=15 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/r2k_tests_pf15.log 2>&1; echo "pf_rc=$?"
echo done
