#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2h_tests_all.log 2>&1; echo "all_rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2h_smoke.log 2>&1; echo "smoke_rc=$?"
SENAS_GATHER_MMA=1 timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2h_bench_mma.json 2> $O/r2h_bench_mma.err
SENAS_GATHER_MMA=1 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "genotype" > $O/r2h_tests_mma_geno.log 2>&1; echo "mma_geno_rc=$?"
timeout 900 python bench.py > $O/r2h_bench_default.json 2> $O/r2h_bench_default.err; echo "bench_rc=$?"
echo done
