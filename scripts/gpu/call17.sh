#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line --defer-wgrad > $O/r2r_bench_defer.json 2> $O/r2r_bench_defer.err
python -m pytest tests -m gpu -q > $O/r2r_tests_all.log 2>&1; echo "all_rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2r_smoke.log 2>&1; echo "smoke_rc=$?"
timeout 900 python bench.py > $O/r2r_bench_default.json 2> $O/r2r_bench_default.err; echo "bench_rc=$?"
echo done
