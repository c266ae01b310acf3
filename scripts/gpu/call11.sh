#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2l_tests_all.log 2>&1; echo "all_rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2l_smoke.log 2>&1; echo "smoke_rc=$?"
timeout 900 python bench.py > $O/r2l_bench_default.json 2> $O/r2l_bench_default.err; echo "bench_rc=$?"
timeout 600 python bench.py --no-cpu --no-ref-gpu --no-fp32-line --arch-grads-only > $O/r2l_bench_archonly.json 2> $O/r2l_bench_archonly.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2l_bench_reference.json 2> $O/r2l_bench_reference.err; echo "ref_rc=$?"
echo done
