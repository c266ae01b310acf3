#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python bench.py --no-cpu --no-ref-gpu --no-fp32-line --arch-grads-only > $O/r2s_bench_archonly.json 2> $O/r2s_bench_archonly.err
timeout 300 python bench.py --no-cpu --no-ref-gpu --no-fp32-line > $O/r2s_bench_plain.json 2> $O/r2s_bench_plain.err
echo done
