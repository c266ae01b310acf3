#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r2m_bench_4gpu.json 2> $O/r2m_bench_4gpu.err; echo "bench4_rc=$?"
tail -3 $O/r2m_bench_4gpu.err
echo done
