#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2j_bench_8gpu.json 2> $O/r2j_bench_8gpu.err; echo "bench8_rc=$?"
tail -5 $O/r2j_bench_8gpu.err
echo done
