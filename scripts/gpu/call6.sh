#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_dp_graphed_gpu.py -m gpu -q > $O/r2g_dp_test.log 2>&1; echo "dp_rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2g_bench_2gpu.json 2> $O/r2g_bench_2gpu.err; echo "bench2_rc=$?"
tail -3 $O/r2g_bench_2gpu.err
echo done
