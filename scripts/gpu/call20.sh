#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "graphed" > gpurun_out/r2u_tests_graphed.log 2>&1; echo "rc=$?"
