#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_optim_mix.py -m gpu -q -x -k "fused_search or dice or adam or sgd" > gpurun_out/r2v_tests.log 2>&1; echo "rc=$?"
