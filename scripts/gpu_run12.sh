#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 ./tests/backend_ops/test_mul_mat_backend > gpurun_out/backend_ops.log 2>&1
echo "backend ops rc=$?"; grep -E "FAIL|PASSED|FAILED|backend " gpurun_out/backend_ops.log | head -20
timeout -k 5 900 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/t_all.log
