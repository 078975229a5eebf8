#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 ./tests/backend_ops/test_mul_mat_backend > gpurun_out/backend_ops.log 2>&1
echo "backend ops rc=$?"; tail -45 gpurun_out/backend_ops.log
