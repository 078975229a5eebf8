#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "IQ2_KS or IQ3_KS or MXFP4" 2>&1 | tail -8
LD_LIBRARY_PATH=ik_llama_cpp_b200:oracle/_ref timeout -k 5 500 tests/backend_ops/test_mul_mat_backend > gpurun_out/backend_ops_harness.log 2>&1
echo "backend ops rc=$?"; grep -c " OK" gpurun_out/backend_ops_harness.log; grep "FAIL\|PASSED\|failed" gpurun_out/backend_ops_harness.log | head
