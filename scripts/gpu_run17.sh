#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 700 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm or fused or k_not or multi" > gpurun_out/t_gemm.log 2>&1
echo "gemm tests rc=$?"; tail -25 gpurun_out/t_gemm.log
LD_LIBRARY_PATH=ik_llama_cpp_b200:oracle/_ref timeout -k 5 400 tests/backend_ops/test_mul_mat_backend > gpurun_out/backend_ops_harness.log 2>&1
echo "backend ops rc=$?"; grep -c " OK" gpurun_out/backend_ops_harness.log; grep "FAIL\|PASSED\|failed" gpurun_out/backend_ops_harness.log | head
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r17.json 2> gpurun_out/bench_r17.err
tail -3 gpurun_out/bench_r17.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r17.json')); print('tg', round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1)); print({k:v for k,v in d.items() if 'pp' in k})"
