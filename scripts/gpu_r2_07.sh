#!/bin/bash
# round 2, GPU run 7: full GPU suite (MoE MUL_MAT_ID, int8 IQ2_BN GEMM, ADD), harness, inter-kernel gap experiments, full bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_07_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_07_pytest.log; tail -15 gpurun_out/r2_07_pytest.log
timeout 900 ./tests/backend_ops/test_mul_mat_backend > gpurun_out/r2_07_backend_ops.log 2>&1; echo "harness rc=$?"; grep -c " OK" gpurun_out/r2_07_backend_ops.log; grep "FAIL\|PASSED\|FAILED\|not supp" gpurun_out/r2_07_backend_ops.log | head -20
timeout 900 python scripts/sweep_decode.py product prefillafter latetrig c23cta1 > gpurun_out/r2_07_sweep.txt 2>&1
cat gpurun_out/r2_07_sweep.txt
for v in trace tr_prefillafter tr_latetrig; do
  env LAYERS=6 B200Q_LIB_PATH=experiments/_variants/libb200q_$v.so timeout 300 python scripts/trace_decode.py > gpurun_out/r2_07_trace_$v.txt 2>&1
  tail -7 gpurun_out/r2_07_trace_$v.txt
done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_07_bench.json 2> gpurun_out/r2_07_bench.err; tail -c 1800 gpurun_out/r2_07_bench.json
