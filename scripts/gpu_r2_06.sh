#!/bin/bash
# round 2, GPU run 6: L2 warm-up of the next launch's weights (decode), rewritten backend plug (harness), full GPU suite
mkdir -p gpurun_out
for v in product pfoff; do
  case $v in product) E="";; pfoff) E="B200Q_PREFETCH_NEXT=0";; esac
  env LAYERS=6 B200Q_LIB_PATH=experiments/_variants/libb200q_trace.so $E timeout 300 python scripts/trace_decode.py > gpurun_out/r2_06_trace_$v.txt 2>&1
  tail -7 gpurun_out/r2_06_trace_$v.txt
done
timeout 900 python scripts/sweep_decode.py product > gpurun_out/r2_06_sweep.txt 2>&1
cat gpurun_out/r2_06_sweep.txt
timeout 900 ./tests/backend_ops/test_mul_mat_backend > gpurun_out/r2_06_backend_ops.log 2>&1; echo "harness rc=$?"; grep -c " OK" gpurun_out/r2_06_backend_ops.log; grep "FAIL\|PASSED\|FAILED" gpurun_out/r2_06_backend_ops.log | head -20
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_06_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_06_pytest.log; tail -5 gpurun_out/r2_06_pytest.log
