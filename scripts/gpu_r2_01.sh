#!/bin/bash
# round 2, GPU run 1: full GPU test-suite on the changed kernels, decode phase trace, decode variant sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_01_box.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_01_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_01_pytest.log
tail -5 gpurun_out/r2_01_pytest.log
LAYERS=4 timeout 300 python scripts/trace_decode.py > gpurun_out/r2_01_trace_product.txt 2>&1
LAYERS=4 B200Q_Q8_HANDOFF=0 timeout 300 python scripts/trace_decode.py > gpurun_out/r2_01_trace_q8off.txt 2>&1
for v in c15s64 c31s64cta1; do LAYERS=4 B200Q_LIB_PATH=experiments/_variants/libb200q_$v.so timeout 300 python scripts/trace_decode.py > gpurun_out/r2_01_trace_$v.txt 2>&1; done
timeout 1200 python scripts/sweep_decode.py > gpurun_out/r2_01_sweep.txt 2>&1
cat gpurun_out/r2_01_sweep.txt
tail -8 gpurun_out/r2_01_trace_product.txt
