#!/bin/bash
# round 2, GPU run 4: copies-in-flight study (membench), merged-pair ring: correctness subset, traces, sweep
mkdir -p gpurun_out
./tools/membench r > gpurun_out/r2_04_membench.txt 2>&1; cat gpurun_out/r2_04_membench.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mat_vec or fused_up_gate or q8 or bias or multi_tensor" > gpurun_out/r2_04_pytest.log 2>&1; tail -3 gpurun_out/r2_04_pytest.log
for v in merged nomerge q8off; do
  case $v in merged) E="";; nomerge) E="B200Q_MERGE_PAIR=0";; q8off) E="B200Q_Q8_HANDOFF=0";; esac
  env LAYERS=6 B200Q_LIB_PATH=experiments/_variants/libb200q_trace.so $E timeout 300 python scripts/trace_decode.py > gpurun_out/r2_04_trace_$v.txt 2>&1
  tail -7 gpurun_out/r2_04_trace_$v.txt
done
timeout 1500 python scripts/sweep_decode.py > gpurun_out/r2_04_sweep.txt 2>&1
cat gpurun_out/r2_04_sweep.txt
