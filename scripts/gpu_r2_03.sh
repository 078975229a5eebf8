#!/bin/bash
# round 2, GPU run 3: decode phase trace at warm clocks + decode variant sweep (incl. the round-1 library as A/B baseline)
mkdir -p gpurun_out
for v in product q8off c15s64 c31s64cta1 c23s64cta1; do
  case $v in product) E="";; q8off) E="B200Q_Q8_HANDOFF=0";; *) E="B200Q_LIB_PATH=experiments/_variants/libb200q_$v.so";; esac
  env LAYERS=6 $E timeout 300 python scripts/trace_decode.py > gpurun_out/r2_03_trace_$v.txt 2>&1
  tail -7 gpurun_out/r2_03_trace_$v.txt
done
timeout 1500 python scripts/sweep_decode.py > gpurun_out/r2_03_sweep.txt 2>&1
cat gpurun_out/r2_03_sweep.txt
