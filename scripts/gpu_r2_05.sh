#!/bin/bash
# round 2, GPU run 5: full GPU suite (46 types, long-row ring, q8 tail hand-off), decode traces, sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_05_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_05_pytest.log; tail -15 gpurun_out/r2_05_pytest.log
for v in product q8off gridfull nopdl; do
  case $v in product) E="";; q8off) E="B200Q_Q8_HANDOFF=0";; gridfull) E="B200Q_GRID_FULL=1";; nopdl) E="B200Q_PDL=0";; esac
  env LAYERS=6 B200Q_LIB_PATH=experiments/_variants/libb200q_trace.so $E timeout 300 python scripts/trace_decode.py > gpurun_out/r2_05_trace_$v.txt 2>&1
  tail -7 gpurun_out/r2_05_trace_$v.txt
done
timeout 1500 python scripts/sweep_decode.py product r1 > gpurun_out/r2_05_sweep.txt 2>&1
cat gpurun_out/r2_05_sweep.txt
