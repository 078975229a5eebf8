#!/bin/bash
# round 2, GPU run 9: cross-CTA dynamic work claiming in the decode ring kernel; n = 1 on tensor-parallel shard shapes (NaN hunt)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mat_vec or up_gate or q8 or llama or qkv or k_not or claiming" > gpurun_out/r2_09_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_09_pytest.log; tail -4 gpurun_out/r2_09_pytest.log
timeout 900 python scripts/sweep_decode.py product "dynoff:B200Q_DYN=0" "dyn40:B200Q_DYN_STATIC=0.4" "dyn80:B200Q_DYN_STATIC=0.8" "q8off_dyn:B200Q_Q8_HANDOFF=0" "q8off_dynoff:B200Q_Q8_HANDOFF=0,B200Q_DYN=0" "q8off_dyn40:B200Q_Q8_HANDOFF=0,B200Q_DYN_STATIC=0.4" > gpurun_out/r2_09_sweep.txt 2>&1
cat gpurun_out/r2_09_sweep.txt
for v in "" "B200Q_Q8_HANDOFF=0"; do
  env LAYERS=3 B200Q_LIB_PATH=experiments/_variants/libb200q_trace.so $v timeout 300 python scripts/trace_cta.py > gpurun_out/r2_09_trace_cta_${v:-default}.txt 2>&1; grep -v "per SM" gpurun_out/r2_09_trace_cta_${v:-default}.txt | tail -10
done
