#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "fused_up_gate_gemm or gemm_multi" 2>&1 | tail -3
timeout 300 python scripts/pp_breakdown.py 512 2>&1 | grep -v "^convert\|single" | tee gpurun_out/pp_breakdown_epi2.txt
echo "== B200Q_FUSE_EPILOGUE=0"; B200Q_FUSE_EPILOGUE=0 timeout 300 python scripts/pp_breakdown.py 512 2>&1 | grep "fused_up_gate\|layer total" | tee gpurun_out/pp_breakdown_noepi.txt
