#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pp_breakdown.py 512 2>&1 | tee gpurun_out/pp_breakdown_default.txt
for s in 2 8; do echo "== B200Q_GEMM_SPLIT=$s"; B200Q_GEMM_SPLIT=$s timeout 300 python scripts/pp_breakdown.py 512 2>&1 | grep -v "^convert" | tee gpurun_out/pp_breakdown_split$s.txt; done
