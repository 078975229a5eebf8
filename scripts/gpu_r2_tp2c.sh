#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/debug_tp_gate.py > gpurun_out/r2_tp2c_debug.txt 2>&1; grep "rank\|Error\|error" gpurun_out/r2_tp2c_debug.txt | cut -c1-330 | head -40
