#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_tp_shapes.py > gpurun_out/r2_10_tp_shapes.txt 2>&1; cat gpurun_out/r2_10_tp_shapes.txt | cut -c1-250 | tail -70
