#!/bin/bash
# round 2, GPU run 11: ring path for K % 256 != 0 (bitnet rows), default-mix bench line, AVX-512 reference CPU baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "k_not or mat_vec or up_gate or q8 or qkv or bitnet" > gpurun_out/r2_11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_11_pytest.log; tail -4 gpurun_out/r2_11_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_11_bench.json 2> gpurun_out/r2_11_bench.err; tail -c 2500 gpurun_out/r2_11_bench.json; tail -5 gpurun_out/r2_11_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_11_bench_ref.json 2>/dev/null; tail -c 900 gpurun_out/r2_11_bench_ref.json
