#!/bin/bash
# first GPU run: parity tests, bench, ncu
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-300 >> gpurun_out/gpu.txt
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "not gemm" -p no:cacheprovider > gpurun_out/t_decode.log 2>&1
echo "decode tests rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "gemm" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1
echo "gemm tests rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-pp > gpurun_out/bench_tg.json 2> gpurun_out/bench_tg.err
echo "bench tg rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench full rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/summary.txt
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:k_mmvq -s 20 -c 5 -o gpurun_out/prof_mmvq_r1 python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu --no-pp > gpurun_out/ncu_mmvq.log 2>&1
echo "ncu mmvq rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_decode.log; tail -5 gpurun_out/t_gemm.log; cat gpurun_out/bench_tg.json | cut -c1-1500; tail -3 gpurun_out/bench_tg.err
