#!/bin/bash
# round-2 evidence run (1 GPU): full GPU test suite, backend harness, bench (ours + reference arm), ncu launch list + two --set full captures.
# Outputs kept under the 64 MiB copy-back limit: raw CSV pages of the ncu reports; the .ncu-rep files only if small.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/r2_box.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/r2_box.txt
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gemm or host_buffer or extension" > gpurun_out/r2_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/r2_pytest_gemm.log >> gpurun_out/summary.txt
timeout -k 5 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_ours.json 2> gpurun_out/bench_ours.err
echo "bench rc=$?" >> gpurun_out/summary.txt
timeout -k 5 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 ncu --set full --import-source on --clock-control none -k regex:k_mmvq_ring -s 18 -c 9 -o gpurun_out/prof_mmvq_r2 python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu --no-pp > gpurun_out/ncu_mmvq.log 2>&1
echo "ncu mmvq rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 ncu --set full --import-source on --clock-control none -k regex:k_gemm_q -s 8 -c 8 -o gpurun_out/prof_gemmq_r2 python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?" >> gpurun_out/summary.txt
timeout -k 5 600 ./tests/backend_ops/test_mul_mat_backend > gpurun_out/r2_backend_ops_harness.log 2>&1; echo "harness rc=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/r2_backend_ops_harness.log >> gpurun_out/summary.txt
for r in prof_mmvq_r2 prof_gemmq_r2; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; done
python scripts/make_traffic.py gpurun_out/prof_mmvq_r2.ncu-rep gpurun_out/prof_gemmq_r2.ncu-rep gpurun_out/r2_traffic.json >> gpurun_out/summary.txt 2>&1
du -sm gpurun_out >> gpurun_out/summary.txt
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_gemmq_r2.ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_mmvq_r2.ncu-rep; fi
cat gpurun_out/summary.txt
# whatever GPU time is left: the rest of the GPU suite (the driver re-runs all of it at round end)
timeout -k 5 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_backend_ops.py -k "not gemm" > gpurun_out/r2_pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -3 gpurun_out/r2_pytest_rest.log
ls -la gpurun_out | head -30
