#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 700 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
LD_LIBRARY_PATH=ik_llama_cpp_b200:oracle/_ref timeout -k 5 400 tests/backend_ops/test_mul_mat_backend > gpurun_out/backend_ops_harness.log 2>&1
echo "backend ops rc=$?"; grep -c " OK" gpurun_out/backend_ops_harness.log; grep "FAIL\|PASSED\|failed" gpurun_out/backend_ops_harness.log | head
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r20.json 2> gpurun_out/bench_r20.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r20.json')); print('default: tg', round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'pp', round(d['pp512']['value']), 'ppfrac', round(d['pp512']['roofline']['frac'],3), 'pp e2e', round(d['pp512']['e2e']['value']))"
for v in seg64 seg32; do
  B200Q_LIB_PATH=$PWD/experiments/_variants/libb200q_$v.so timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-pp > gpurun_out/bench_r20_$v.json 2> gpurun_out/bench_r20_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_r20_$v.json')); print('$v: tg', round(d['value'],1), 'frac', round(d['roofline']['frac'],3))" || tail -3 gpurun_out/bench_r20_$v.err
done
B200Q_LIB_PATH=$PWD/experiments/_variants/libb200q_seg64.so timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider -x -k "mat_vec or fused_up_gate or multi_tensor" 2>&1 | tail -2
