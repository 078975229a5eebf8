#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-pp > gpurun_out/bench_r25.json 2> gpurun_out/bench_r25.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r25.json')); print('default: tg', round(d['value'],1), 'frac', round(d['roofline']['frac'],3))"
export B200Q_LIB_PATH=$PWD/experiments/_variants/libb200q_imadshr.so
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "not gemm" 2>&1 | tail -2
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-pp > gpurun_out/bench_r25_imadshr.json 2> gpurun_out/bench_r25_imadshr.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r25_imadshr.json')); print('imadshr: tg', round(d['value'],1), 'frac', round(d['roofline']['frac'],3))"
LAYERS=2 timeout 300 python scripts/trace_decode.py 2>&1 | tail -11
