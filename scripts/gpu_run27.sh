#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout -k 5 200 python -m pytest tests/test_gpu_tp.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -2
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r27_tp$N.json 2> gpurun_out/bench_r27_tp$N.err
echo "rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r27_tp$N.json') if l.startswith('{')][-1]); print('TP$N: tg', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'pp', round(d['pp512']['value']), d['config']['reduce'][:40])" || tail -8 gpurun_out/bench_r27_tp$N.err
