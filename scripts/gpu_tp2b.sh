#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_tp.py -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/t_tp.log 2>&1
echo "tp test rc=$?"; tail -12 gpurun_out/t_tp.log
for nccl in 0 1; do
B200Q_NCCL_REDUCE=$nccl timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$nccl bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_tp2_nccl$nccl.json 2> gpurun_out/bench_tp2_nccl$nccl.err
echo "bench nccl=$nccl rc=$?"; tail -1 gpurun_out/bench_tp2_nccl$nccl.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('tg',round(d['value'],1),'pp',round(d['pp512']['value']), d['config'].get('reduce'))
except Exception as e: print('ERR',e)"
tail -3 gpurun_out/bench_tp2_nccl$nccl.err
done
