#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_tp.py -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -12
for f in 1; do
  B200Q_TP_FUSED=$f timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r23_tp2_f$f.json 2> gpurun_out/bench_r23_tp2_f$f.err
  echo "rc=$?"
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r23_tp2_f$f.json') if l.startswith('{')][-1]); print('TP2 fused=$f: tg', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'pp', round(d['pp512']['value']))" || tail -5 gpurun_out/bench_r23_tp2_f$f.err
done
