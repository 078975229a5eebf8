#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt; nvidia-smi topo -m >> gpurun_out/gpus2.txt 2>&1
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_tp2.json 2> gpurun_out/bench_tp2.err
echo "tp2 rc=$?"; cut -c1-700 gpurun_out/bench_tp2.json; tail -5 gpurun_out/bench_tp2.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_tp2.json")); print("tg", d["value"], "pp", d["pp512"]["value"])
except Exception as e: print("ERR", e)
PY
