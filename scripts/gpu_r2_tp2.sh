#!/bin/bash
# round 2, 2-GPU run: TP test (NVLS one-shot / two-shot bf16 / fused decode reduce), bench --gpus 2 with the correctness gate, TP decode timeline
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_tp2d_box.txt
timeout 600 python -m pytest tests/test_gpu_tp.py -m gpu -q -x -s > gpurun_out/r2_tp2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tp2d_pytest.log; tail -6 gpurun_out/r2_tp2d_pytest.log | cut -c1-600
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_tp2d_bench_$name.json 2> gpurun_out/r2_tp2d_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2_tp2d_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "tg", round(l["value"], 1), "pp512", round(l.get("pp512", {}).get("value", 0)), "gate", l["config"].get("tp_gate"))
except Exception as e:
    print("$name: no line", e); print(open("gpurun_out/r2_tp2d_bench_$name.err").read()[-1500:])
PY
}
run default X=1

run unfused B200Q_TP_FUSED=0
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_tp2d_bench_n1.json 2>/dev/null; python -c "
import json; l=json.loads(open('gpurun_out/r2_tp2d_bench_n1.json').read().strip().splitlines()[-1]); print('n1 tg', round(l['value'],1), 'pp512', round(l['pp512']['value']))"
env LAYERS=6 B200Q_LIB_PATH=experiments/_variants/libb200q_trace.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/trace_decode.py > gpurun_out/r2_tp2d_trace.txt 2>&1; tail -8 gpurun_out/r2_tp2d_trace.txt
