#!/bin/bash
# round 2, last 2-GPU run: unicast (coalesced peer stores) vs multicast variant of the tagged-slot exchange
mkdir -p gpurun_out
export B200Q_LIB_PATH=$PWD/experiments/_variants/libb200q_tp.so
run() { name=$1; shift; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-pp > gpurun_out/r2_tp2u_$name.json 2> gpurun_out/r2_tp2u_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2_tp2u_$name.json").read().strip().splitlines()[-1])
    print("$name", "tg", round(l["value"], 1), "gate", l["config"].get("tp_gate"))
except Exception as e:
    print("$name: no line", e); print(open("gpurun_out/r2_tp2u_$name.err").read()[-800:])
PY
}
run unicast X=1
run multicast B200Q_TP_UNICAST=0
run unicast_b X=1
timeout 100 python -m pytest tests/test_gpu_tp.py -m gpu -q -x > gpurun_out/r2_tp2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_tp2u_pytest.log | cut -c1-300
