#!/bin/bash
# round 2, N-GPU run (NG = 4 or 8): TP test for every world size that fits, bench.py --gpus N with the correctness gate, decode timeline of rank 0.
# Uses the bench-types-only build (B200Q_LIB_PATH) so that the snapshot pushed to the multi-GPU box stays small.
NG=${NG:-4}
mkdir -p gpurun_out
export B200Q_LIB_PATH=$PWD/experiments/_variants/libb200q_tp.so
nvidia-smi -L > gpurun_out/r2_tp${NG}_box.txt
timeout 600 python -m pytest tests/test_gpu_tp.py -m gpu -q -x -s > gpurun_out/r2_tp${NG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tp${NG}_pytest.log; tail -4 gpurun_out/r2_tp${NG}_pytest.log | cut -c1-300
run() { n=$1; name=$2; shift; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_tp${NG}_bench_$name.json 2> gpurun_out/r2_tp${NG}_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2_tp${NG}_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "tg", round(l["value"], 1), "pp512", round(l.get("pp512", {}).get("value", 0)), "gate", l["config"].get("tp_gate"))
except Exception as e:
    print("$name: no line", e); print(open("gpurun_out/r2_tp${NG}_bench_$name.err").read()[-1500:])
PY
}
run $NG n$NG X=1
if [ "$NG" -gt 2 ]; then run 2 n2 X=1; fi
if [ "$NG" -gt 4 ]; then run 4 n4 X=1; fi
env LAYERS=6 B200Q_LIB_PATH=$PWD/experiments/_variants/libb200q_trace.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 scripts/trace_decode.py > gpurun_out/r2_tp${NG}_trace.txt 2>&1; tail -7 gpurun_out/r2_tp${NG}_trace.txt
if [ "$NG" = "2" ]; then run 2 n2_norowbuf B200Q_TP_ROWBUF=0; fi
