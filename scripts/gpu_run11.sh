#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm" > gpurun_out/t_gemm.log 2>&1
echo "gemm tests rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r11.json 2> gpurun_out/bench_r11.err
echo "bench rc=$?" >> gpurun_out/summary.txt
B200Q_FUSED_GEMM=0 timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r11_unfused.json 2> gpurun_out/bench_r11_unfused.err
cat gpurun_out/summary.txt; tail -8 gpurun_out/t_gemm.log; python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r11*.json")):
    try:
        d=json.load(open(f)); print(f, "tg", round(d["value"],1), "pp", round(d["pp512"]["value"]), round(d["pp512"]["roofline"]["frac"],3), "e2e pp", round(d["pp512"]["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-800:])
PY
