#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "tests rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
echo "bench rc=$?" >> gpurun_out/summary.txt
B200Q_PDL=0 timeout -k 5 400 python bench.py --steps 20 --warmup 3 --no-cpu --no-pp > gpurun_out/bench_r2_nopdl.json 2> gpurun_out/bench_r2_nopdl.err
echo "bench nopdl rc=$?" >> gpurun_out/summary.txt
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu --no-pp > gpurun_out/ncu_launch.log 2>&1
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:k_mmvq -s 20 -c 5 -o gpurun_out/prof_mmvq_r2 python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu --no-pp > gpurun_out/ncu_mmvq.log 2>&1
echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/t_all.log; python - <<'PY'
import json
for f in ("gpurun_out/bench_r2.json","gpurun_out/bench_r2_nopdl.json"):
    try:
        d=json.load(open(f)); print(f, "tg", round(d["value"],1), "tok/s frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "pp", d.get("pp512",{}).get("value"), d.get("pp512",{}).get("roofline",{}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/bench_r2.err
