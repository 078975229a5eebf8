#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider -x -k "not gemm" > gpurun_out/t_all.log 2>&1
echo "tests rc=$?" >> gpurun_out/summary.txt
for cfg in "2 1" "2 0"; do set -- $cfg
B200Q_CTAS_PER_SM=$1 B200Q_PDL=$2 timeout -k 5 400 python bench.py --steps 20 --warmup 3 --no-cpu --no-pp > gpurun_out/bench_r9_cps$1_pdl$2.json 2> gpurun_out/bench_r9_cps$1_pdl$2.err
echo "bench cps=$1 pdl=$2 rc=$?" >> gpurun_out/summary.txt
done
B200Q_CTAS_PER_SM=2 B200Q_PDL=0 timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:k_mmvq -s 20 -c 5 -o gpurun_out/prof_mmvq_r9 python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu --no-pp > gpurun_out/ncu_mmvq.log 2>&1
echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r9_*.json")):
    try:
        d=json.load(open(f)); print(f, "tg", round(d["value"],1), "tok/s frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(f, "ERR", e, open(f.replace('.json','.err')).read()[-500:])
PY
