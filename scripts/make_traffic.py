#!/usr/bin/env python
"""DRAM traffic per bench step from the ncu --set full captures of a 2-layer run (scripts/gpu_profile_r2.sh), scaled to the 32-layer
model: profiles/r2_traffic.json, read by bench.py for roofline.traffic.  usage: make_traffic.py <mmvq.ncu-rep> <gemm.ncu-rep> <out.json>"""
import csv
import json
import subprocess
import sys


def rows(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr = r[0]
    out = []
    for x in r[2:]:
        d = dict(zip(hdr, x))
        unit = {h: u for h, u in zip(hdr, r[1])}
        def val(k):
            v = float(d[k].replace(",", "")); u = unit[k].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        out.append({"name": d["Kernel Name"][:60], "grid": d["Grid Size"], "bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                    "us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(unit["gpu__time_duration.sum"].lower(), 1)})
    return out


def main():
    mm, gg, out = sys.argv[1:4]
    n_layer = 32
    res = {"source": "ncu --set full --clock-control none, bench.py --layers 2 (scripts/gpu_profile_r2.sh), scaled to 32 layers"}
    r = rows(mm)
    if len(r) >= 9:
        step = r[-9:]                                   # one whole tg step of the 2-layer model: 8 layer launches + head
        head = max(step, key=lambda x: x["bytes"])
        layers = sum(x["bytes"] for x in step) - head["bytes"]
        res["tg"] = {"captured_launches": [{k: x[k] for k in ("name", "grid", "bytes", "us")} for x in step],
                     "dram_bytes_per_step": layers / 2 * n_layer + head["bytes"]}
    g = [x for x in rows(gg)]
    if len(g) >= 8:
        step = g[-8:]                                   # two layers of the pp512 step: 4 k_gemm_q launches per layer
        res["pp"] = {"captured_launches": [{k: x[k] for k in ("name", "grid", "bytes", "us")} for x in step],
                     "dram_bytes_per_step_gemm_only": sum(x["bytes"] for x in step) / 2 * n_layer}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "captured_launches"}) for k, v in res.items()}))


if __name__ == "__main__":
    main()
