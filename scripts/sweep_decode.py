#!/usr/bin/env python
"""Run bench.py's tg step (tok/s only) for the product library and every tuning variant under experiments/_variants,
each in its own process (the library is chosen at import time), optionally with extra environment knobs."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
configs = [("product", None, {}), ("product pfoff", None, {"B200Q_PREFETCH_NEXT": "0"}), ("product q8off", None, {"B200Q_Q8_HANDOFF": "0"})]
vd = os.path.join(ROOT, "experiments", "_variants")
for f in sorted(os.listdir(vd)) if os.path.isdir(vd) else []:
    if f.endswith(".so"):
        name = f[len("libb200q_"):-3]
        configs.append((name, os.path.join(vd, f), {}))
        if name not in ("r1", "trace"):
            configs.append((name + " q8off", os.path.join(vd, f), {"B200Q_Q8_HANDOFF": "0"}))
# extra product configurations from the command line: name:ENV=val,ENV2=val
for arg in [a for a in sys.argv[1:] if ":" in a]:
    nm, kv = arg.split(":", 1)
    configs.append((nm, None, dict(x.split("=", 1) for x in kv.split(","))))
only = [a.split(":", 1)[0] for a in sys.argv[1:]]
for name, lib, env in configs:
    if only and not any(o == name or (":" not in o and o in name) for o in only):
        continue
    e = dict(os.environ); e.update(env)
    if lib: e["B200Q_LIB_PATH"] = lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-pp", "--no-cpu", "--no-mix", "--steps", "20", "--warmup", "3"], env=e, capture_output=True, text=True)
    try:
        line = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{name:28s} tg {line['value']:8.1f} tok/s  frac {line['roofline']['frac']:.3f}  e2e {line['e2e']['value']:8.1f}", flush=True)
    except Exception:
        print(f"{name:28s} FAILED rc={r.returncode}: {r.stderr[-400:]}", flush=True)
