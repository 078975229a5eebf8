#!/usr/bin/env python
"""Phase timeline of the decode graph (debug): per mat-vec launch, %globaltimer at entry / after griddepcontrol.wait /
prologue done / last consumer done (CTA 0), from one CUDA-graph replay of bench.py's tg step."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ik_llama_cpp_b200 import backend as be, _lib

L = _lib.lib()
L.b200q_debug_trace.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
nl = int(os.environ.get("LAYERS", "8"))
torch.cuda.set_device(0)
model = bench.Model(be, torch, nl)
model.alloc(1)
model.x.normal_()
L.b200q_debug_trace(1, None, 0)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    model.step_tg(); model.step_tg()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
L.b200q_debug_trace(2, None, 0)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    model.step_tg()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
n = model.launches_tg
out = np.zeros((n, 4), np.uint64)
L.b200q_debug_trace(0, out.ctypes.data, n)
t = out.astype(np.int64)
t0 = t[0, 0]
names = ["qkv", "wo", "upgate", "down"]
print("launch  name    entry   wait_ret  prologue   done   | dur(entry->done) wait->done  gap(prev done -> wait_ret)")
for i in range(n):
    nm = names[i % 4] if i < n - 1 else "head"
    e, w, p, d = (t[i] - t0) / 1000.0
    gap = (t[i, 1] - t[i - 1, 3]) / 1000.0 if i else 0.0
    print(f"{i:4d}  {nm:7s} {e:8.2f} {w:8.2f} {p:8.2f} {d:8.2f} | {d-e:8.2f} {d-w:8.2f} {gap:8.2f}")
print("total us", (t[n - 1, 3] - t0) / 1000.0)
