#!/usr/bin/env python
"""Per-CTA timeline of decode launches (tuning builds with -DB200Q_TRACE_FINE=1): when does each CTA finish, on which SM, with how many units.
Shows how unevenly the SMs are served by the memory system (static split of the units over the CTAs)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ik_llama_cpp_b200 import backend as be, _lib

L = _lib.lib()
L.b200q_debug_trace.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
nl = int(os.environ.get("LAYERS", "4"))
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
for _ in range(400):
    g.replay()
torch.cuda.synchronize()
L.b200q_debug_trace(3, None, 0)
g.replay()
torch.cuda.synchronize()
n = model.launches_tg
out = np.zeros((n, 8), np.uint64)
L.b200q_debug_trace(0, out.ctypes.data, n)
t = out.astype(np.int64)
names = ["qkv", "wo", "upgate", "down"]
for i in range(4 * (nl - 1), n):
    nm = names[i % 4] if i < n - 1 else "head"
    c = np.zeros((512, 4), np.uint64)
    L.b200q_debug_trace(4, c.ctypes.data, i)
    c = c.astype(np.int64)
    live = c[:, 0] > 0
    end = (c[live, 0] - t[i, 2]) / 1000.0
    first = (((1 << 62) - c[live, 3]) - t[i, 2]) / 1000.0
    smid = c[live, 1]; units = c[live, 2]
    print(f"launch {i} {nm}: {live.sum()} CTAs, units/CTA {units.min()}..{units.max()}, CTA end (us after CTA0's barrier): min {end.min():.2f} p10 {np.percentile(end,10):.2f} "
          f"median {np.median(end):.2f} p90 {np.percentile(end,90):.2f} max {end.max():.2f}; first-warp end median {np.median(first):.2f}")
    # by SM: both CTAs of an SM, and by groups of 2 SMs (TPC) / 18-20 SMs
    order = np.argsort(smid, kind="stable")
    per_sm = {}
    for sm, e, u in zip(smid[order], end[order], units[order]):
        per_sm.setdefault(int(sm), []).append((float(e), int(u)))
    line = []
    for sm in sorted(per_sm):
        line.append(f"{sm}:" + "/".join(f"{e:.1f}" for e, _ in per_sm[sm]))
    print("   per SM (end of each resident CTA): " + " ".join(line))
    # time per unit by SM
    tpu = np.array([np.mean([e / max(u, 1) for e, u in per_sm[sm]]) for sm in sorted(per_sm)])
    print(f"   us per unit by SM: min {tpu.min():.3f} median {np.median(tpu):.3f} max {tpu.max():.3f}")
