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
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:                      # tensor-parallel timeline (torchrun): every rank traces, rank 0 prints
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
model = bench.Model(be, torch, nl, tp=world, rank=rank)
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
for _ in range(int(os.environ.get("TRACE_WARMUP", "400"))):      # bring the SM clock up: the timestamps of a cold, short run are taken at idle clocks
    g.replay()
torch.cuda.synchronize()
L.b200q_debug_trace(3, None, 0)          # clear the accumulated min/max slots, then ONE traced replay at warm clocks
g.replay()
torch.cuda.synchronize()
n = model.launches_tg if (world == 1 or model.fused_tp) else 4 * nl + 1     # (unfused TP: the reduce kernels are not traced)
if rank != 0:
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
out = np.zeros((n, 8), np.uint64)
L.b200q_debug_trace(0, out.ctypes.data, n)
t = out.astype(np.int64)
t0 = t[0, 0]
names = ["qkv", "wo", "upgate", "down"]
print("launch  name    entry   wait_ret  prologue   done   | dur(entry->done) wait->done  gap(prev done -> wait_ret) | wait->x_loaded  ->quantised  ->barrier | barrier->1st unit (CTA0/w1)  first warp done  last warp done (after barrier)")
agg = {}
for i in range(n):
    nm = names[i % 4] if i < n - 1 else "head"
    e, w, p, d = (t[i, :4] - t0) / 1000.0
    gap = (t[i, 1] - t[i - 1, 3]) / 1000.0 if i else 0.0
    xl = (t[i, 4] - t[i, 1]) / 1000.0 if t[i, 4] else float("nan")
    qd = (t[i, 5] - t[i, 1]) / 1000.0 if t[i, 5] else float("nan")
    fu = (t[i, 7] - t[i, 2]) / 1000.0 if t[i, 7] else float("nan")
    fd = (((1 << 62) - t[i, 6]) - t[i, 2]) / 1000.0 if t[i, 6] else float("nan")
    print(f"{i:4d}  {nm:7s} {e:8.2f} {w:8.2f} {p:8.2f} {d:8.2f} | {d-e:8.2f} {d-w:8.2f} {gap:8.2f} | {xl:6.2f} {qd:6.2f} {p-w:6.2f} | {fu:6.2f} {fd:6.2f} {d-p:6.2f}")
    if i >= 4:
        agg.setdefault(nm, []).append((d - w, gap, p - w, d - p, fd))
print("total us", (t[n - 1, 3] - t0) / 1000.0)
print("median per kernel (layers >= 1): wait->done, gap, prologue, main(last), main(first warp)")
for nm, v in agg.items():
    a = np.median(np.array(v), axis=0)
    print(f"  {nm:7s} " + " ".join(f"{x:7.2f}" for x in a))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
