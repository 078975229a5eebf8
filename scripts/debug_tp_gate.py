#!/usr/bin/env python
"""torchrun, 2 GPUs: the tg (n = 1) tensor-parallel step of bench.py, layer by layer, against shards rebuilt on every rank: where does it go wrong?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from ik_llama_cpp_b200 import backend as be

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
N_EMBD = bench.N_EMBD
model = bench.Model(be, torch, 2, tp=world, rank=rank)
shards = [model if r == rank else bench.Model(be, torch, 2, tp=world, rank=r, collective=False) for r in range(world)]
model.alloc(1)
gen = torch.Generator(device="cuda"); gen.manual_seed(777)
x0 = torch.randn(1, N_EMBD, device="cuda", generator=gen)

def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
def nm(t, ref):
    t, ref = t.double(), ref.double()
    return f"nan {int(torch.isnan(t).sum())} absmax {float(torch.nan_to_num(t).abs().max()):.4g} nmse {float(((t - ref) ** 2).sum() / (ref ** 2).sum()):.3g}"

# reference, layer by layer (every rank computes the same thing)
refs = []
x = x0.clone()
for li in range(2):
    part = torch.zeros(1, N_EMBD, dtype=torch.float64, device="cuda")
    for sh in shards:
        L = sh.layers[li]; part += be.mul_mat(L["wo"], be.mul_mat(L["wq"], x)).double()
    h = part.float()
    part = torch.zeros(1, N_EMBD, dtype=torch.float64, device="cuda")
    for sh in shards:
        L = sh.layers[li]; part += be.mul_mat(L["down"], be.fused_up_gate(L["up"], L["gate"], h, "silu")).double()
    x = part.float()
    refs.append((h, x))
torch.cuda.synchronize(); dist.barrier()

# (1) unfused: plain kernels + NVLS all-reduce, op by op
r = model.reducer
x = x0.clone()
for li, L in enumerate(model.layers):
    be.mul_mat_multi([L["wq"], L["wk"], L["wv"]], x, [model.q, model.kk, model.v])
    be.mul_mat(L["wo"], model.q, out=model.h)
    hp = model.h.clone(); r.all_reduce(model.h)
    allp = [torch.empty_like(hp) for _ in range(world)]; dist.all_gather(allp, hp)
    say(f"unfused layer {li}: h partial nan {int(torch.isnan(hp).sum())}; NVLS sum vs NCCL-gathered sum: {nm(model.h, sum(p.double() for p in allp))}; vs ref {nm(model.h, refs[li][0])}")
    be.fused_up_gate(L["up"], L["gate"], model.h, "silu", out=model.a, q8_out=model.q8a)
    be.mul_mat(L["down"], model.a, out=model.x2, q8_in=model.q8a)
    xp = model.x2.clone(); r.all_reduce(model.x2)
    allp = [torch.empty_like(xp) for _ in range(world)]; dist.all_gather(allp, xp)
    say(f"unfused layer {li}: x2 partial nan {int(torch.isnan(xp).sum())} (q8 hand-off valid {model.q8a.valid}); NVLS sum vs gathered {nm(model.x2, sum(p.double() for p in allp))}; vs ref {nm(model.x2, refs[li][1])}")
    x = model.x2
torch.cuda.synchronize(); dist.barrier()

# (2) fused: reduce inside the mat-vec kernels, launch by launch
first = True
for li, L in enumerate(model.layers):
    be.mul_mat_vec_tp([L["wq"], L["wk"], L["wv"]], x0 if first else None, [model.q, model.kk, model.v], r, reduce_in=not first)
    be.mul_mat_vec_tp([L["wo"]], model.q, None, r, reduce_out=True)
    be.mul_mat_vec_tp([L["up"]], None, [model.a], r, reduce_in=True, gate=L["gate"], unary="silu")
    say(f"fused layer {li}: reduced h vs ref {nm(r.reduced_view(N_EMBD)[None, :], refs[li][0])}")
    be.mul_mat_vec_tp([L["down"]], model.a, None, r, reduce_out=True)
    be.mul_mat_vec_tp([model.layers[0]["wq"]], None, [model.q], r, reduce_in=True) if li == 1 else None
    if li == 1:
        say(f"fused layer {li}: reduced x2 vs ref {nm(r.reduced_view(N_EMBD)[None, :], refs[li][1])}")
    first = False
torch.cuda.synchronize(); dist.barrier()
# (3) the gate itself
err, ok = bench.tp_correctness_gate(be, torch, dist, model, rank, world, 1)
say("gate tg:", err, ok)
dist.destroy_process_group()
