"""Warm per-op timing of one Llama-3-8B prefill layer (pp512) through the public operator mirror: which GEMM costs what.
Cycles over NL distinct layers so that weights are not L2-resident.  Usage: python scripts/pp_breakdown.py [n_tokens]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ik_llama_cpp_b200 import backend as be

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NL = 6
m = bench.Model(be, torch, NL)
m.alloc(n)
m.x.normal_()
E, F = bench.N_EMBD, bench.N_FF

def timeit(name, fn, flops=None, reps=5):
    for L in m.layers: fn(L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for L in m.layers: fn(L)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (reps * NL)
    tf = f"  {flops / us / 1e6:7.1f} TFLOP/s" if flops else ""
    print(f"{name:44s} {us:8.1f} us{tf}", flush=True)
    return us

g = lambda M, K: 2.0 * M * K * n
be.convert_activations(m.x, m.xb); be.convert_activations(m.q, m.qb); be.convert_activations(m.h, m.hb); be.convert_activations(m.a, m.ab)
tot = 0
tot += timeit("convert x (E)", lambda L: be.convert_activations(m.x, m.xb))
timeit("wq single", lambda L: be.mul_mat(L["wq"], m.x, out=m.q, x_bf16=m.xb), g(E, E))
timeit("wk single", lambda L: be.mul_mat(L["wk"], m.x, out=m.kk, x_bf16=m.xb), g(1024, E))
tot += timeit("qkv multi", lambda L: be.mul_mat_multi([L["wq"], L["wk"], L["wv"]], m.x, [m.q, m.kk, m.v], x_bf16=m.xb), g(E + 2048, E))
tot += timeit("convert q", lambda L: be.convert_activations(m.q, m.qb))
tot += timeit("wo", lambda L: be.mul_mat(L["wo"], m.q, out=m.h, x_bf16=m.qb), g(E, E))
tot += timeit("convert h", lambda L: be.convert_activations(m.h, m.hb))
timeit("up single", lambda L: be.mul_mat(L["up"], m.h, out=m.u, x_bf16=m.hb), g(F, E))
timeit("up+gate multi (plain)", lambda L: be.mul_mat_multi([L["up"], L["gate"]], m.h, [m.u, m.g], x_bf16=m.hb), 2 * g(F, E))
tot += timeit("fused_up_gate (epilogue, +bf16)", lambda L: be.fused_up_gate(L["up"], L["gate"], m.h, "silu", out=m.a, x_bf16=m.hb, out_bf16=m.ab), 2 * g(F, E))
timeit("fused_up_gate (epilogue, no bf16)", lambda L: be.fused_up_gate(L["up"], L["gate"], m.h, "silu", out=m.a, x_bf16=m.hb), 2 * g(F, E))
timeit("convert a (F)", lambda L: be.convert_activations(m.a, m.ab))
tot += timeit("down", lambda L: be.mul_mat(L["down"], m.a, out=m.x2, x_bf16=m.ab), g(E, F))
print(f"layer total (ops on the bench path): {tot:.1f} us -> {n / (tot * 32 * 1e-6):.0f} tok/s for 32 layers (head excluded)")
