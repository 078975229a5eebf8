#!/usr/bin/env python
"""One GPU: the n = 1 path on the shapes a tensor-parallel shard sees (tp = 2, 4, 8), without collectives: where does a NaN / mismatch appear?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ik_llama_cpp_b200 import backend as be

torch.cuda.set_device(0)
for tp in (2, 4, 8):
    m = bench.Model(be, torch, 2, tp=tp, rank=0, collective=False)
    gen = torch.Generator(device="cuda"); gen.manual_seed(777)
    x = torch.randn(1, bench.N_EMBD, device="cuda", generator=gen)
    xn = torch.randn(16, bench.N_EMBD, device="cuda", generator=gen); xn[0] = x[0]
    for li, L in enumerate(m.layers):
        q = be.mul_mat(L["wq"], x); qn = be.mul_mat(L["wq"], xn)
        h = be.mul_mat(L["wo"], q); hn = be.mul_mat(L["wo"], qn)
        a = be.fused_up_gate(L["up"], L["gate"], h, "silu"); an = be.fused_up_gate(L["up"], L["gate"], hn, "silu")
        q8 = be.Q8Scratch(bench.N_FF // tp)
        a2 = be.fused_up_gate(L["up"], L["gate"], h, "silu", q8_out=q8)
        y = be.mul_mat(L["down"], a); yn = be.mul_mat(L["down"], an)
        y2 = be.mul_mat(L["down"], a2, q8_in=q8)
        torch.cuda.synchronize()
        def st(name, v, ref):
            v, ref = v.double(), ref.double()
            print(f"tp={tp} layer {li} {name:10s} shape {tuple(v.shape)} nan {int(torch.isnan(v).sum())} inf {int(torch.isinf(v).sum())} "
                  f"nmse vs GEMM row 0 {float(((v - ref) ** 2).sum() / (ref ** 2).sum()):.3g}")
        st("q", q, qn[:1]); st("h", h, hn[:1]); st("a", a, an[:1]); st("a(q8out)", a2, an[:1]); st("y", y, yn[:1]); st("y(q8in)", y2, yn[:1])
        x = y; xn = yn
