#!/usr/bin/env python
"""One GPU: hunt for the NaN seen in the tg correctness gate on tensor-parallel shard shapes (down: K = 7168 / 3584)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ik_llama_cpp_b200 import backend as be

torch.cuda.set_device(0)
def stats(name, v):
    v = v.double()
    print(f"   {name:28s} shape {tuple(v.shape)} nan {int(torch.isnan(v).sum())} inf {int(torch.isinf(v).sum())} absmax {float(v[torch.isfinite(v)].abs().max()) if torch.isfinite(v).any() else float('nan'):.4g}", flush=True)

for tp in (2, 4):
    m = bench.Model(be, torch, 2, tp=tp, rank=0, collective=False)
    gen = torch.Generator(device="cuda"); gen.manual_seed(777)
    x = torch.randn(1, bench.N_EMBD, device="cuda", generator=gen)
    for li, L in enumerate(m.layers):
        print(f"tp={tp} layer {li}")
        for nm in ("wq", "wo", "up", "gate", "down"):
            w = L[nm]
            wire = be.get_tensor(w)
            d = np.frombuffer(wire.tobytes(), np.uint8).reshape(-1, 18)[:, :2].copy().view(np.float16).astype(np.float32)
            dq = be.dequantize_bf16(w)
            print(f"   {nm}: m={w.m} k={w.k} wire d: finite {bool(np.isfinite(d).all())} min {d.min():.3g} max {d.max():.3g}; dequant nan {int(torch.isnan(dq).sum())} inf {int(torch.isinf(dq).sum())} absmax {float(dq.float().abs().max()):.4g}", flush=True)
        q = be.mul_mat(L["wq"], x); stats("q = wq x", q)
        h = be.mul_mat(L["wo"], q); stats("h = wo q", h)
        a = be.fused_up_gate(L["up"], L["gate"], h, "silu"); stats("a = glu(up h, gate h)", a)
        y = be.mul_mat(L["down"], a); stats("y = down a (mat-vec)", y)
        a16 = a.repeat(16, 1).contiguous(); y16 = be.mul_mat(L["down"], a16); stats("y (GEMM, 16 copies of a)", y16)
        ref = (be.dequantize_bf16(L["down"]).double() @ a.double().T).T; stats("y (torch f64 on dequant)", ref)
        r = torch.randn(1, L["down"].k, device="cuda", generator=gen); stats("down . randn (mat-vec)", be.mul_mat(L["down"], r))
        a2 = a.clone(); stats("down . a.clone()", be.mul_mat(L["down"], a2))
        stats("layer-0 down . a", be.mul_mat(m.layers[0]["down"], a))
        # which blocks of a are unusual?
        ab = a.view(-1, 32).abs().amax(dim=1)
        print(f"   a: per-32 amax min {float(ab.min()):.3g} (zeros: {int((ab == 0).sum())}, < 1e-30: {int((ab < 1e-30).sum())}) max {float(ab.max()):.3g}", flush=True)
        x = torch.nan_to_num(y)
