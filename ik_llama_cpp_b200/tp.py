"""Tensor-parallel ("split mode graph") host logic for the quantized mat-mul path.

Mirrors the reference's sharding of MUL_MAT weights (src/llama-load-tensors.cpp:395-440 create_split,
:4647-4705 prepare_split_tensors, :5452-5477 attention/FFN policy; scatter in ggml_backend_cuda_split_buffer_set_tensor,
ggml/src/ggml-cuda.cu:1003-1192):
  * split_dim = 1 (rows of W: wq/wk/wv/ffn_up/ffn_gate/output): shard = contiguous row range, no exchange;
  * split_dim = 0 (columns/K of W: wo/ffn_down): shard = K range (multiple of the granularity, >= quant block) of EVERY row,
    partial outputs are summed across ranks (GGML_OP_REDUCE, ggml-cuda/reduce.cu:125);
  * per-row headers (row_meta_size: IQ4_KS, IQ2_BN, ...) are replicated into every K-shard.
Pure numpy: usable in CPU tests with the oracle as the compute stand-in, and by bench.py / the backend for real shards.
"""
from __future__ import annotations

import numpy as np

# (block elements, block bytes, row meta bytes) — ggml type traits (ggml/src/ggml.c:640-1460)
GEOM = {2: (32, 18, 0), 3: (32, 20, 0), 6: (32, 22, 0), 7: (32, 24, 0), 133: (32, 26, 0), 8: (32, 34, 0), 12: (256, 144, 0), 13: (256, 176, 0), 14: (256, 210, 0), 20: (32, 18, 0), 23: (256, 136, 0),
        135: (64, 16, 4), 139: (256, 144, 0), 140: (256, 176, 0), 144: (256, 136, 4),
        10: (256, 84, 0), 11: (256, 110, 0), 137: (256, 76, 0), 138: (256, 110, 0), 39: (32, 17, 0), 152: (256, 168, 4), 145: (256, 70, 2), 156: (256, 102, 2),
        # wire-layout types (ik_llama_cpp_b200/csrc/b200q_wire.cuh)
        16: (256, 66, 0), 17: (256, 74, 0), 18: (256, 98, 0), 22: (256, 82, 0), 21: (256, 110, 0), 19: (256, 50, 0), 29: (256, 56, 0), 141: (256, 212, 0), 146: (256, 128, 4),
        157: (256, 86, 2), 134: (64, 13, 2), 158: (256, 56, 4), 153: (256, 68, 4), 154: (256, 100, 4), 155: (256, 128, 4),
        219: (32, 6, 2), 229: (32, 7, 2), 337: (256, 76, 0), 338: (256, 110, 0), 339: (256, 144, 0), 340: (256, 176, 0), 344: (256, 136, 4), 352: (256, 168, 4)}
# rows interleaved on the wire (the _R4 repacks): a wire "row group" of 4 rows = {4 row headers}{blocks of 4 rows}
INTERLEAVE = {219: 4, 229: 4, 337: 4, 338: 4, 339: 4, 340: 4, 344: 4, 352: 4}


def create_split(nr: int, granularity: int, world: int) -> list[int]:
    """Even split of `nr` in units of `granularity` (reference create_split with uniform `splits` and equal memory use):
    chunks are handed out round(p*nchunk) per device, the remainder goes to the first devices."""
    if granularity < 0:
        return [nr] * world
    assert nr % granularity == 0, (nr, granularity)
    nchunk = nr // granularity
    base, rem = divmod(nchunk, world)
    return [(base + (1 if i < rem else 0)) * granularity for i in range(world)]


def row_size(ggml_type: int, k: int) -> int:
    qk, bs, meta = GEOM[ggml_type]
    assert k % qk == 0
    return meta + (k // qk) * bs


def shard_rows(wire: np.ndarray, ggml_type: int, m: int, k: int, world: int, rank: int, granularity: int = 1):
    """split_dim = 1: rows [r0, r1) of the wire tensor.  Returns (shard_bytes, m_shard)."""
    granularity = max(granularity, INTERLEAVE.get(ggml_type, 1)) if granularity > 0 else granularity
    sizes = create_split(m, granularity, world)
    r0 = sum(sizes[:rank])
    rs = row_size(ggml_type, k)
    w = np.ascontiguousarray(wire, np.uint8).reshape(m, rs)
    return np.ascontiguousarray(w[r0:r0 + sizes[rank]]).reshape(-1), sizes[rank]


def shard_cols(wire: np.ndarray, ggml_type: int, m: int, k: int, world: int, rank: int, granularity: int | None = None):
    """split_dim = 0: K range [k0, k1) of every row (granularity >= quant block), row header replicated.
    Returns (shard_bytes, k_shard, k0)."""
    qk, bs, meta = GEOM[ggml_type]
    g = max(qk, granularity or qk)
    assert g % qk == 0
    sizes = create_split(k, g, world)
    k0 = sum(sizes[:rank]); ks = sizes[rank]
    rs = row_size(ggml_type, k)
    il = INTERLEAVE.get(ggml_type, 1)                 # row groups: the K range of a group of `il` interleaved rows is contiguous on the wire
    assert m % il == 0
    w = np.ascontiguousarray(wire, np.uint8).reshape(m // il, il * rs)
    out = np.empty((m // il, il * (meta + (ks // qk) * bs)), np.uint8)
    out[:, :il * meta] = w[:, :il * meta]
    out[:, il * meta:] = w[:, il * (meta + (k0 // qk) * bs): il * (meta + ((k0 + ks) // qk) * bs)]
    return out.reshape(-1), ks, k0


def llama_layer_plan(n_embd: int, n_ff: int, n_head: int, n_head_kv: int, world: int, ggml_type: int):
    """Shard sizes of one Llama layer under -sm graph (src/llama-load-tensors.cpp:5452-5477, :5619-5633)."""
    head_dim = n_embd // n_head
    gqa = n_head // n_head_kv
    qk = GEOM[ggml_type][0]
    gran_kq = head_dim * gqa                       # wq rows per KV-head group
    gran_vo = max(head_dim * gqa, qk)              # wo columns
    q_rows = create_split(n_embd, gran_kq, world)
    kv_rows = [r // gqa for r in q_rows]
    o_cols = create_split(n_embd, gran_vo, world)
    ff = create_split(n_ff, max(qk, 1), world)     # ffn_up/gate rows == ffn_down columns
    return {"wq_rows": q_rows, "wkv_rows": kv_rows, "wo_cols": o_cols, "ffn_rows": ff, "ffn_down_cols": ff}
