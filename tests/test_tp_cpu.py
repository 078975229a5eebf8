"""CPU tests (gloo, world_size 2) of the tensor-parallel host logic: row / K sharding of wire tensors + all-reduce of the
row-parallel partials reproduce the unsharded result.  The oracle stands in for the GPU kernels (tests only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import random_wire
from ik_llama_cpp_b200 import tp
from oracle.oracle import GGML_TYPE, Oracle


def test_create_split_matches_reference_rules():
    assert tp.create_split(4096, 512, 8) == [512] * 8                      # Llama-3-8B wq at G=8
    assert tp.create_split(14336, 256, 8) == [1792] * 8                    # ffn_down K = 7 x 256 per GPU
    assert sum(tp.create_split(25600, 256, 8)) == 25600                    # Qwen3-32B: uneven but complete, multiples of 256
    assert all(s % 256 == 0 for s in tp.create_split(25600, 256, 8))
    assert tp.create_split(100, -1, 4) == [100] * 4                        # replicate
    plan = tp.llama_layer_plan(4096, 14336, 32, 8, 8, GGML_TYPE["IQ4_NL"])
    assert plan["wq_rows"] == [512] * 8 and plan["wkv_rows"] == [128] * 8 and plan["wo_cols"] == [512] * 8


def test_geometry_table_covers_every_supported_type():
    from conftest import ALL_TYPES
    O = Oracle()
    for name in ALL_TYPES:
        t = GGML_TYPE[name]
        assert t in tp.GEOM, name
        assert tp.row_size(t, 1024) == O.row_size(t, 1024), name


@pytest.mark.parametrize("name", ["IQ4_NL", "Q4_K", "IQ4_KS", "IQ2_BN", "Q2_K", "IQ3_K", "MXFP4", "IQ5_KS", "IQ3_KS"])
def test_shards_partition_the_tensor(name):
    t = GGML_TYPE[name]
    m, k, world = 12, 2048, 4
    wire = random_wire(name, m, k, np.random.default_rng(3))
    O = Oracle()
    full = O.dequantize(t, wire, m, k)
    rows = [tp.shard_rows(wire, t, m, k, world, r, granularity=1) for r in range(world)]
    assert sum(ms for _, ms in rows) == m
    got = np.concatenate([O.dequantize(t, b, ms, k) for b, ms in rows if ms])
    assert np.array_equal(got, full)
    cols = [tp.shard_cols(wire, t, m, k, world, r, granularity=256) for r in range(world)]
    got = np.concatenate([O.dequantize(t, b, m, ks) for b, ks, _ in cols], axis=1)
    assert np.array_equal(got, full), "K shards (with replicated row header) must dequantise to the column slices"


def _worker(rank, world, port, name, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t = GGML_TYPE[name]
        m, k, n = 24, 1024, 2
        rng = np.random.default_rng(5)                        # same tensors on every rank
        w_up = random_wire(name, m, k, rng)                   # column-parallel (row split): no exchange
        kd = 768
        w_down = random_wire(name, 16, kd, rng)               # row-parallel (K split): partial sums need the all-reduce
        x = rng.standard_normal((n, k)).astype(np.float32)
        O = Oracle()
        # stage 1: row-split mat-mul, every rank owns a slice of the output
        shard, ms = tp.shard_rows(w_up, t, m, k, world, rank)
        y_loc = O.mul_mat_exact(t, shard, x, ms)
        # stage 2: K-split mat-mul on a (replicated) activation + all-reduce of the partials (GGML_OP_REDUCE)
        a = rng.standard_normal((n, kd)).astype(np.float32)
        shard2, ks, k0 = tp.shard_cols(w_down, t, 16, kd, world, rank, granularity=256)
        part = torch.from_numpy(O.mul_mat_exact(t, shard2, a[:, k0:k0 + ks], 16))
        dist.all_reduce(part)
        ref = O.mul_mat_exact(t, w_down, a, 16)
        err = float(np.abs(part.numpy() - ref).max() / max(np.abs(ref).max(), 1e-30))
        ref1 = O.mul_mat_exact(t, w_up, x, m)
        r0 = sum(tp.create_split(m, 1, world)[:rank])
        err1 = float(np.abs(y_loc - ref1[:, r0:r0 + ms]).max())
        out_q.put((rank, err, err1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["IQ4_NL", "IQ4_KS"])
def test_row_parallel_allreduce_world2(name):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, err1 in res:
        assert err <= 1e-5, f"rank {rank}: all-reduced K-split partials differ from the unsharded result ({err})"
        assert err1 == 0.0, f"rank {rank}: row shard result must equal the corresponding slice bit-for-bit"
