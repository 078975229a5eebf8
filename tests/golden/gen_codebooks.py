"""Extract the IQ2_XXS codebook (iq2xxs_grid: 256 x 8 magnitudes, ksigns_iq2xs: 128 sign masks) by RUNNING the unmodified reference:
crafted blocks go through its own to_float (oracle/_ref) and the outputs are the table entries.  Nothing is copied from the reference
sources.  Run where /root/reference exists:  python tests/golden/gen_codebooks.py  -> tests/golden/iq2xxs_codebook.npz
block_iq2_xxs = {half d; u16 qs[32]} (ggml-common.h:439-442): per 32 weights two u32: [4 grid indices][4 x 7-bit sign index | 4-bit scale << 28];
dequantize_row_iq2_xxs (ggml-quants.c:3674-3698): y = d * (0.5 + scale) * 0.25 * grid[idx][j] * (+-1)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.oracle import GGML_TYPE, RefLib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def blocks(groups):
    """groups: list of (4 grid indices, 28-bit sign word) per 32 weights; d = 8.0 and scale 0 -> db = 1."""
    assert len(groups) % 8 == 0
    out = []
    for b in range(len(groups) // 8):
        blk = bytearray(np.float16(8.0).tobytes())
        for idx, signs in groups[8 * b: 8 * b + 8]:
            blk += bytes(idx) + int(signs).to_bytes(4, "little")
        assert len(blk) == 66
        out.append(bytes(blk))
    return np.frombuffer(b"".join(out), np.uint8)


def main():
    R = RefLib(); t = GGML_TYPE["IQ2_XXS"]
    wire = blocks([((i, i, i, i), 0) for i in range(256)])
    y = R.to_float(t, wire, 1, 256 * 32).reshape(256, 32)
    assert (y >= 0).all(), "ksigns[0] is expected to be 'all positive'"
    grid = y[:, :8].astype(np.uint8)
    assert np.array_equal(grid.astype(np.float32), y[:, :8]) and np.array_equal(y[:, :8], y[:, 8:16])
    wire = blocks([((0, 0, 0, 0), k) for k in range(128)])
    y = R.to_float(t, wire, 1, 128 * 32).reshape(128, 32)
    neg = (y[:, :8] < 0)
    assert np.array_equal(np.abs(y[:, :8]), np.broadcast_to(grid[0].astype(np.float32), (128, 8)))
    ksigns = (neg * (1 << np.arange(8))).sum(1).astype(np.uint8)
    # IQ2_XS {half d; u16 qs[32]; u8 scales[8]} (ggml-quants.c:3702-3725): qs = 9-bit index into iq2xs_grid (512 x 8) | 7-bit sign index << 9;
    # db = d * (0.5 + scale nibble) * 0.25 -> d = 8, scales = 0 gives db = 1
    t2 = GGML_TYPE["IQ2_XS"]
    blk = []
    for b in range(512 // 32):
        qs = np.arange(32 * b, 32 * b + 32, dtype=np.uint16)
        blk.append(np.float16(8.0).tobytes() + qs.tobytes() + bytes(8))
    y = R.to_float(t2, np.frombuffer(b"".join(blk), np.uint8), 1, 512 * 8).reshape(512, 8)
    grid_xs = y.astype(np.uint8); assert np.array_equal(grid_xs.astype(np.float32), y)
    # IQ3_XXS {half d; u8 qs[64]; u8 scales_and_signs[32]} (ggml-quants.c:3761-3789): qs = index into iq3xxs_grid (256 x 4);
    # db = d * (0.5 + scale) * 0.5 -> d = 4, scale 0 gives db = 1
    t3 = GGML_TYPE["IQ3_XXS"]
    blk = []
    for b in range(256 // 64):
        blk.append(np.float16(4.0).tobytes() + np.arange(64 * b, 64 * b + 64, dtype=np.uint8).tobytes() + bytes(32))
    y = R.to_float(t3, np.frombuffer(b"".join(blk), np.uint8), 1, 256 * 4).reshape(256, 4)
    grid_3xxs = y.astype(np.uint8); assert np.array_equal(grid_3xxs.astype(np.float32), y)
    # IQ2_S {half d; u8 qs[64] (32 index bytes + 32 sign bytes); u8 qh[8]; u8 scales[8]} (ggml-quants.c:3727-3757):
    # index_l = qs[l] | ((qh[ib32] >> 2l) & 3) << 8 into iq2s_grid (1024 x 8); d = 8, scales 0 -> db = 1
    t2s = GGML_TYPE["IQ2_S"]
    blk = []
    for b in range(1024 // 32):                       # 32 grid entries per block (8 groups x 4)
        idx = np.arange(32 * b, 32 * b + 32)
        qs = (idx & 255).astype(np.uint8).tobytes() + bytes(32)
        qh = bytes(int(sum(((idx[4 * g + l] >> 8) & 3) << (2 * l) for l in range(4))) for g in range(8))
        blk.append(np.float16(8.0).tobytes() + qs + qh + bytes(8))
    y = R.to_float(t2s, np.frombuffer(b"".join(blk), np.uint8), 1, 1024 * 8).reshape(1024, 8)
    grid_2s = y.astype(np.uint8); assert np.array_equal(grid_2s.astype(np.float32), y)
    # IQ3_S {half d; u8 qs[64]; u8 qh[8]; u8 signs[32]; u8 scales[4]} (ggml-quants.c:3793-3835): index = qs[i] | 9th bit from qh (bit i%8 of qh[i/8])
    # into iq3s_grid (512 x 4); db = d * (1 + 2 * scale) -> d = 1, scales 0 gives db = 1
    t3s = GGML_TYPE["IQ3_S"]
    blk = []
    for b in range(512 // 64):
        idx = np.arange(64 * b, 64 * b + 64)
        qs = (idx & 255).astype(np.uint8).tobytes()
        qh = bytes(int(sum(((idx[8 * g + i] >> 8) & 1) << i for i in range(8))) for g in range(8))
        blk.append(np.float16(1.0).tobytes() + qs + qh + bytes(32) + bytes(4))
    y = R.to_float(t3s, np.frombuffer(b"".join(blk), np.uint8), 1, 512 * 4).reshape(512, 4)
    grid_3s = y.astype(np.uint8); assert np.array_equal(grid_3s.astype(np.float32), y)
    # IQ1_S {half d; u8 qs[32]; u16 qh[8]} (ggml-quants.c:3836-3859): per 32 weights 4 groups, index_l = qs[l] | ((qh >> 3l) & 7) << 8 into
    # iq1s_grid (2048 x 8, values in {-1,0,1}); y = d * (2*((qh >> 12) & 7) + 1) * (grid + (qh & 0x8000 ? -delta : delta)); d = 1, scale 0, sign 0
    t1s = GGML_TYPE["IQ1_S"]
    blk = []
    for b in range(2048 // 32):
        idx = np.arange(32 * b, 32 * b + 32)
        qs = (idx & 255).astype(np.uint8).tobytes()
        qh = np.array([sum(((int(idx[4 * g + l]) >> 8) & 7) << (3 * l) for l in range(4)) for g in range(8)], np.uint16).tobytes()
        blk.append(np.float16(1.0).tobytes() + qs + qh)
    y = R.to_float(t1s, np.frombuffer(b"".join(blk), np.uint8), 1, 2048 * 8).reshape(2048, 8)
    delta = float(y.ravel()[0] - np.round(y.ravel()[0]))
    grid_1s = np.round(y - delta).astype(np.int8)
    assert np.array_equal(grid_1s.astype(np.float32) + np.float32(delta), y) and set(grid_1s.ravel().tolist()) <= {-1, 0, 1}, delta
    # IQ2_KL row = {half d; blocks {u16 scales_h; u8 scales_l[4]; u8 qs[64]; u8 qh[16]}} (iqk_quantize.cpp:2243-2275): 5-bit index
    # (nibble of qs | bit of qh << 4) into iq2kl_values (32 PAIRS of int8); 6-bit scales - 32: scale 33 -> dl = d = 1
    t2kl = GGML_TYPE["IQ2_KL"]
    qs = bytes(((j & 15) | ((j & 15) << 4)) for j in range(16)) + bytes(48)       # ib64 = 0: low nibbles -> entries 0..15, high nibbles -> 16..31 with qh bit 1
    qh = bytes([0b10] * 16)
    row = np.float16(1.0).tobytes() + np.uint16(0b1010).tobytes() + bytes([0x11, 0x11, 0, 0]) + qs + qh
    y = R.to_float(t2kl, np.frombuffer(row, np.uint8), 1, 256)[0]
    kl = np.empty((32, 2), np.int8)
    for j in range(16):
        kl[j] = y[2 * j: 2 * j + 2]; kl[16 + j] = y[2 * j + 32: 2 * j + 34]
    assert np.array_equal(kl[:16].astype(np.float32).ravel(), y[:32]) and np.array_equal(kl[16:].astype(np.float32).ravel(), y[32:64])
    np.savez_compressed(os.path.join(HERE, "iq2xxs_codebook.npz"), grid=grid, ksigns=ksigns, iq2xs_grid=grid_xs, iq3xxs_grid=grid_3xxs,
                        iq2s_grid=grid_2s, iq3s_grid=grid_3s, iq1s_grid=grid_1s, iq1s_delta=np.float32(delta), iq2kl_values=kl)
    print("iq1s", grid_1s.shape, "delta", delta, "iq2kl", kl.ravel().tolist())
    print("iq2s", grid_2s.shape, sorted(set(grid_2s.ravel().tolist())), "iq3s", grid_3s.shape, sorted(set(grid_3s.ravel().tolist())))
    print("grid", grid.shape, sorted(set(grid.ravel().tolist())), "ksigns", ksigns[:8].tolist(), "...",
          "iq2xs", grid_xs.shape, sorted(set(grid_xs.ravel().tolist())), "iq3xxs", grid_3xxs.shape, sorted(set(grid_3xxs.ravel().tolist())))


if __name__ == "__main__":
    main()
