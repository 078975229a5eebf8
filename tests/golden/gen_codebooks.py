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
    np.savez_compressed(os.path.join(HERE, "iq2xxs_codebook.npz"), grid=grid, ksigns=ksigns)
    print("grid", grid.shape, sorted(set(grid.ravel().tolist())), "ksigns", ksigns[:8].tolist(), "...")


if __name__ == "__main__":
    main()
