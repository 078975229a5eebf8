import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
ORACLE_ONLY_TYPES = ["IQ2_XXS", "IQ2_XS", "IQ3_XXS", "IQ2_S", "IQ3_S", "IQ6_K", "IQ1_BN", "IQ4_KSS"]      # oracle pinned against the reference, device kernel not built yet
ALL_TYPES = ["Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q6_0", "Q8_0", "Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K", "IQ4_NL", "IQ4_XS", "IQ2_K", "IQ3_K", "IQ4_K", "IQ5_K", "IQ4_KS", "IQ5_KS", "IQ2_KS", "IQ3_KS", "MXFP4", "IQ2_BN"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    """The unmodified reference CPU library, if oracle/_ref was built (needs /root/reference at build time)."""
    from oracle.oracle import RefLib
    if RefLib.find(prefer_native=False) is None:
        pytest.skip("oracle/_ref not built")
    return RefLib()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    return {k: z[k] for k in z.files}


def make_wire(oracle_or_ref, name, m, k, seed, reflib=None):
    """Wire bytes for a random tensor.  With the reference library: real ggml_quantize_chunk output.
    Without it (GPU box may lack it): random but VALID wire blocks (every bit pattern of the payload is a valid
    encoding; scales are drawn as sane fp16/f32 values)."""
    from oracle.oracle import GGML_TYPE
    t = GGML_TYPE[name]
    rng = np.random.default_rng(seed)
    if reflib is not None:
        w = (rng.standard_normal((m, k)) * 0.02).astype(np.float32)
        if name == "IQ2_BN":
            w = (rng.integers(-1, 2, (m, k)) * 0.043).astype(np.float32)
        return reflib.quantize(t, w)
    return random_wire(name, m, k, rng)


# byte offsets of fp16 scale fields inside one wire block, per type: (block_bytes, [offsets of half fields], row_meta)
_GEOM = {
    "Q4_0": (18, [0], 0), "Q4_1": (20, [0, 2], 0), "Q5_0": (22, [0], 0), "Q5_1": (24, [0, 2], 0), "Q6_0": (26, [0], 0), "Q8_0": (34, [0], 0), "IQ4_NL": (18, [0], 0), "Q4_K": (144, [0, 2], 0), "Q5_K": (176, [0, 2], 0),
    "IQ2_K": (76, [0], 0), "IQ3_K": (110, [0], 0), "Q2_K": (84, [80, 82], 0), "Q3_K": (110, [108], 0), "Q6_K": (210, [208], 0), "IQ4_XS": (136, [0], 0), "IQ4_K": (144, [0], 0), "IQ5_K": (176, [0], 0), "IQ4_KS": (136, [], 4), "IQ5_KS": (168, [], 4), "IQ2_KS": (70, [], 2), "IQ3_KS": (102, [], 2), "MXFP4": (17, [], 0), "IQ2_BN": (16, [], 4),
}
_QK = {"Q4_0": 32, "Q4_1": 32, "Q5_0": 32, "Q5_1": 32, "Q6_0": 32, "Q8_0": 32, "IQ4_NL": 32, "MXFP4": 32, "IQ2_BN": 64}


def random_wire(name, m, k, rng):
    bs, halfs, meta = _GEOM[name]
    qk = _QK.get(name, 256)
    nb = k // qk
    rows = np.empty((m, meta + nb * bs), np.uint8)
    blocks = rng.integers(0, 256, (m, nb, bs), dtype=np.uint8)
    for off in halfs:
        sc = (rng.uniform(0.5, 2.0, (m, nb)) * 1e-3).astype(np.float16)
        if off in (2, 82):   # dmin of Q4_K/Q5_K/Q2_K: keep it small
            sc = (rng.uniform(0.0, 1.0, (m, nb)) * 1e-4).astype(np.float16)
        blocks[:, :, off:off + 2] = sc.view(np.uint8).reshape(m, nb, 2)
    if name == "MXFP4":   # E8M0 block exponent: keep the scale in a sane range (2^-18 .. 2^-4), every other bit pattern is payload
        blocks[:, :, 0] = rng.integers(110, 125, (m, nb), dtype=np.uint8)
    rows[:, meta:] = blocks.reshape(m, nb * bs)
    if meta == 2:       # IQ2_KS / IQ3_KS: ggml_half row scale
        rs = (rng.uniform(0.5, 2.0, m) * 1e-3).astype(np.float16)
        rows[:, :2] = rs.view(np.uint8).reshape(m, 2)
    if meta == 4:
        rs = (rng.uniform(0.5, 2.0, m) * 1e-3).astype(np.float32)
        rows[:, :4] = rs.view(np.uint8).reshape(m, 4)
    return rows.reshape(-1)
