"""CPU tests of the PRODUCT's format code: ik_llama_cpp_b200/csrc/b200q_types.cuh is __host__ __device__, so the
very same repack / PRMT-LUT / decode arithmetic the CUDA kernels execute is compiled with g++ (tests/host_emul/emul.cpp)
and checked against the oracle: wire->planes->wire bijection, bit-exact dequantisation, mat-vec arithmetic."""
import ctypes
import os
import subprocess
from ctypes import c_int, c_long, c_void_p

import numpy as np
import pytest

from conftest import ALL_TYPES, PLANE_TYPES, WIRE_TYPES, ROOT, load_golden, random_wire
from oracle.oracle import GGML_TYPE, _p, nmse


@pytest.fixture(scope="session")
def emul():
    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    so = os.path.join(ROOT, "tests", "host_emul", "libemul.so")
    hdrs = [os.path.join(ROOT, "ik_llama_cpp_b200", "csrc", h) for h in ("b200q_types.cuh", "b200q_wire.cuh", "b200q_codebooks.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    E = ctypes.CDLL(so)
    E.emul_layout_bytes.restype = c_long
    E.emul_layout_bytes.argtypes = [c_int, c_long, c_long]
    E.emul_wire_row_size.restype = c_long
    E.emul_wire_row_size.argtypes = [c_int, c_long]
    E.emul_repack.argtypes = [c_int, c_void_p, c_void_p, c_long, c_long, c_int]
    E.emul_dequant.argtypes = [c_int, c_void_p, c_long, c_long, c_void_p]
    E.emul_mul_mat_vec.argtypes = [c_int, c_void_p, c_long, c_long, c_void_p, c_void_p, c_void_p, c_long, c_void_p]
    E.emul_wire_dequant.argtypes = [c_int, c_void_p, c_long, c_long, c_void_p]
    return E


def _emul_all(E, oracle, t, wire, m, k, x):
    nb = E.emul_layout_bytes(t, m, k)
    planes = np.zeros(nb, np.uint8)
    assert E.emul_repack(t, _p(wire), _p(planes), m, k, 0) == 0
    back = np.full_like(wire, 0x5A)
    assert E.emul_repack(t, _p(back), _p(planes), m, k, 1) == 0
    deq = np.empty((m, k), np.float32)
    assert E.emul_dequant(t, _p(planes), m, k, _p(deq)) == 0
    n = x.shape[0]
    q, d = oracle.quantize_q8_1_b200(x)
    xd = np.ascontiguousarray(d.astype(np.float32))
    q16 = q.reshape(n, k // 32, 2, 16).astype(np.int32).sum(-1)
    xis = np.ascontiguousarray(((q16[..., 0] & 0xFFFF) | (q16[..., 1] << 16)).astype(np.int32))
    y = np.empty((n, m), np.float32)
    assert E.emul_mul_mat_vec(t, _p(planes), m, k, _p(q), _p(xd), _p(xis), n, _p(y)) == 0
    return nb, back, deq, y


@pytest.mark.parametrize("name", PLANE_TYPES)
def test_emulated_kernel_arithmetic_on_golden(emul, oracle, name):
    g = load_golden(name)
    t, m, k = int(g["ggml_type"]), int(g["m"]), int(g["k"])
    wire, x = g["wire"], g["x"]
    assert emul.emul_wire_row_size(t, k) == int(g["row_size"])
    nb, back, deq, y = _emul_all(emul, oracle, t, wire, m, k, x)
    assert nb >= wire.size and nb <= wire.size + 5 * 256            # same bytes, only 256-B plane alignment on top
    assert np.array_equal(back, wire), "planes -> wire must restore the GGUF bytes bit-for-bit"
    if name in ("IQ4_KS", "IQ5_KS"):
        np.testing.assert_allclose(deq, g["dequant_ref"], rtol=2e-7)
    else:
        assert np.array_equal(deq, g["dequant_ref"]), "canonical decode must equal the reference to_float bit-for-bit"
    yq = oracle.mul_mat_q8_1(t, wire, x, m, variant="b200")
    rms = np.sqrt((yq.astype(np.float64) ** 2).mean())
    assert np.abs(y - yq).max() <= 2e-5 * rms, "mat-vec arithmetic differs from the restated reference MMVQ beyond f32 summation order"
    assert nmse(y, oracle.mul_mat_exact(t, wire, x, m)) <= 5e-4


@pytest.mark.parametrize("name", PLANE_TYPES)
def test_emulated_random_bit_patterns(emul, oracle, name):
    """Every payload bit pattern is a valid block: fuzz the codecs with random bytes (catches index/LUT corner cases)."""
    t = GGML_TYPE[name]
    rng = np.random.default_rng(4242 + t)
    m, k = 6, 1024
    wire = random_wire(name, m, k, rng)
    x = rng.standard_normal((2, k)).astype(np.float32)
    nb, back, deq, y = _emul_all(emul, oracle, t, wire, m, k, x)
    assert np.array_equal(back, wire)
    ref = oracle.dequantize(t, wire, m, k)
    if name in ("IQ4_KS", "IQ5_KS", "IQ2_BN"):   # dl*q - ml vs dl*(q - c): one rounding apart for out-of-codebook bit patterns
        np.testing.assert_allclose(deq, ref, rtol=3e-7, atol=1e-12)
    else:
        assert np.array_equal(deq, ref)
    yq = oracle.mul_mat_q8_1(t, wire, x, m, variant="b200")
    rms = np.sqrt((yq.astype(np.float64) ** 2).mean())
    assert np.abs(y - yq).max() <= 2e-5 * rms


@pytest.mark.parametrize("name", PLANE_TYPES)
def test_emulated_edge_shapes(emul, oracle, name):
    """Edge shapes of the codecs: a single block per row, an odd block count, one row, the row-scale types with every K — the
    bijection and the canonical decode must hold for each (the reference's quantiser is not needed: any payload is a valid block)."""
    from conftest import _QK
    t = GGML_TYPE[name]
    qk = _QK.get(name, 256)
    rng = np.random.default_rng(977 + t)
    for m, nblk in ((1, 1), (3, 1), (2, 3), (5, 7)):
        k = nblk * max(qk, 32)
        if k % qk:
            continue
        wire = random_wire(name, m, k, rng)
        x = rng.standard_normal((1, k)).astype(np.float32)
        nb, back, deq, y = _emul_all(emul, oracle, t, wire, m, k, x)
        assert np.array_equal(back, wire), (name, m, k)
        ref = oracle.dequantize(t, wire, m, k)
        if name in ("IQ4_KS", "IQ5_KS", "IQ2_BN"):
            np.testing.assert_allclose(deq, ref, rtol=3e-7, atol=1e-12)
        else:
            assert np.array_equal(deq, ref), (name, m, k)
        yq = oracle.mul_mat_q8_1(t, wire, x, m, variant="b200")
        rms = max(float(np.sqrt((yq.astype(np.float64) ** 2).mean())), 1e-30)
        assert np.abs(y - yq).max() <= 2e-5 * rms, (name, m, k)


@pytest.mark.parametrize("name", WIRE_TYPES)
def test_wire_decode32_is_bit_exact_on_golden(emul, oracle, name):
    """Wire-layout types: the product's b200q_wire_decode32 (the function the CUDA kernels call) == the reference's own to_float output
    (golden, generated from oracle/_ref) bit for bit, and == the oracle restatement."""
    g = load_golden(name)
    t, m, k = int(g["ggml_type"]), int(g["m"]), int(g["k"])
    assert emul.emul_wire_row_size(t, k) == int(g["row_size"])
    assert emul.emul_layout_bytes(t, m, k) >= g["wire"].size
    deq = np.empty((m, k), np.float32)
    assert emul.emul_wire_dequant(t, _p(np.ascontiguousarray(g["wire"])), m, k, _p(deq)) == 0
    assert np.array_equal(deq, oracle.dequantize(t, g["wire"], m, k))          # product decode == oracle restatement, bit for bit
    if name == "IQ6_K":     # the reference build contracts the float cubic of its to_float into FMAs (tests/test_oracle.py FMA_DEPENDENT)
        np.testing.assert_allclose(deq, g["dequant_ref"], rtol=3e-6, atol=1e-5 * float(np.abs(g["dequant_ref"]).max()))
    else:
        assert np.array_equal(deq, g["dequant_ref"]), f"{name}: max |diff| {np.abs(deq - g['dequant_ref']).max()}"
