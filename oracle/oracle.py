"""oracle/oracle.py — TEST INFRASTRUCTURE: ctypes front-end for the CPU oracle and the reference library.

* ``Oracle``   -> oracle/_build/liboracle.so, compiled from oracle/oracle_quants.c (our CPU restatement).
* ``RefLib``   -> oracle/_ref/libggml_ref_<variant>.so, the UNMODIFIED reference CPU ggml built by
                  oracle/Makefile.ref (+ oracle/ref_shim.c).  Exists wherever it was built in a container
                  that has /root/reference; it travels to the GPU box as a prebuilt file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint16, c_void_p, POINTER

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REFDIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

# ggml_type ids (reference ggml/include/ggml.h:391-492)
GGML_TYPE = {
    "F32": 0, "F16": 1, "Q4_0": 2, "Q4_1": 3, "Q5_0": 6, "Q5_1": 7, "Q8_0": 8, "Q2_K": 10, "Q3_K": 11,
    "Q4_K": 12, "Q5_K": 13, "Q6_K": 14, "IQ2_XXS": 16, "IQ2_XS": 17, "IQ3_XXS": 18, "IQ1_S": 19,
    "IQ4_NL": 20, "IQ3_S": 21, "IQ2_S": 22, "IQ4_XS": 23, "IQ1_M": 29, "BF16": 30, "MXFP4": 39,
    "Q6_0": 133, "IQ1_BN": 134, "IQ2_BN": 135, "IQ2_K": 137, "IQ3_K": 138, "IQ4_K": 139, "IQ5_K": 140,
    "IQ6_K": 141, "IQ4_KS": 144, "IQ2_KS": 145, "IQ4_KSS": 146, "IQ5_KS": 152, "IQ2_KT": 153,
    "IQ3_KT": 154, "IQ4_KT": 155, "IQ3_KS": 156, "IQ2_KL": 157, "IQ1_KT": 158,
    "IQ1_S_R4": 219, "IQ1_M_R4": 229, "IQ2_K_R4": 337, "IQ3_K_R4": 338, "IQ4_K_R4": 339, "IQ5_K_R4": 340, "IQ4_KS_R4": 344, "IQ5_KS_R4": 352,
}
ROWS_INTERLEAVED = {n: 4 for n in ("IQ1_S_R4", "IQ1_M_R4", "IQ2_K_R4", "IQ3_K_R4", "IQ4_K_R4", "IQ5_K_R4", "IQ4_KS_R4", "IQ5_KS_R4")}
TYPE_NAME = {v: k for k, v in GGML_TYPE.items()}


def build_oracle(force: bool = False) -> str:
    """Compile oracle_quants.c -> _build/liboracle.so (plain gcc, no reference needed)."""
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "liboracle.so")
    src = os.path.join(HERE, "oracle_quants.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-fvisibility=hidden",
                               "-ffp-contract=off", "-o", so, src, "-lm"])
    return so


def build_ref(variant: str = "avx2", jobs: int = 8) -> str | None:
    """Build oracle/_ref from the reference sources if they are present (this container only)."""
    so = os.path.join(REFDIR, f"libggml_ref_{variant}.so")
    if not os.path.isdir(REFERENCE_ROOT):
        return so if os.path.exists(so) else None
    args = ["make", "-f", os.path.join(HERE, "Makefile.ref"), f"-j{jobs}", f"VARIANT={variant}"]
    if variant == "native":
        args.append("ARCHFLAGS=-march=native")
    if variant == "avx512":       # the reference's AVX-512 / VNNI kernels with a PORTABLE flag set (Ice Lake server and later, Zen 4): safe on the GPU pool's hosts
        args.append("ARCHFLAGS=-march=icelake-server -mtune=generic")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return so


def _p(a: np.ndarray, ty=c_void_p):
    return a.ctypes.data_as(ty)


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.oracle_row_size.restype = c_int64
        L.oracle_row_size.argtypes = [c_int, c_int64]
        L.oracle_type_supported.argtypes = [c_int]
        L.oracle_dequantize_row.argtypes = [c_int, c_void_p, c_void_p, c_int64]
        L.oracle_quantize_q8_1.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]
        L.oracle_mul_mat_exact.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64]
        L.oracle_mul_mat_q8_1.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64]
        L.oracle_mul_mat_q8_1_b200.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64]
        L.oracle_quantize_q8_1_b200.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p]
        L.oracle_h2f.restype = c_float
        L.oracle_h2f.argtypes = [c_uint16]
        L.oracle_f2h.restype = c_uint16
        L.oracle_f2h.argtypes = [c_float]
        # codebook fixtures extracted from the running reference (tests/golden/gen_codebooks.py); kept alive for the library's lifetime
        cb = os.path.join(os.path.dirname(HERE), "tests", "golden", "iq2xxs_codebook.npz")
        if os.path.exists(cb):
            z = np.load(cb)
            self._iq2xxs = (np.ascontiguousarray(z["grid"], np.uint8), np.ascontiguousarray(z["ksigns"], np.uint8))
            L.oracle_set_iq2xxs_codebook.argtypes = [c_void_p, c_void_p]
            L.oracle_set_iq2xxs_codebook(_p(self._iq2xxs[0]), _p(self._iq2xxs[1]))
            L.oracle_set_grid.argtypes = [c_int, c_void_p]
            self._grids = {}
            for tid, key in ((17, "iq2xs_grid"), (18, "iq3xxs_grid"), (22, "iq2s_grid"), (21, "iq3s_grid"), (19, "iq1s_grid"), (157, "iq2kl_values")):
                if key in z.files:
                    self._grids[tid] = np.ascontiguousarray(z[key]).view(np.uint8)
                    L.oracle_set_grid(tid, _p(self._grids[tid]))
        L.oracle_dequantize_matrix.argtypes = [c_int, c_void_p, c_void_p, c_int64, c_int64]
        L.oracle_rows_interleaved.argtypes = [c_int]

    def supported(self, t: int) -> bool:
        return bool(self.lib.oracle_type_supported(t))

    def row_size(self, t: int, k: int) -> int:
        r = self.lib.oracle_row_size(t, k)
        if r < 0:
            raise ValueError(f"oracle: type {t} / k {k} unsupported")
        return r

    def dequantize(self, t: int, wire: np.ndarray, m: int, k: int) -> np.ndarray:
        rs = self.row_size(t, k)
        wire = np.ascontiguousarray(wire, dtype=np.uint8).reshape(m, rs)
        out = np.empty((m, k), np.float32)
        assert self.lib.oracle_dequantize_matrix(t, _p(wire), _p(out), m, k) == 0      # (handles the 4-row groups of the _R4 repacks)
        return out

    def rows_interleaved(self, t: int) -> int:
        return int(self.lib.oracle_rows_interleaved(t))

    def quantize_q8_1(self, x: np.ndarray):
        x = np.ascontiguousarray(x, np.float32)
        n, k = x.shape
        q = np.empty((n, k), np.int8)
        d = np.empty((n, k // 32), np.uint16)
        s = np.empty((n, k // 32), np.uint16)
        assert self.lib.oracle_quantize_q8_1(_p(x), n, k, _p(q), _p(d), _p(s)) == 0
        return q, d.view(np.float16), s.view(np.float16)

    def mul_mat_exact(self, t: int, wire: np.ndarray, x: np.ndarray, m: int) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        n, k = x.shape
        wire = np.ascontiguousarray(wire, np.uint8)
        assert wire.size == m * self.row_size(t, k)
        out = np.empty((n, m), np.float32)
        assert self.lib.oracle_mul_mat_exact(t, _p(wire), _p(x), _p(out), m, k, n) == 0
        return out

    def mul_mat_q8_1(self, t: int, wire: np.ndarray, x: np.ndarray, m: int, variant: str = "reference") -> np.ndarray:
        """dequant(W) . dequant_q8_1(x) in f64.  variant="reference": quantize_q8_1 exactly as ggml-cuda/quantize.cu;
        variant="b200": the product's quantiser (one division per block, round-half-even)."""
        x = np.ascontiguousarray(x, np.float32)
        n, k = x.shape
        wire = np.ascontiguousarray(wire, np.uint8)
        assert wire.size == m * self.row_size(t, k)
        out = np.empty((n, m), np.float32)
        fn = self.lib.oracle_mul_mat_q8_1 if variant == "reference" else self.lib.oracle_mul_mat_q8_1_b200
        assert fn(t, _p(wire), _p(x), _p(out), m, k, n) == 0
        return out

    def quantize_q8_1_b200(self, x: np.ndarray):
        x = np.ascontiguousarray(x, np.float32)
        n, k = x.shape
        q = np.empty((n, k), np.int8)
        d = np.empty((n, k // 32), np.uint16)
        assert self.lib.oracle_quantize_q8_1_b200(_p(x), n, k, _p(q), _p(d)) == 0
        return q, d.view(np.float16)


def _cpu_flags() -> set[str]:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


class RefLib:
    """The reference's own CPU ggml (IQK path), through oracle/ref_shim.c."""

    @staticmethod
    def find(prefer_native: bool = True) -> str | None:
        cands = []
        # (/proc/cpuinfo spells it avx512_vnni; round 1 looked for "avx512vnni" and therefore always fell back to the AVX2 build)
        if prefer_native and {"avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512_vnni", "avx512vbmi", "avx512_vbmi2", "avx512_bitalg", "avx512_vpopcntdq"} <= _cpu_flags():
            cands.append(os.path.join(REFDIR, "libggml_ref_avx512.so"))
        cands.append(os.path.join(REFDIR, "libggml_ref_avx2.so"))
        for c in cands:
            if os.path.exists(c):
                return c
        return None

    def __init__(self, path: str | None = None):
        path = path or self.find(prefer_native=False)
        if path is None:
            raise FileNotFoundError("oracle/_ref not built (needs /root/reference: `make -f oracle/Makefile.ref`)")
        self.path = path
        self.lib = L = ctypes.CDLL(path)
        L.refshim_blck_size.restype = c_int64
        L.refshim_blck_size.argtypes = [c_int]
        L.refshim_type_size.restype = c_size_t
        L.refshim_type_size.argtypes = [c_int]
        L.refshim_row_size.restype = c_size_t
        L.refshim_row_size.argtypes = [c_int, c_int64]
        L.refshim_type_name.restype = c_char_p
        L.refshim_type_name.argtypes = [c_int]
        L.refshim_row_meta_size.restype = c_int64
        L.refshim_row_meta_size.argtypes = [c_int]
        L.refshim_quantize.restype = c_size_t
        L.refshim_quantize.argtypes = [c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p]
        L.refshim_to_float.argtypes = [c_int, c_void_p, c_void_p, c_int64]
        L.refshim_mul_mat.restype = c_double
        L.refshim_mul_mat.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int]
        L.refshim_chain_new.restype = c_void_p
        L.refshim_chain_new.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int]
        L.refshim_chain_set.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
        L.refshim_chain_get.argtypes = [c_void_p, c_int, c_void_p]
        L.refshim_chain_run.restype = c_double
        L.refshim_chain_run.argtypes = [c_void_p]
        L.refshim_chain_free.argtypes = [c_void_p]
        if hasattr(L, "refshim_fused_up_gate"):
            L.refshim_fused_up_gate.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_float, c_int]
            L.refshim_unary_op_id.argtypes = [c_char_p]

    def row_size(self, t: int, k: int) -> int:
        return int(self.lib.refshim_row_size(t, k))

    def row_meta_size(self, t: int) -> int:
        return int(self.lib.refshim_row_meta_size(t))

    def blck_size(self, t: int) -> int:
        return int(self.lib.refshim_blck_size(t))

    def type_name(self, t: int) -> str:
        return self.lib.refshim_type_name(t).decode()

    def quantize(self, t: int, w: np.ndarray) -> np.ndarray:
        w = np.ascontiguousarray(w, np.float32)
        m, k = w.shape
        out = np.zeros(m * self.row_size(t, k), np.uint8)
        n = self.lib.refshim_quantize(t, _p(w), _p(out), m, k, None)
        assert n == out.size, (n, out.size)
        return out

    # to_float of the ternary types takes a BLOCK pointer and ignores the row scale (SURVEY.md §8c pitfall 1,
    # iqk_quantize.cpp:375,418); every other row-meta type (IQ4_KS, ...) takes the ROW start and applies it itself.
    TO_FLOAT_IGNORES_ROW_SCALE = (134, 135)

    def to_float(self, t: int, wire: np.ndarray, m: int, k: int, apply_row_scale: bool = True) -> np.ndarray:
        """Reference to_float per row (row scale applied by hand for IQ1_BN/IQ2_BN when apply_row_scale)."""
        rs = self.row_size(t, k)
        meta = self.row_meta_size(t)
        wire = np.ascontiguousarray(wire, np.uint8).reshape(m, rs)
        out = np.empty((m, k), np.float32)
        if TYPE_NAME.get(t) in ROWS_INTERLEAVED:       # to_float of an _R4 type takes a group of 4 rows and n = 4 * k
            assert m % 4 == 0
            for g in range(m // 4):
                grp = np.ascontiguousarray(wire[4 * g: 4 * g + 4]).reshape(-1)
                assert self.lib.refshim_to_float(t, _p(grp), _p(out[4 * g: 4 * g + 4]), 4 * k) == 0
            return out
        for i in range(m):
            skip = meta if t in self.TO_FLOAT_IGNORES_ROW_SCALE else 0
            row = np.ascontiguousarray(wire[i, skip:])
            assert self.lib.refshim_to_float(t, _p(row), _p(out[i]), k) == 0
            if skip and apply_row_scale:
                scale = np.frombuffer(wire[i, :4].tobytes(), np.float32)[0] if meta == 4 else np.frombuffer(wire[i, :2].tobytes(), np.float16)[0].astype(np.float32)
                out[i] *= scale
        return out

    def mul_mat(self, t: int, wire: np.ndarray, x: np.ndarray, m: int, n_threads: int = 4, reps: int = 1):
        x = np.ascontiguousarray(x, np.float32)
        n, k = x.shape
        wire = np.ascontiguousarray(wire, np.uint8)
        out = np.empty((n, m), np.float32)
        sec = self.lib.refshim_mul_mat(t, _p(wire), _p(x), _p(out), m, k, n, n_threads, reps)
        assert sec >= 0, sec
        return out, sec


    def fused_up_gate(self, t: int, w_up: np.ndarray, w_gate: np.ndarray, x: np.ndarray, m: int, unary: str = "silu", limit: float = 0.0, n_threads: int = 4):
        """GGML_OP_FUSED_UP_GATE through the reference CPU backend (op_params[1] = limit)."""
        x = np.ascontiguousarray(x, np.float32)
        n, k = x.shape
        out = np.empty((n, m), np.float32)
        op = self.lib.refshim_unary_op_id(unary.encode())
        assert op >= 0, unary
        rc = self.lib.refshim_fused_up_gate(t, _p(np.ascontiguousarray(w_up, np.uint8)), _p(np.ascontiguousarray(w_gate, np.uint8)), _p(x), _p(out), m, k, n, op, float(limit), n_threads)
        assert rc == 0, rc
        return out


def nmse(a: np.ndarray, b: np.ndarray) -> float:
    """Normalised mean squared error, as tests/test-backend-ops.cpp:39-50 of the reference."""
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    return float(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300))
