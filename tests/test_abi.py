"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/b200q.h declares,
agrees with the oracle on wire geometry, and FAILS LOUDLY (no CPU fallback) when there is no GPU."""
import ctypes

import numpy as np
import pytest

import ik_llama_cpp_b200 as pkg
from conftest import ALL_TYPES
from oracle.oracle import GGML_TYPE


def test_library_exports_every_declared_symbol():
    L = pkg.lib()
    syms = pkg.header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"include/b200q.h declares {s} but libb200q.so does not export it"
    assert L.b200q_abi_version() == 1


@pytest.mark.parametrize("name", ALL_TYPES)
def test_geometry_matches_oracle(oracle, name):
    L = pkg.lib()
    t = GGML_TYPE[name]
    assert L.b200q_type_supported(t) == 1
    for k in (256, 512, 4096, 14336):
        assert L.b200q_wire_row_size(t, k) == oracle.row_size(t, k)
        m = 8 if name.endswith("_R4") else 7            # the row-interleaved repacks exist only in groups of 4 rows
        pb, wire = L.b200q_plane_bytes(t, m, k), m * oracle.row_size(t, k)
        assert wire <= pb <= wire + 5 * 256, "plane layout must not inflate the tensor"
    assert L.b200q_wire_row_size(t, 100) == -1          # K not a multiple of the block: error, like ggml's assert
    assert L.b200q_type_supported(99999) == 0


def test_bad_arguments_are_errors_not_crashes():
    L = pkg.lib()
    rc = L.b200q_mul_mat_vec(GGML_TYPE["IQ4_NL"], None, None, None, 16, 256, 1, 256, None, None)
    assert rc == -4 and b"bad argument" in L.b200q_last_error()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ik_llama_cpp_b200 import backend
    with pytest.raises(pkg.B200QError):
        backend.set_tensor(GGML_TYPE["IQ4_NL"], np.zeros(18 * 8, np.uint8), 1, 256)
    L = pkg.lib()
    assert L.b200q_device_count() == 0
    x = np.zeros(256, np.float32); y = np.zeros(16, np.float32); w = np.zeros(16 * 144, np.uint8)
    rc = L.b200q_mul_mat_vec(GGML_TYPE["IQ4_NL"], w.ctypes.data, x.ctypes.data, y.ctypes.data, 16, 256, 1, 256, None, None)
    assert rc == -3, "compute entry points must fail with B200Q_E_CUDA when there is no device"


def _plug():
    import os
    from ik_llama_cpp_b200.build import PLUG_LIB
    if not os.path.exists(PLUG_LIB):
        pytest.skip("libggml_b200.so not built (needs the reference headers at build time)")
    # the plug is a ggml plug-in: its ggml_* core symbols come from the host process.  In the tests the host's libggml is the
    # reference CPU build of oracle/_ref (test infrastructure), loaded first with RTLD_GLOBAL
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libggml_ref_avx2.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built: no libggml to host the plug")
    ctypes.CDLL(ref, mode=ctypes.RTLD_GLOBAL)
    return PLUG_LIB


def test_backend_plug_exports_the_ggml_cuda_h_symbols():
    """The backend-level boundary: every symbol of the reference's ggml-cuda.h (repeated in include/ggml-b200.h) is exported."""
    import os
    import re
    lib = ctypes.CDLL(_plug(), mode=ctypes.RTLD_GLOBAL)
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ggml-b200.h")).read()
    names = sorted(set(re.findall(r"\b(ggml_backend_(?:cuda_\w+|is_cuda))\s*\(", hdr)))
    assert len(names) == 13, names
    for n in names:
        assert hasattr(lib, n), f"libggml_b200.so does not export {n}"


def test_backend_plug_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = ctypes.CDLL(_plug(), mode=ctypes.RTLD_GLOBAL)
    lib.ggml_backend_cuda_get_device_count.restype = ctypes.c_int
    lib.ggml_backend_cuda_init.restype = ctypes.c_void_p
    lib.ggml_backend_cuda_init.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p]
    assert lib.ggml_backend_cuda_get_device_count() == 0
    assert lib.ggml_backend_cuda_init(0, None, None) is None      # init returns nullptr on a bad device, like the reference
