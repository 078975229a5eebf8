"""GPU test of the backend-level drop-in boundary: the reference's own ggml graph API drives our backend
(ggml_backend_cuda_init -> ggml_backend_graph_compute) and ggml_backend_compare_graph_backend checks every MUL_MAT /
FUSED_UP_GATE node against the reference CPU backend (NMSE <= 5e-4) — the semantics of tests/test-backend-ops.cpp test_mul_mat."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_backend_ops_mul_mat_through_ggml_backend_api():
    exe = os.path.join(ROOT, "tests", "backend_ops", "test_mul_mat_backend")
    if not os.path.exists(exe):
        pytest.skip("harness not built (needs the reference headers at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    print(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASSED: 0 failures" in r.stdout
