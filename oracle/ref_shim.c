// oracle/ref_shim.c — TEST INFRASTRUCTURE (never linked into the product).
//
// A thin plain-C window onto the UNMODIFIED reference CPU library (oracle/_ref/libggml_ref_*.so,
// built by oracle/Makefile.ref from the sources under /root/reference).  It lets Python (ctypes)
//   * quantize f32 rows into any ggml wire format      -> ggml_quantize_chunk   (ggml/include/ggml.h:3124-3132)
//   * dequantize wire rows back to f32                  -> type_traits.to_float  (ggml/include/ggml.h:3318)
//   * run GGML_OP_MUL_MAT through the reference's CPU backend (the IQK path):
//       ggml_mul_mat + ggml_backend_cpu_init + ggml_backend_graph_compute
//       (ggml/src/ggml.c:17863 ggml_compute_forward_mul_mat -> iqk_mul_mat_4d, ggml/src/iqk/iqk_mul_mat.cpp:503)
// The shim is what pins the oracle restatement (golden vectors) and what times the
// reference CPU path for bench.py's cpu_baseline / --impl reference.
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>

#define SHIM_API __attribute__((visibility("default")))

// GGML_FP16_TO_FP32 reads a table that only ggml_init() fills (ggml/src/ggml.c, first-call branch of ggml_init)
static void shim_init(void) {
    static int done = 0;
    if (done) return;
    struct ggml_init_params ip = { 1024, NULL, true };
    struct ggml_context * c = ggml_init(ip);
    if (c) ggml_free(c);
    done = 1;
}

SHIM_API int64_t refshim_blck_size(int type)            { return ggml_blck_size((enum ggml_type)type); }
SHIM_API size_t  refshim_type_size(int type)            { return ggml_type_size((enum ggml_type)type); }
SHIM_API size_t  refshim_row_size(int type, int64_t k)  { return ggml_row_size((enum ggml_type)type, k); }
SHIM_API const char * refshim_type_name(int type)       { return ggml_type_name((enum ggml_type)type); }
SHIM_API int64_t refshim_row_meta_size(int type)        { return ggml_internal_get_type_traits((enum ggml_type)type).row_meta_size; }
SHIM_API int     refshim_vec_dot_type(int type)         { return (int)ggml_internal_get_type_traits((enum ggml_type)type).vec_dot_type; }
SHIM_API int     refshim_requires_imatrix(int type)     { return ggml_quantize_requires_imatrix((enum ggml_type)type) ? 1 : 0; }

// f32 [nrows][k] -> wire bytes; imatrix may be NULL (all-ones is substituted for types that demand one,
// as tests/test-backend-ops.cpp:73 of the reference does).
SHIM_API size_t refshim_quantize(int type, const float * src, void * dst, int64_t nrows, int64_t k, const float * imatrix) {
    shim_init();
    float * ones = NULL;
    if (!imatrix && ggml_quantize_requires_imatrix((enum ggml_type)type)) {
        ones = (float *)malloc(sizeof(float)*k);
        for (int64_t i = 0; i < k; ++i) ones[i] = 1.0f;
        imatrix = ones;
    }
    size_t n = ggml_quantize_chunk((enum ggml_type)type, src, dst, 0, nrows, k, imatrix, NULL);
    free(ones);
    return n;
}

// wire bytes of ONE row -> f32 [k], via the reference's own to_float.  NOTE (SURVEY.md §8c pitfall 1):
// for types with row_meta_size > 0 the caller decides whether `src` points at the row start or past the header;
// to_float of IQ1_BN/IQ2_BN ignores the row scale.
SHIM_API int refshim_to_float(int type, const void * src, float * dst, int64_t k) {
    shim_init();
    ggml_type_traits_t tt = ggml_internal_get_type_traits((enum ggml_type)type);
    if (!tt.to_float) return -1;
    tt.to_float(src, dst, k);
    return 0;
}

// dst[n][m] = sum_k W[m][k] * x[n][k] through the reference CPU backend.  W: wire bytes (m rows of
// ggml_row_size(type,k)), x: f32 [n][k], dst: f32 [n][m].  Returns seconds of the best of `reps`
// graph_compute calls (wall clock), or <0 on error.
SHIM_API double refshim_mul_mat(int type, const void * W, const float * x, float * dst,
                                int64_t m, int64_t k, int64_t n, int n_threads, int reps) {
    struct ggml_init_params ip = { ggml_tensor_overhead()*8 + ggml_graph_overhead() + 4096, NULL, true };
    struct ggml_context * ctx = ggml_init(ip);
    if (!ctx) return -1.0;
    struct ggml_tensor * a = ggml_new_tensor_2d(ctx, (enum ggml_type)type, k, m);
    struct ggml_tensor * b = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, k, n);
    struct ggml_tensor * c = ggml_mul_mat(ctx, a, b);
    struct ggml_cgraph * gf = ggml_new_graph(ctx);
    ggml_build_forward_expand(gf, c);
    ggml_backend_t cpu = ggml_backend_cpu_init();
    if (!cpu) { ggml_free(ctx); return -2.0; }
    ggml_backend_cpu_set_n_threads(cpu, n_threads);
    ggml_backend_buffer_t buf = ggml_backend_alloc_ctx_tensors(ctx, cpu);
    if (!buf) { ggml_backend_free(cpu); ggml_free(ctx); return -3.0; }
    ggml_backend_tensor_set(a, W, 0, ggml_nbytes(a));
    ggml_backend_tensor_set(b, x, 0, ggml_nbytes(b));
    double best = 1e30;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps; ++r) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        if (ggml_backend_graph_compute(cpu, gf) != GGML_STATUS_SUCCESS) { best = -4.0; break; }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        double dt = (t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec);
        if (dt < best) best = dt;
    }
    ggml_backend_tensor_get(c, dst, 0, ggml_nbytes(c));
    ggml_backend_buffer_free(buf);
    ggml_backend_free(cpu);
    ggml_free(ctx);
    return best;
}

// A persistent multi-matrix harness for timing the reference CPU path on a model-shaped
// sequence of mat-muls without re-uploading weights: create once, run many times.
struct refshim_chain {
    struct ggml_context * ctx;
    ggml_backend_t cpu;
    ggml_backend_buffer_t buf;
    struct ggml_cgraph * gf;
    int n_mats;
    struct ggml_tensor ** a;
    struct ggml_tensor ** b;
    struct ggml_tensor ** c;
};

// types[i], m[i], k[i]: one independent MUL_MAT per entry, all with the same batch n.
SHIM_API struct refshim_chain * refshim_chain_new(int n_mats, const int * types, const int64_t * m, const int64_t * k, int64_t n, int n_threads) {
    struct refshim_chain * ch = (struct refshim_chain *)calloc(1, sizeof(*ch));
    size_t gsize = ggml_graph_overhead_custom(n_mats*4 + 64, false);
    struct ggml_init_params ip = { ggml_tensor_overhead()*(size_t)(n_mats*3 + 8) + gsize + 4096, NULL, true };
    ch->ctx = ggml_init(ip);
    ch->n_mats = n_mats;
    ch->a = (struct ggml_tensor **)calloc(n_mats, sizeof(void*));
    ch->b = (struct ggml_tensor **)calloc(n_mats, sizeof(void*));
    ch->c = (struct ggml_tensor **)calloc(n_mats, sizeof(void*));
    ch->gf = ggml_new_graph_custom(ch->ctx, n_mats*4 + 64, false);
    for (int i = 0; i < n_mats; ++i) {
        ch->a[i] = ggml_new_tensor_2d(ch->ctx, (enum ggml_type)types[i], k[i], m[i]);
        ch->b[i] = ggml_new_tensor_2d(ch->ctx, GGML_TYPE_F32, k[i], n);
        ch->c[i] = ggml_mul_mat(ch->ctx, ch->a[i], ch->b[i]);
        ggml_build_forward_expand(ch->gf, ch->c[i]);
    }
    ch->cpu = ggml_backend_cpu_init();
    ggml_backend_cpu_set_n_threads(ch->cpu, n_threads);
    ch->buf = ggml_backend_alloc_ctx_tensors(ch->ctx, ch->cpu);
    if (!ch->buf) return NULL;
    return ch;
}
SHIM_API void refshim_chain_set(struct refshim_chain * ch, int i, const void * W, const float * x) {
    if (W) ggml_backend_tensor_set(ch->a[i], W, 0, ggml_nbytes(ch->a[i]));
    if (x) ggml_backend_tensor_set(ch->b[i], x, 0, ggml_nbytes(ch->b[i]));
}
SHIM_API void refshim_chain_get(struct refshim_chain * ch, int i, float * dst) {
    ggml_backend_tensor_get(ch->c[i], dst, 0, ggml_nbytes(ch->c[i]));
}
SHIM_API double refshim_chain_run(struct refshim_chain * ch) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (ggml_backend_graph_compute(ch->cpu, ch->gf) != GGML_STATUS_SUCCESS) return -1.0;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec);
}
SHIM_API void refshim_chain_free(struct refshim_chain * ch) {
    if (!ch) return;
    ggml_backend_buffer_free(ch->buf);
    ggml_backend_free(ch->cpu);
    ggml_free(ch->ctx);
    free(ch->a); free(ch->b); free(ch->c); free(ch);
}

// GGML_OP_FUSED_UP_GATE through the reference CPU backend: dst[n][m] = unary(gate.x) * (up.x), op_params[1] = swiglu limit
// (ggml.c ggml_fused_up_gate; CPU compute ggml.c:16895-16990).  Pins the activation / clamp order of operations.
SHIM_API int refshim_fused_up_gate(int type, const void * W_up, const void * W_gate, const float * x, float * dst,
                                   int64_t m, int64_t k, int64_t n, int unary_op, float limit, int n_threads) {
    struct ggml_init_params ip = { ggml_tensor_overhead()*16 + ggml_graph_overhead() + 4096, NULL, true };
    struct ggml_context * ctx = ggml_init(ip);
    if (!ctx) return -1;
    struct ggml_tensor * up = ggml_new_tensor_2d(ctx, (enum ggml_type)type, k, m);
    struct ggml_tensor * gate = ggml_new_tensor_2d(ctx, (enum ggml_type)type, k, m);
    struct ggml_tensor * b = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, k, n);
    struct ggml_tensor * c = ggml_fused_up_gate(ctx, up, gate, b, (enum ggml_unary_op)unary_op);
    memcpy((char *)c->op_params + sizeof(int32_t), &limit, sizeof(float));
    struct ggml_cgraph * gf = ggml_new_graph(ctx);
    ggml_build_forward_expand(gf, c);
    ggml_backend_t cpu = ggml_backend_cpu_init();
    if (!cpu) { ggml_free(ctx); return -2; }
    ggml_backend_cpu_set_n_threads(cpu, n_threads);
    ggml_backend_buffer_t buf = ggml_backend_alloc_ctx_tensors(ctx, cpu);
    if (!buf) { ggml_backend_free(cpu); ggml_free(ctx); return -3; }
    ggml_backend_tensor_set(up, W_up, 0, ggml_nbytes(up));
    ggml_backend_tensor_set(gate, W_gate, 0, ggml_nbytes(gate));
    ggml_backend_tensor_set(b, x, 0, ggml_nbytes(b));
    int rc = ggml_backend_graph_compute(cpu, gf) == GGML_STATUS_SUCCESS ? 0 : -4;
    if (rc == 0) ggml_backend_tensor_get(c, dst, 0, ggml_nbytes(c));
    ggml_backend_buffer_free(buf);
    ggml_backend_free(cpu);
    ggml_free(ctx);
    return rc;
}
SHIM_API int refshim_unary_op_id(const char * name) {
    if (!strcmp(name, "silu")) return GGML_UNARY_OP_SILU;
    if (!strcmp(name, "gelu")) return GGML_UNARY_OP_GELU;
    if (!strcmp(name, "relu")) return GGML_UNARY_OP_RELU;
    if (!strcmp(name, "swiglu_oai")) return GGML_UNARY_OP_SWIGLU_OAI;
    return -1;
}
