// oracle/oracle_quants.c — TEST INFRASTRUCTURE.  CPU restatement of the reference's arithmetic for
// GGML_OP_MUL_MAT on block-quantized weights.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg may load this; the product (libb200q.so) never does and has no CPU fallback.
//
// What is restated, and from where (paths relative to /root/reference):
//   * wire formats            ggml/src/ggml-common.h:166-775 (block_* structs), tables :2212-2250
//   * dequantize_row_<type>   ggml/src/ggml-quants.c (legacy / K / IQ types), ggml/src/iqk/iqk_quantize.cpp (IQK types)
//   * quantize_q8_1 (CUDA)    ggml/src/ggml-cuda/quantize.cu:13-47
//   * MMVQ result             ggml/src/ggml-cuda/mmvq-templates.cuh:68-150 + vecdotq.cuh: for every type the
//                             kernel evaluates  sum_k dequant(W)[m][k] * (d8_b * q8[k])  with integer
//                             partial sums, i.e. the exact dot of the dequantized weights with the
//                             q8_1-dequantized activations (d8 rounded to half) — oracle_mul_mat_q8_1 below.
//   * exact result            dst[n][m] = sum_k dequant(W)[m][k] * x[n][k] in f64 — the ground truth the
//                             reference's own test uses modulo NMSE (tests/test-backend-ops.cpp:979-981).
// Pinning: tests/test_oracle.py checks every dequantizer here bit-for-bit against the reference's own
// to_float (oracle/_ref, when built) and against committed golden vectors in tests/golden/.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>

#define ORACLE_API __attribute__((visibility("default")))

// ggml_type ids (ggml/include/ggml.h:391-492)
enum {
    T_F32 = 0, T_F16 = 1, T_Q4_0 = 2, T_Q4_1 = 3, T_Q5_0 = 6, T_Q5_1 = 7, T_Q8_0 = 8,
    T_Q2_K = 10, T_Q3_K = 11, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14,
    T_IQ2_XXS = 16, T_IQ2_XS = 17, T_IQ3_XXS = 18, T_IQ3_S = 21, T_IQ2_S = 22, T_IQ6_K = 141, T_IQ1_BN = 134, T_IQ4_KSS = 146, T_IQ4_NL = 20, T_IQ4_XS = 23, T_Q6_0 = 133, T_IQ2_BN = 135, T_IQ2_K = 137, T_IQ3_K = 138, T_MXFP4 = 39, T_IQ5_KS = 152, T_IQ2_KS = 145, T_IQ3_KS = 156,
    T_IQ4_K = 139, T_IQ5_K = 140, T_IQ4_KS = 144,
    T_IQ1_S = 19, T_IQ1_M = 29, T_IQ2_KT = 153, T_IQ3_KT = 154, T_IQ4_KT = 155, T_IQ2_KL = 157, T_IQ1_KT = 158,
    // row-interleaved (x4) repacks the CUDA back-end accepts (ggml-cuda.cu:4906-4913)
    T_IQ1_S_R4 = 219, T_IQ1_M_R4 = 229, T_IQ2_K_R4 = 337, T_IQ3_K_R4 = 338, T_IQ4_K_R4 = 339, T_IQ5_K_R4 = 340, T_IQ4_KS_R4 = 344, T_IQ5_KS_R4 = 352,
};

static float h2f(uint16_t h) {                    // IEEE binary16 -> binary32 (GGML_FP16_TO_FP32)
    uint32_t s = (uint32_t)(h & 0x8000) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ff, u;
    if (e == 0) {
        if (m == 0) u = s;
        else { int sh = 0; while (!(m & 0x400)) { m <<= 1; ++sh; } m &= 0x3ff; u = s | ((uint32_t)(113 - sh) << 23) | (m << 13); }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}
static uint16_t f2h(float f) {                    // binary32 -> binary16, round-to-nearest-even (__float2half_rn)
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t s = (x >> 16) & 0x8000; int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15; uint32_t m = x & 0x7fffff;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(s | 0x7c00 | (m ? 0x200 : 0));
    if (e >= 31) return (uint16_t)(s | 0x7c00);
    if (e <= 0) {
        if (e < -10) return (uint16_t)s;
        m |= 0x800000; int sh = 14 - e; uint32_t r = m >> sh, rem = m & ((1u << sh) - 1), half = 1u << (sh - 1);
        if (rem > half || (rem == half && (r & 1))) ++r;
        return (uint16_t)(s | r);
    }
    uint32_t r = (uint32_t)(e << 10) | (m >> 13), rem = m & 0x1fff;
    if (rem > 0x1000 || (rem == 0x1000 && (r & 1))) ++r;
    return (uint16_t)(s | r);
}
ORACLE_API float    oracle_h2f(uint16_t h) { return h2f(h); }
ORACLE_API uint16_t oracle_f2h(float f)    { return f2h(f); }

static uint16_t rd16(const uint8_t * p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// non-linear value tables (ggml-common.h: kvalues_iq4nl, iq4k_values :2227, iq5nl_values :2232)
static const int8_t k_iq4nl[16] = {-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113};
static int8_t k_iq4k[32];
static int8_t k_iq5nl[64];
static int tables_ready = 0;
static void init_tables(void) {
    if (tables_ready) return;
    // iq4k_values = kvalues_iq4nl, then the same +4 (second row of the table at ggml-common.h:2227-2230)
    for (int i = 0; i < 16; ++i) { k_iq4k[i] = k_iq4nl[i]; k_iq4k[16 + i] = (int8_t)(k_iq4nl[i] + 4); }
    // iq5nl_values (ggml-common.h:2232-2235): 32 values, then the same +2
    static const int8_t v5[32] = {-126, -114, -103, -92, -83, -74, -65, -57, -50, -43, -36, -30, -24, -18, -12, -6,
                                  -1, 5, 11, 17, 23, 29, 36, 43, 51, 59, 68, 77, 87, 97, 109, 121};
    for (int i = 0; i < 32; ++i) { k_iq5nl[i] = v5[i]; k_iq5nl[32 + i] = (int8_t)(v5[i] + 2); }
    tables_ready = 1;
}

// ---- IQ2_XXS codebook: NOT restated from the reference sources.  The 256 x 8 magnitude grid and the 128 sign masks are extracted by
// running the reference's own to_float on crafted blocks (tests/golden/gen_codebooks.py -> tests/golden/iq2xxs_codebook.npz) and handed
// in at run time.  Device kernel for this type: round 2; the oracle is ready and pinned.
static const uint8_t * g_iq2xxs_grid = NULL;     // [256][8]
static const uint8_t * g_iq2xxs_signs = NULL;    // [128]
static const uint8_t * g_iq2xs_grid = NULL;      // [512][8]
static const uint8_t * g_iq3xxs_grid = NULL;     // [256][4]
static const uint8_t * g_iq2s_grid = NULL;       // [1024][8]
static const uint8_t * g_iq3s_grid = NULL;       // [512][4]
static const int8_t * g_iq1s_grid = NULL;        // [2048][8], values in {-1, 0, 1} (IQ1_S / IQ1_M and their _R4 repacks)
static const int8_t * g_iq2kl = NULL;            // [32][2] value pairs of IQ2_KL
ORACLE_API void oracle_set_iq2xxs_codebook(const uint8_t * grid, const uint8_t * ksigns) { g_iq2xxs_grid = grid; g_iq2xxs_signs = ksigns; }
ORACLE_API void oracle_set_grid(int which, const uint8_t * grid) { if (which == 17) g_iq2xs_grid = grid; else if (which == 18) g_iq3xxs_grid = grid; else if (which == 22) g_iq2s_grid = grid; else if (which == 21) g_iq3s_grid = grid; else if (which == 19) g_iq1s_grid = (const int8_t *)grid; else if (which == 157) g_iq2kl = (const int8_t *)grid; }

// ---- wire geometry: {block elements, block bytes, row meta bytes} (ggml.c type_traits :640-1460) ----
static int geom(int type, int * qk, int * bs, int * meta) {
    *meta = 0;
    switch (type) {
        case T_Q4_0:   *qk = 32;  *bs = 18;  return 0;
        case T_Q4_1:   *qk = 32;  *bs = 20;  return 0;
        case T_Q5_0:   *qk = 32;  *bs = 22;  return 0;
        case T_Q5_1:   *qk = 32;  *bs = 24;  return 0;
        case T_Q6_0:   *qk = 32;  *bs = 26;  return 0;
        case T_Q8_0:   *qk = 32;  *bs = 34;  return 0;
        case T_IQ2_XXS: if (!g_iq2xxs_grid) return -1; *qk = 256; *bs = 66; return 0;     // needs the codebook fixture (oracle_set_iq2xxs_codebook)
        case T_IQ2_XS:  if (!g_iq2xs_grid || !g_iq2xxs_signs) return -1; *qk = 256; *bs = 74; return 0;
        case T_IQ3_XXS: if (!g_iq3xxs_grid || !g_iq2xxs_signs) return -1; *qk = 256; *bs = 98; return 0;
        case T_IQ2_S:   if (!g_iq2s_grid) return -1; *qk = 256; *bs = 82; return 0;
        case T_IQ3_S:   if (!g_iq3s_grid) return -1; *qk = 256; *bs = 110; return 0;
        case T_IQ6_K:   *qk = 256; *bs = 212; return 0;
        case T_IQ1_BN:  *qk = 64;  *bs = 13;  *meta = 2; return 0;
        case T_IQ4_KSS: *qk = 256; *bs = 128; *meta = 4; return 0;
        case T_Q2_K:   *qk = 256; *bs = 84;  return 0;
        case T_Q3_K:   *qk = 256; *bs = 110; return 0;
        case T_Q4_K:   *qk = 256; *bs = 144; return 0;
        case T_Q5_K:   *qk = 256; *bs = 176; return 0;
        case T_Q6_K:   *qk = 256; *bs = 210; return 0;
        case T_IQ4_NL: *qk = 32;  *bs = 18;  return 0;
        case T_IQ4_XS: *qk = 256; *bs = 136; return 0;
        case T_IQ2_K:  *qk = 256; *bs = 76;  return 0;
        case T_IQ3_K:  *qk = 256; *bs = 110; return 0;
        case T_IQ4_K:  *qk = 256; *bs = 144; return 0;
        case T_IQ5_K:  *qk = 256; *bs = 176; return 0;
        case T_IQ4_KS: *qk = 256; *bs = 136; *meta = 4; return 0;
        case T_IQ5_KS: *qk = 256; *bs = 168; *meta = 4; return 0;
        case T_MXFP4:  *qk = 32;  *bs = 17;  return 0;
        case T_IQ2_KS: *qk = 256; *bs = 70;  *meta = 2; return 0;
        case T_IQ3_KS: *qk = 256; *bs = 102; *meta = 2; return 0;
        case T_IQ2_BN: *qk = 64;  *bs = 16;  *meta = 4; return 0;
        case T_IQ1_S:  if (!g_iq1s_grid) return -1; *qk = 256; *bs = 50; return 0;
        case T_IQ1_M:  if (!g_iq1s_grid) return -1; *qk = 256; *bs = 56; return 0;
        case T_IQ2_KL: if (!g_iq2kl) return -1; *qk = 256; *bs = 86; *meta = 2; return 0;
        case T_IQ1_KT: *qk = 256; *bs = 56;  *meta = 4; return 0;
        case T_IQ2_KT: *qk = 256; *bs = 68;  *meta = 4; return 0;
        case T_IQ3_KT: *qk = 256; *bs = 100; *meta = 4; return 0;
        case T_IQ4_KT: *qk = 256; *bs = 128; *meta = 4; return 0;
        // _R4: per-ROW figures (the wire interleaves 4 rows: a group of 4 rows = 4 x row bytes, the 4 row scales first)
        case T_IQ1_S_R4: if (!g_iq1s_grid) return -1; *qk = 32; *bs = 6; *meta = 2; return 0;
        case T_IQ1_M_R4: if (!g_iq1s_grid) return -1; *qk = 32; *bs = 7; *meta = 2; return 0;
        case T_IQ2_K_R4: *qk = 256; *bs = 76;  return 0;
        case T_IQ3_K_R4: *qk = 256; *bs = 110; return 0;
        case T_IQ4_K_R4: *qk = 256; *bs = 144; return 0;
        case T_IQ5_K_R4: *qk = 256; *bs = 176; return 0;
        case T_IQ4_KS_R4: *qk = 256; *bs = 136; *meta = 4; return 0;
        case T_IQ5_KS_R4: *qk = 256; *bs = 168; *meta = 4; return 0;
        default: return -1;
    }
}
ORACLE_API int64_t oracle_row_size(int type, int64_t k) {
    int qk, bs, meta; if (geom(type, &qk, &bs, &meta) || k % qk) return -1;
    return (int64_t)meta + (k / qk) * bs;
}
ORACLE_API int oracle_type_supported(int type) { int a, b, c; return geom(type, &a, &b, &c) == 0; }
// rows interleaved on the wire: 1, or 4 for the _R4 repacks (then rows are addressable only in groups of 4)
ORACLE_API int oracle_rows_interleaved(int type) {
    switch (type) { case T_IQ1_S_R4: case T_IQ1_M_R4: case T_IQ2_K_R4: case T_IQ3_K_R4: case T_IQ4_K_R4: case T_IQ5_K_R4: case T_IQ4_KS_R4: case T_IQ5_KS_R4: return 4; default: return 1; }
}
// trellis generator of the IQx_KT types (QuantizerIQKT<...>::set_values, integer variant, iqk/iqk_quantize.cpp:8626-8640):
// x <- 0xCBAC1FED * x; value = (sum of the four 6-bit fields of x) - 126
static void kt_values(uint32_t idx, uint32_t offset, int n, float scale, int is_abs, float * out) {
    uint32_t x = idx + offset;
    for (int k = 0; k < n; ++k) {
        x *= 0xCBAC1FEDu;
        const uint32_t s = x & 0x3f3f3f3fu;
        const float v = (float)((int)(s & 0xff) + (int)((s >> 8) & 0xff) + (int)((s >> 16) & 0xff) + (int)(s >> 24)) - 126.f;
        out[k] = scale * (is_abs ? fabsf(v) : v);
    }
}

// get_scale_min_k4 (ggml-quants.c:2036-2044)
static void scale_min_k4(int j, const uint8_t * q, int * sc, int * m) {
    if (j < 4) { *sc = q[j] & 63; *m = q[j + 4] & 63; }
    else { *sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); *m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

// Dequantize one wire row (pointer at row start, i.e. including any row-meta header) into y[k].
ORACLE_API int oracle_dequantize_row(int type, const uint8_t * row, float * y, int64_t k) {
    int qk, bs, meta; if (geom(type, &qk, &bs, &meta) || k % qk) return -1;
    init_tables();
    const int64_t nb = k / qk;
    const uint8_t * x = row + meta;
    float row_scale = 1.0f;
    if (meta == 4) memcpy(&row_scale, row, 4);
    if (meta == 2) row_scale = h2f(rd16(row));           // IQ2_KS / IQ3_KS: ggml_half row scale
    for (int64_t i = 0; i < nb; ++i, x += bs, y += qk) {
        switch (type) {
        case T_Q4_0: {  // ggml-quants.c:1581-1599  {half d; u8 qs[16]}
            const float d = h2f(rd16(x)); const uint8_t * qs = x + 2;
            for (int j = 0; j < 16; ++j) { y[j] = ((qs[j] & 0xF) - 8) * d; y[j + 16] = ((qs[j] >> 4) - 8) * d; }
        } break;
        case T_Q4_1: {  // ggml-quants.c:1601-1620  {half d, m; u8 qs[16]}
            const float d = h2f(rd16(x)), m = h2f(rd16(x + 2)); const uint8_t * qs = x + 4;
            for (int j = 0; j < 16; ++j) { y[j] = (qs[j] & 0xF) * d + m; y[j + 16] = (qs[j] >> 4) * d + m; }
        } break;
        case T_Q5_0: {  // ggml-quants.c:1622-1646  {half d; u8 qh[4]; u8 qs[16]}
            const float d = h2f(rd16(x)); uint32_t qh; memcpy(&qh, x + 2, 4); const uint8_t * qs = x + 6;
            for (int j = 0; j < 16; ++j) {
                const uint8_t xh0 = ((qh >> (j + 0)) << 4) & 0x10, xh1 = ((qh >> (j + 12))) & 0x10;
                y[j] = (((qs[j] & 0xF) | xh0) - 16) * d; y[j + 16] = (((qs[j] >> 4) | xh1) - 16) * d;
            }
        } break;
        case T_Q5_1: {  // ggml-quants.c:1648-1673
            const float d = h2f(rd16(x)), m = h2f(rd16(x + 2)); uint32_t qh; memcpy(&qh, x + 4, 4); const uint8_t * qs = x + 8;
            for (int j = 0; j < 16; ++j) {
                const uint8_t xh0 = ((qh >> (j + 0)) << 4) & 0x10, xh1 = ((qh >> (j + 12))) & 0x10;
                y[j] = ((qs[j] & 0xF) | xh0) * d + m; y[j + 16] = ((qs[j] >> 4) | xh1) * d + m;
            }
        } break;
        case T_Q6_0: {  // ggml-quants.c:1675-1695  {half d; u8 qh[8]; u8 qs[16]}
            const float d = h2f(rd16(x)); const uint8_t * qh = x + 2; const uint8_t * qs = x + 10;
            for (int j = 0; j < 16; ++j) {
                const uint8_t h = qh[j % 8] >> 4 * (j / 8);
                y[j] = (((qs[j] & 0xF) | ((h << 4) & 0x30)) - 32) * d;
                y[j + 16] = (((qs[j] >> 4) | ((h << 2) & 0x30)) - 32) * d;
            }
        } break;
        case T_Q8_0: {  // ggml-quants.c:1697-1711  {half d; i8 qs[32]}
            const float d = h2f(rd16(x)); const int8_t * qs = (const int8_t *)(x + 2);
            for (int j = 0; j < 32; ++j) y[j] = qs[j] * d;
        } break;
        case T_Q4_K: {  // ggml-quants.c:2797-2822  {half d, dmin; u8 scales[12]; u8 qs[128]}
            const float d = h2f(rd16(x)), mn = h2f(rd16(x + 2)); const uint8_t * sc = x + 4; const uint8_t * q = x + 16;
            float * yy = y; int is = 0;
            for (int j = 0; j < 256; j += 64) {
                int s1, m1, s2, m2; scale_min_k4(is, sc, &s1, &m1); scale_min_k4(is + 1, sc, &s2, &m2);
                const float d1 = d * s1, mm1 = mn * m1, d2 = d * s2, mm2 = mn * m2;
                for (int l = 0; l < 32; ++l) *yy++ = d1 * (q[l] & 0xF) - mm1;
                for (int l = 0; l < 32; ++l) *yy++ = d2 * (q[l] >> 4) - mm2;
                q += 32; is += 2;
            }
        } break;
        case T_Q5_K: {  // ggml-quants.c:3015-3041  {half d, dmin; u8 scales[12]; u8 qh[32]; u8 qs[128]}
            const float d = h2f(rd16(x)), mn = h2f(rd16(x + 2)); const uint8_t * sc = x + 4; const uint8_t * qh = x + 16; const uint8_t * ql = x + 48;
            float * yy = y; int is = 0; uint8_t u1 = 1, u2 = 2;
            for (int j = 0; j < 256; j += 64) {
                int s1, m1, s2, m2; scale_min_k4(is, sc, &s1, &m1); scale_min_k4(is + 1, sc, &s2, &m2);
                const float d1 = d * s1, mm1 = mn * m1, d2 = d * s2, mm2 = mn * m2;
                for (int l = 0; l < 32; ++l) *yy++ = d1 * ((ql[l] & 0xF) + (qh[l] & u1 ? 16 : 0)) - mm1;
                for (int l = 0; l < 32; ++l) *yy++ = d2 * ((ql[l] >> 4) + (qh[l] & u2 ? 16 : 0)) - mm2;
                ql += 32; is += 2; u1 <<= 2; u2 <<= 2;
            }
        } break;
        case T_IQ2_XXS: {  // ggml-quants.c:3674-3698  {half d; u16 qs[32]}: per 32 weights [4 grid indices][4 x 7-bit sign index | scale << 28]
            const float d = h2f(rd16(x)); float * yy = y;
            for (int ib32 = 0; ib32 < 8; ++ib32) {
                uint32_t aux32[2]; memcpy(aux32, x + 2 + 8 * ib32, 8); const uint8_t * aux8 = (const uint8_t *)aux32;
                const float db = d * (0.5f + (aux32[1] >> 28)) * 0.25f;
                for (int l = 0; l < 4; ++l) {
                    const uint8_t * grid = g_iq2xxs_grid + 8 * aux8[l];
                    const uint8_t signs = g_iq2xxs_signs[(aux32[1] >> 7 * l) & 127];
                    for (int j = 0; j < 8; ++j) yy[j] = db * grid[j] * (signs & (1 << j) ? -1.f : 1.f);
                    yy += 8;
                }
            }
        } break;
        case T_IQ2_XS: {  // ggml-quants.c:3702-3725  {half d; u16 qs[32]; u8 scales[8]}
            const float d = h2f(rd16(x)); const uint8_t * sc = x + 66; float * yy = y;
            for (int ib32 = 0; ib32 < 8; ++ib32) {
                const float db[2] = { d * (0.5f + (sc[ib32] & 0xf)) * 0.25f, d * (0.5f + (sc[ib32] >> 4)) * 0.25f };
                for (int l = 0; l < 4; ++l) {
                    const uint16_t q = rd16(x + 2 + 2 * (4 * ib32 + l));
                    const uint8_t * grid = g_iq2xs_grid + 8 * (q & 511); const uint8_t signs = g_iq2xxs_signs[q >> 9];
                    for (int j = 0; j < 8; ++j) yy[j] = db[l / 2] * grid[j] * (signs & (1 << j) ? -1.f : 1.f);
                    yy += 8;
                }
            }
        } break;
        case T_IQ3_XXS: {  // ggml-quants.c:3761-3789  {half d; u8 qs[64]; u8 scales_and_signs[32]}
            const float d = h2f(rd16(x)); const uint8_t * qs = x + 2; const uint8_t * sas = x + 66; float * yy = y;
            for (int ib32 = 0; ib32 < 8; ++ib32) {
                uint32_t aux32; memcpy(&aux32, sas + 4 * ib32, 4);
                const float db = d * (0.5f + (aux32 >> 28)) * 0.5f;
                for (int l = 0; l < 4; ++l) {
                    const uint8_t signs = g_iq2xxs_signs[(aux32 >> 7 * l) & 127];
                    const uint8_t * g1 = g_iq3xxs_grid + 4 * qs[2 * l + 0]; const uint8_t * g2 = g_iq3xxs_grid + 4 * qs[2 * l + 1];
                    for (int j = 0; j < 4; ++j) { yy[j] = db * g1[j] * (signs & (1 << j) ? -1.f : 1.f); yy[j + 4] = db * g2[j] * (signs & (1 << (j + 4)) ? -1.f : 1.f); }
                    yy += 8;
                }
                qs += 8;
            }
        } break;
        case T_IQ2_S: {  // ggml-quants.c:3727-3757  {half d; u8 qs[64]; u8 qh[8]; u8 scales[8]}, qs[32..63] = sign bytes
            const float d = h2f(rd16(x)); const uint8_t * qs = x + 2; const uint8_t * signs = qs + 32; const uint8_t * qh = x + 66; const uint8_t * sc = x + 74; float * yy = y;
            for (int ib32 = 0; ib32 < 8; ++ib32) {
                const float db[2] = { d * (0.5f + (sc[ib32] & 0xf)) * 0.25f, d * (0.5f + (sc[ib32] >> 4)) * 0.25f };
                for (int l = 0; l < 4; ++l) {
                    const uint8_t * grid = g_iq2s_grid + 8 * (qs[l] | ((qh[ib32] << (8 - 2 * l)) & 0x300));
                    for (int j = 0; j < 8; ++j) yy[j] = db[l / 2] * grid[j] * (signs[l] & (1 << j) ? -1.f : 1.f);
                    yy += 8;
                }
                qs += 4; signs += 4;
            }
        } break;
        case T_IQ3_S: {  // ggml-quants.c:3793-3835  {half d; u8 qs[64]; u8 qh[8]; u8 signs[32]; u8 scales[4]}
            const float d = h2f(rd16(x)); const uint8_t * qs = x + 2; const uint8_t * qh = x + 66; const uint8_t * signs = x + 74; const uint8_t * sc = x + 106; float * yy = y;
            for (int ib32 = 0; ib32 < 8; ib32 += 2) {
                const float dbs[2] = { d * (1 + 2 * (sc[ib32 / 2] & 0xf)), d * (1 + 2 * (sc[ib32 / 2] >> 4)) };
                for (int half = 0; half < 2; ++half) {
                    for (int l = 0; l < 4; ++l) {
                        const uint8_t * g1 = g_iq3s_grid + 4 * (qs[2 * l + 0] | ((qh[half] << (8 - 2 * l)) & 256));
                        const uint8_t * g2 = g_iq3s_grid + 4 * (qs[2 * l + 1] | ((qh[half] << (7 - 2 * l)) & 256));
                        for (int j = 0; j < 4; ++j) { yy[j] = dbs[half] * g1[j] * (signs[l] & (1 << j) ? -1.f : 1.f); yy[j + 4] = dbs[half] * g2[j] * (signs[l] & (1 << (j + 4)) ? -1.f : 1.f); }
                        yy += 8;
                    }
                    qs += 8; signs += 4;
                }
                qh += 2;
            }
        } break;
        case T_IQ6_K: {  // iqk/iqk_quantize.cpp:3442-3486  {half d; u16 extra; i8 scales[16]; u8 qs[128]; u8 qh[64]}: cubic codebook A + q(B + q(-C + qD)) (+1 with the extra bit)
            const float A = -127.f, B = 6.2568f, C = 0.11218f, D = 0.0011972f, S = 1.f;
            const float d = h2f(rd16(x)); uint16_t extra = rd16(x + 2); const int8_t * sl = (const int8_t *)(x + 4); const uint8_t * qs = x + 20; const uint8_t * qh = x + 148; float * yy = y;
            int shift = 0;
            for (int ib64 = 0; ib64 < 4; ++ib64) {
                const float dl1 = d * sl[4 * ib64 + 0], dl2 = d * sl[4 * ib64 + 1], dl3 = d * sl[4 * ib64 + 2], dl4 = d * sl[4 * ib64 + 3];
                const float m1 = extra & 1 ? S : 0, m2 = extra & 2 ? S : 0, m3 = extra & 4 ? S : 0, m4 = extra & 8 ? S : 0;
                for (int j = 0; j < 16; ++j) {
                    const float q1 = ((qs[j] & 0xf) | (((qh[j] >> shift) & 0x03) << 4)), q2 = ((qs[j + 16] & 0xf) | (((qh[j + 16] >> shift) & 0x03) << 4));
                    const float q3 = ((qs[j] >> 4) | (((qh[j] >> shift) & 0x0c) << 2)), q4 = ((qs[j + 16] >> 4) | (((qh[j + 16] >> shift) & 0x0c) << 2));
                    yy[j]      = dl1 * (A + q1 * (B + q1 * (-C + q1 * D)) + m1);
                    yy[j + 16] = dl2 * (A + q2 * (B + q2 * (-C + q2 * D)) + m2);
                    yy[j + 32] = dl3 * (A + q3 * (B + q3 * (-C + q3 * D)) + m3);
                    yy[j + 48] = dl4 * (A + q4 * (B + q4 * (-C + q4 * D)) + m4);
                }
                yy += 64; qs += 32; extra >>= 4; shift += 4;
                if (shift == 8) { qh += 32; shift = 0; }
            }
        } break;
        case T_Q2_K: {  // ggml-quants.c:2162-2190  {u8 scales[16]; u8 qs[64]; half d, dmin}
            const uint8_t * sc = x; const uint8_t * q = x + 16; const float d = h2f(rd16(x + 80)), dmin = h2f(rd16(x + 82));
            float * yy = y; int is = 0;
            for (int n = 0; n < 256; n += 128) {
                int shift = 0;
                for (int j = 0; j < 4; ++j) {
                    float dl = d * (sc[is] & 0xF), ml = dmin * (sc[is] >> 4); ++is;
                    for (int l = 0; l < 16; ++l) *yy++ = dl * ((int8_t)((q[l] >> shift) & 3)) - ml;
                    dl = d * (sc[is] & 0xF); ml = dmin * (sc[is] >> 4); ++is;
                    for (int l = 0; l < 16; ++l) *yy++ = dl * ((int8_t)((q[l + 16] >> shift) & 3)) - ml;
                    shift += 2;
                }
                q += 32;
            }
        } break;
        case T_Q3_K: {  // ggml-quants.c:2563-2605  {u8 hmask[32]; u8 qs[64]; u8 scales[12]; half d}
            const uint8_t * hm = x; const uint8_t * q = x + 32; const float d_all = h2f(rd16(x + 108));
            uint32_t aux[4]; memcpy(aux, x + 96, 12);
            const uint32_t kmask1 = 0x03030303, kmask2 = 0x0f0f0f0f; const uint32_t tmp = aux[2];
            aux[2] = ((aux[0] >> 4) & kmask2) | (((tmp >> 4) & kmask1) << 4);
            aux[3] = ((aux[1] >> 4) & kmask2) | (((tmp >> 6) & kmask1) << 4);
            aux[0] = (aux[0] & kmask2) | (((tmp >> 0) & kmask1) << 4);
            aux[1] = (aux[1] & kmask2) | (((tmp >> 2) & kmask1) << 4);
            const int8_t * scales = (const int8_t *)aux;
            float * yy = y; int is = 0; uint8_t m = 1;
            for (int n = 0; n < 256; n += 128) {
                int shift = 0;
                for (int j = 0; j < 4; ++j) {
                    float dl = d_all * (scales[is++] - 32);
                    for (int l = 0; l < 16; ++l) *yy++ = dl * ((int8_t)((q[l] >> shift) & 3) - ((hm[l] & m) ? 0 : 4));
                    dl = d_all * (scales[is++] - 32);
                    for (int l = 0; l < 16; ++l) *yy++ = dl * ((int8_t)((q[l + 16] >> shift) & 3) - ((hm[l + 16] & m) ? 0 : 4));
                    shift += 2; m <<= 1;
                }
                q += 32;
            }
        } break;
        case T_Q6_K: {  // ggml-quants.c:3231-3260  {u8 ql[128]; u8 qh[64]; i8 scales[16]; half d}
            const uint8_t * ql = x; const uint8_t * qh = x + 128; const int8_t * sc = (const int8_t *)(x + 192); const float d = h2f(rd16(x + 208));
            float * yy = y;
            for (int n = 0; n < 256; n += 128) {
                for (int l = 0; l < 32; ++l) {
                    const int is = l / 16;
                    const int8_t q1 = (int8_t)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                    const int8_t q2 = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    const int8_t q3 = (int8_t)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                    const int8_t q4 = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    yy[l] = d * sc[is] * q1; yy[l + 32] = d * sc[is + 2] * q2; yy[l + 64] = d * sc[is + 4] * q3; yy[l + 96] = d * sc[is + 6] * q4;
                }
                yy += 128; ql += 64; qh += 32; sc += 8;
            }
        } break;
        case T_IQ4_NL: {  // ggml-quants.c:3913-3929  {half d; u8 qs[16]}
            const float d = h2f(rd16(x)); const uint8_t * qs = x + 2;
            for (int j = 0; j < 16; ++j) { y[j] = d * k_iq4nl[qs[j] & 0xf]; y[j + 16] = d * k_iq4nl[qs[j] >> 4]; }
        } break;
        case T_IQ4_XS: {  // ggml-quants.c:3931-3954  {half d; u16 scales_h; u8 scales_l[4]; u8 qs[128]}
            const float d = h2f(rd16(x)); const uint16_t sh = rd16(x + 2); const uint8_t * sl = x + 4; const uint8_t * qs = x + 8;
            float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                const int ls = ((sl[ib / 2] >> 4 * (ib % 2)) & 0xf) | (((sh >> 2 * ib) & 3) << 4);
                const float dl = d * (ls - 32);
                for (int j = 0; j < 16; ++j) { yy[j] = dl * k_iq4nl[qs[j] & 0xf]; yy[j + 16] = dl * k_iq4nl[qs[j] >> 4]; }
                yy += 32; qs += 16;
            }
        } break;
        case T_IQ2_K: {  // iqk/iqk_quantize.cpp:1356-1385  {half d; u16 extra; u8 scales[8]; u8 qs[64]}; iq2nl_values ggml-common.h:2212
            static const int8_t v2[8] = {-31, -13, 1, 17, -26, -8, 6, 22};
            const float d = h2f(rd16(x)); uint16_t extra = rd16(x + 2); const uint8_t * sc = x + 4; const uint8_t * qs = x + 12;
            float * yy = y; int shift = 0;
            for (int ib = 0; ib < 8; ++ib) {
                const float dl1 = d * ((sc[ib] & 0xf) - 8), dl2 = d * ((sc[ib] >> 4) - 8);
                const int8_t * va = extra & 1 ? v2 + 4 : v2; const int8_t * vb = extra & 2 ? v2 + 4 : v2; extra >>= 2;
                for (int j = 0; j < 16; ++j) { yy[j] = dl1 * va[(qs[j] >> shift) & 3]; yy[j + 16] = dl2 * vb[(qs[j + 16] >> shift) & 3]; }
                yy += 32; shift += 2; if (shift == 8) { qs += 32; shift = 0; }
            }
        } break;
        case T_IQ3_K: {  // iqk/iqk_quantize.cpp:2534-2565  {half d; u16 extra; u16 scales_h; u8 scales_l[8]; u8 qs[64]; u8 qh[32]}; iq3nl_values :2222
            static const int8_t v3[16] = {-63, -40, -23, -10, 1, 13, 28, 47, -59, -36, -19, -6, 5, 17, 32, 51};
            const float d = h2f(rd16(x)); uint16_t extra = rd16(x + 2); uint16_t sh = rd16(x + 4); const uint8_t * sl = x + 6; const uint8_t * qs = x + 14; const uint8_t * qh = x + 78;
            float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                const float dl1 = d * ((2 * (sl[ib] & 0xf) + 1) * ((sh & 1) ? -1 : 1)), dl2 = d * ((2 * (sl[ib] >> 4) + 1) * ((sh & 2) ? -1 : 1)); sh >>= 2;
                const int8_t * va = extra & 1 ? v3 + 8 : v3; const int8_t * vb = extra & 2 ? v3 + 8 : v3; extra >>= 2;
                const int shl = 2 * (ib % 4), shh = ib % 8;
                for (int j = 0; j < 16; ++j) {
                    yy[j]      = dl1 * va[((qs[j] >> shl) & 3) | (((qh[j] >> shh) & 1) << 2)];
                    yy[j + 16] = dl2 * vb[((qs[j + 16] >> shl) & 3) | (((qh[j + 16] >> shh) & 1) << 2)];
                }
                yy += 32; if (shl == 6) qs += 32;
            }
        } break;
        case T_IQ4_K: {  // iqk/iqk_quantize.cpp:2822-2850  {half d; u16 extra; u8 scales_h[4]; u8 scales_l[8]; u8 qs[128]}
            const float d = h2f(rd16(x)); uint16_t extra = rd16(x + 2); const uint8_t * sh = x + 4; const uint8_t * sl = x + 8; const uint8_t * qs = x + 16;
            float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                const uint8_t h = sh[ib / 2] >> 4 * (ib % 2);
                const int ls1 = ((sl[ib] & 0xf) | ((h << 4) & 0x30)) - 32;
                const int ls2 = ((sl[ib] >> 4) | ((h << 2) & 0x30)) - 32;
                const float dl1 = d * ls1, dl2 = d * ls2;
                const int8_t * v1 = k_iq4k + ((extra & 1) << 4); const int8_t * v2 = k_iq4k + ((extra & 2) << 3); extra >>= 2;
                for (int j = 0; j < 16; ++j) { yy[j] = dl1 * v1[qs[j] & 0xf]; yy[j + 16] = dl2 * v2[qs[j] >> 4]; }
                yy += 32; qs += 16;
            }
        } break;
        case T_IQ5_K: {  // iqk/iqk_quantize.cpp:3112-3150  {half d; u16 extra; u8 scales_h[4]; u8 scales_l[8]; u8 qs[128]; u8 qh[32]}
            const float d = h2f(rd16(x)); uint16_t extra = rd16(x + 2); const uint8_t * sh = x + 4; const uint8_t * sl = x + 8;
            const uint8_t * qs = x + 16; const uint8_t * qh = x + 144;
            float * yy = y; int shift = 0;
            for (int ib64 = 0; ib64 < 4; ++ib64) {
                const float dl1 = d * (((sl[2 * ib64] & 0xf) | ((sh[ib64] << 4) & 0x30)) - 32);
                const float dl2 = d * (((sl[2 * ib64] >> 4) | ((sh[ib64] << 2) & 0x30)) - 32);
                const float dl3 = d * (((sl[2 * ib64 + 1] & 0xf) | ((sh[ib64] >> 0) & 0x30)) - 32);
                const float dl4 = d * (((sl[2 * ib64 + 1] >> 4) | ((sh[ib64] >> 2) & 0x30)) - 32);
                const int8_t * v1 = k_iq5nl + ((extra & 1) << 5); const int8_t * v2 = k_iq5nl + ((extra & 2) << 4);
                const int8_t * v3 = k_iq5nl + ((extra & 4) << 3); const int8_t * v4 = k_iq5nl + ((extra & 8) << 2);
                for (int j = 0; j < 16; ++j) {
                    yy[j]      = dl1 * v1[(qs[j] & 0xf) | (((qh[j] >> shift) & 1) << 4)];
                    yy[j + 16] = dl2 * v2[(qs[j + 16] & 0xf) | (((qh[j + 16] >> shift) & 1) << 4)];
                    yy[j + 32] = dl3 * v3[(qs[j] >> 4) | (((qh[j] >> shift) & 2) << 3)];
                    yy[j + 48] = dl4 * v4[(qs[j + 16] >> 4) | (((qh[j + 16] >> shift) & 2) << 3)];
                }
                yy += 64; qs += 32; extra >>= 4; shift += 2;
                if (shift == 8) { qh += 32; shift = 0; }
            }
        } break;
        case T_IQ4_KS: {  // iqk/iqk_quantize.cpp:4555-4580  row = {float d; blocks {u8 scales[8]; u8 qs[128]}}
            const uint8_t * sc = x; const uint8_t * qs = x + 8; float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                const float dl = row_scale * ((int)(sc[ib] & 254) - 127);
                const int8_t * v = k_iq4k + ((sc[ib] & 1) << 4);
                for (int j = 0; j < 16; ++j) { yy[j] = dl * v[qs[j] & 0xf]; yy[j + 16] = dl * v[qs[j] >> 4]; }
                yy += 32; qs += 16;
            }
        } break;
        case T_IQ5_KS: {  // iqk/iqk_quantize.cpp:4798-4822  row = {float d; blocks {u8 scales[8]; u8 qs[128]; u8 qh[32]}}
            const uint8_t * sc = x; const uint8_t * qs = x + 8; const uint8_t * qh = x + 136; float * yy = y;
            for (int ib64 = 0; ib64 < 4; ++ib64) {
                const float dl1 = row_scale * ((int)(sc[2 * ib64] & 254) - 127), dl2 = row_scale * ((int)(sc[2 * ib64 + 1] & 254) - 127);
                const int8_t * v1 = k_iq5nl + ((sc[2 * ib64] & 1) << 5); const int8_t * v2 = k_iq5nl + ((sc[2 * ib64 + 1] & 1) << 5);
                for (int j = 0; j < 32; ++j) {
                    yy[j]      = dl1 * v1[(qs[j] & 0xf) | (((qh[j] >> (2 * ib64)) & 1) << 4)];
                    yy[j + 32] = dl2 * v2[(qs[j] >> 4) | (((qh[j] >> (2 * ib64 + 1)) & 1) << 4)];
                }
                yy += 64; qs += 32;
            }
        } break;
        case T_MXFP4: {  // iqk/iqk_quantize.cpp:4224-4236  {u8 e; u8 qs[16]}; kvalues_mxfp4 ggml-common.h:2250; E8M0/2: ggml-impl.h:40-45
            static const int8_t kv[16] = {0, 1, 2, 3, 4, 6, 8, 12, 0, -1, -2, -3, -4, -6, -8, -12};
            const uint8_t e = x[0]; uint32_t u = e >= 2 ? (uint32_t)(e - 1) << 23 : (e == 0 ? 0x00200000u : 0x00400000u); float d; memcpy(&d, &u, 4);
            for (int j = 0; j < 16; ++j) { y[j] = d * kv[x[1 + j] & 0xf]; y[j + 16] = d * kv[x[1 + j] >> 4]; }
        } break;
        case T_IQ2_KS: {  // iqk/iqk_quantize.cpp:1877-1907  row = {half d; blocks {u16 extra; u8 scales[4]; u8 qs[64]}}
            static const int8_t v2[8] = {-31, -13, 1, 17, -26, -8, 6, 22};
            uint16_t extra = rd16(x); const uint8_t * sc = x + 2; const uint8_t * qs = x + 6; float * yy = y; int shift = 0;
            for (int ib64 = 0; ib64 < 4; ++ib64) {
                const float dl1 = row_scale * (((sc[ib64] & 0xf) | ((extra >> 4) & 0x10)) - 16), dl2 = row_scale * (((sc[ib64] >> 4) | ((extra >> 5) & 0x10)) - 16);
                const int8_t * va = extra & 1 ? v2 + 4 : v2; const int8_t * vb = extra & 2 ? v2 + 4 : v2; extra >>= 2;
                for (int j = 0; j < 32; ++j) { yy[j] = dl1 * va[(qs[j] >> (shift + 0)) & 3]; yy[j + 32] = dl2 * vb[(qs[j] >> (shift + 2)) & 3]; }
                yy += 64; shift += 4; if (shift == 8) { qs += 32; shift = 0; }
            }
        } break;
        case T_IQ3_KS: {  // iqk/iqk_quantize.cpp:2774-2803  row = {half d; blocks {u16 extra; u8 scales[4]; u8 qs[64]; u8 qh[32]}}
            static const int8_t v3[16] = {-63, -40, -23, -10, 1, 13, 28, 47, -59, -36, -19, -6, 5, 17, 32, 51};
            const uint16_t extra = rd16(x); const uint8_t * sc = x + 2; const uint8_t * qs = x + 6; const uint8_t * qh = x + 70; float dl[8]; float * yy = y;
            for (int j = 0; j < 4; ++j) {
                dl[j]     = row_scale * (((sc[j] & 0xf) | (((extra >> (j + 0)) & 1) << 4)) - 16);
                dl[j + 4] = row_scale * (((sc[j] >> 4) | (((extra >> (j + 4)) & 1) << 4)) - 16);
            }
            for (int i128 = 0; i128 < 2; ++i128) {
                for (int ib = 0; ib < 4; ++ib) {
                    const int8_t * v = v3 + (((extra >> (8 + 4 * i128 + ib)) & 1) << 3);
                    for (int j = 0; j < 32; ++j) yy[j] = dl[4 * i128 + ib] * v[((qs[j] >> (2 * ib)) & 3) | (((qh[j] >> (4 * i128 + ib)) & 1) << 2)];
                    yy += 32;
                }
                qs += 32;
            }
        } break;
        case T_IQ1_BN: {  // iqk/iqk_quantize.cpp:375-396 (+ half row scale, ggml.c:1273; to_float itself ignores it, SURVEY §8c pitfall 1)
                          // {u8 ql[12]; u8 extra}: 5 ternary digits per byte, digit j of byte b = ((v + (v >> 1)) >> 7) with v = u8(b * {81,27,9,3,1}[j]); w = row_scale * (digit - 1)
            static const uint8_t k_mult[5] = {81, 27, 9, 3, 1};
            const uint8_t extra = x[12]; const uint8_t * ql = x; float * yy = y;
            for (int i16 = 0; i16 < 4; ++i16) {
                for (int kk = 0; kk < 3; ++kk) for (int j = 0; j < 5; ++j) { const uint8_t v = (uint8_t)(ql[kk] * k_mult[j]); *yy++ = row_scale * (float)((int8_t)((v + (v >> 1)) >> 7) - 1); }
                ql += 3;
                const uint8_t v = (uint8_t)(extra * k_mult[i16]); *yy++ = row_scale * (float)((int8_t)((v + (v >> 1)) >> 7) - 1);
            }
        } break;
        case T_IQ4_KSS: {  // iqk/iqk_quantize.cpp:5161-5187  row = {float d; blocks {u32 qs[32]}}: per 32 weights eight u16, bit 0 of each = one bit of the
                           // scale byte, the other 15 bits hold 4 nibbles XOR-folded (v ^= v >> 1); codebook iq4k_values (+4 variant by the scale's bit 0)
            float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                uint16_t aux16[8]; int ls = 0;
                for (int kk = 0; kk < 8; ++kk) { const uint16_t q = rd16(x + 2 * (8 * ib + kk)); aux16[kk] = q & 0xfffe; aux16[kk] ^= (aux16[kk] >> 1); ls |= (q & 1) << kk; }
                const uint8_t * aux8 = (const uint8_t *)aux16;
                const int8_t * v = k_iq4k + ((ls & 1) << 4);
                const float dl = row_scale * ((ls & 254) - 127);
                for (int j = 0; j < 16; ++j) { yy[j] = dl * v[aux8[j] & 0xf]; yy[j + 16] = dl * v[aux8[j] >> 4]; }
                yy += 32;
            }
        } break;
        case T_IQ2_BN: {  // iqk/iqk_quantize.cpp:418-436 + row scale written at :227-229 (SURVEY §8c pitfall 1):
                          // w = row_scale * (q - 1), q = 2-bit field (j div 16) of byte (j mod 16)
            for (int j = 0; j < 64; ++j) y[j] = row_scale * (float)(((x[j % 16] >> (2 * (j / 16))) & 3) - 1);
        } break;
        case T_IQ1_S: {  // ggml-quants.c:3836-3859  {half d; u8 qs[32]; u16 qh[8]}: 11-bit grid index, 3-bit scale, sign of the +-1/8 shift
            const float d = h2f(rd16(x)); const uint8_t * qs = x + 2; float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                const uint16_t qh = rd16(x + 34 + 2 * ib);
                const float dl = d * (2 * ((qh >> 12) & 7) + 1), delta = qh & 0x8000 ? -0.125f : 0.125f;
                for (int l = 0; l < 4; ++l) {
                    const int8_t * grid = g_iq1s_grid + 8 * (qs[l] | (((qh >> 3 * l) & 7) << 8));
                    for (int j = 0; j < 8; ++j) yy[j] = dl * (grid[j] + delta);
                    yy += 8;
                }
                qs += 4;
            }
        } break;
        case T_IQ1_M: {  // ggml-quants.c:3861-3911  {u8 qs[32]; u8 qh[16]; u8 scales[8]}: the half super-scale is spread over the top nibbles of the 4 u16 scale words
            const uint8_t * qs = x; const uint8_t * qh = x + 32; uint16_t sc[4]; for (int i = 0; i < 4; ++i) sc[i] = rd16(x + 48 + 2 * i);
            const float d = h2f((uint16_t)((sc[0] >> 12) | ((sc[1] >> 8) & 0x00f0) | ((sc[2] >> 4) & 0x0f00) | (sc[3] & 0xf000)));
            float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                const float dl[2] = { d * (2 * ((sc[ib / 2] >> (6 * (ib % 2) + 0)) & 7) + 1), d * (2 * ((sc[ib / 2] >> (6 * (ib % 2) + 3)) & 7) + 1) };
                for (int l = 0; l < 4; ++l) {
                    const uint8_t h = qh[l / 2] >> (4 * (l % 2));
                    const int8_t * grid = g_iq1s_grid + 8 * (qs[l] | ((h & 7) << 8));
                    const float delta = h & 8 ? -0.125f : 0.125f;
                    for (int j = 0; j < 8; ++j) yy[j] = dl[l / 2] * (grid[j] + delta);
                    yy += 8;
                }
                qs += 4; qh += 2;
            }
        } break;
        case T_IQ2_KL: {  // iqk/iqk_quantize.cpp:2243-2275  row = {half d; blocks {u16 scales_h; u8 scales_l[4]; u8 qs[64]; u8 qh[16]}}: 5-bit index -> PAIR of values
            const uint16_t scales_h = rd16(x); const uint8_t * sl = x + 2; const uint8_t * qs = x + 6; const uint8_t * qh = x + 70; float * yy = y;
            for (int ib64 = 0; ib64 < 4; ++ib64) {
                const float dl1 = row_scale * (float)((int)(((sl[(2 * ib64 + 0) % 4] >> 4 * (ib64 / 2)) & 0xf) | (((scales_h >> (4 * ib64 + 0)) & 3) << 4)) - 32);
                const float dl2 = row_scale * (float)((int)(((sl[(2 * ib64 + 1) % 4] >> 4 * (ib64 / 2)) & 0xf) | (((scales_h >> (4 * ib64 + 2)) & 3) << 4)) - 32);
                for (int j = 0; j < 16; ++j) {
                    const int8_t * v1 = g_iq2kl + 2 * ((qs[j] & 0xf) | (((qh[j] >> (2 * ib64 + 0)) & 1) << 4));
                    const int8_t * v2 = g_iq2kl + 2 * ((qs[j] >> 4) | (((qh[j] >> (2 * ib64 + 1)) & 1) << 4));
                    yy[2 * j] = dl1 * v1[0]; yy[2 * j + 1] = dl1 * v1[1]; yy[2 * j + 32] = dl2 * v2[0]; yy[2 * j + 33] = dl2 * v2[1];
                }
                yy += 64; qs += 16;
            }
        } break;
        case T_IQ1_KT: {  // iqk/iqk_quantize.cpp:9470-9491  row = {float d; blocks {u8 sh[8]; u8 ql[32]; u8 qh[16]}}: 13-bit trellis index per 8 weights
            const uint8_t * sh = x; const uint8_t * ql = x + 8; const uint8_t * qh = x + 40; float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                const float sl = row_scale * k_iq4k[sh[ib] & 0xf];
                for (int ig = 0; ig < 4; ++ig) {
                    uint32_t idx = ql[ib * 4 + ig] | ((qh[(ib % 4) * 4 + ig] << (8 - 4 * (ib / 4))) & 0xf00);
                    idx |= ((uint32_t)sh[ib] << (8 - ig)) & 0x1000;
                    kt_values(idx, 4096, 8, sl, 0, yy); yy += 8;
                }
            }
        } break;
        case T_IQ2_KT: case T_IQ3_KT: {  // iqk/iqk_quantize.cpp:9751-9779 / :10021-10058  row = {float d; blocks {u8 scales[4]; u16 ql[32]; [u8 qh[32]]}}:
                                        // 16-bit trellis index per 8 weights; first 16 indices = weights 0..127, next 16 = 128..255; IQ3_KT: |value| and a sign bit plane
            const int q3 = type == T_IQ3_KT; const uint8_t * sc = x; const uint8_t * ql = x + 4; const uint8_t * qh = x + 68;
            for (int ib = 0; ib < 4; ++ib) {
                const float sl = row_scale * (q3 ? (float)(sc[ib] & 0xf) : (float)k_iq4k[sc[ib] & 0xf]), shh = row_scale * (q3 ? (float)(sc[ib] >> 4) : (float)k_iq4k[sc[ib] >> 4]);
                for (int ig = 0; ig < 4; ++ig) {
                    const int g = 4 * ib + ig; float * yl = y + 8 * g; float * yh = yl + 128;
                    kt_values(rd16(ql + 2 * g), 4096, 8, sl, q3, yl); kt_values(rd16(ql + 32 + 2 * g), 4096, 8, shh, q3, yh);
                    if (q3) for (int j = 0; j < 8; ++j) { if (qh[8 * ig + j] & (1 << ib)) yl[j] = -yl[j]; if (qh[8 * ig + j] & (16 << ib)) yh[j] = -yh[j]; }
                }
            }
        } break;
        case T_IQ4_KT: {  // iqk/iqk_quantize.cpp:10286-10313  row = {float d; blocks {u32 shb[8]; u8 ql[64]; u8 qh[32]}}: 15-bit index per 4 weights,
                          // 7-bit block scale - 64, bit 0 of shb selects the second half of the trellis (offset + 32768)
            const uint8_t * ql = x + 32; const uint8_t * qh = x + 96; float * yy = y;
            for (int ib = 0; ib < 8; ++ib) {
                uint32_t shb; memcpy(&shb, x + 4 * ib, 4);
                const uint32_t offset = shb & 1 ? 32768 + 4096 : 4096; const float sl = row_scale * (float)((int)((shb & 0xff) >> 1) - 64);
                for (int ig = 0; ig < 8; ++ig) {
                    const int jj = ib * 8 + ig;
                    const uint32_t idx = ql[jj] | ((qh[jj % 32] << (8 - 4 * (jj / 32))) & 0xf00) | (((shb >> (8 + 3 * ig)) & 7) << 12);
                    kt_values(idx, offset, 4, sl, 0, yy); yy += 4;
                }
            }
        } break;
        default: return -1;
        }
    }
    return 0;
}

// The _R4 repacks: 4 rows interleaved.  `g` points at a group of 4 rows (= 4 x row bytes, the four row scales first where the type has them),
// y receives the 4 rows [4][k].  Each function follows the reference's dequantize_row_<type>_r4 (iqk/iqk_quantize.cpp, lines cited per case).
static int dequant_group4(int type, const uint8_t * g, float * y, int64_t k) {
    int qk, bs, meta; if (geom(type, &qk, &bs, &meta) || k % qk) return -1;
    init_tables();
    const int64_t nb = k / qk;
    float d4[4] = {1.f, 1.f, 1.f, 1.f};
    if (meta == 4) memcpy(d4, g, 16);
    if (meta == 2) for (int r = 0; r < 4; ++r) d4[r] = h2f(rd16(g + 2 * r));
    const uint8_t * x = g + 4 * meta;
    for (int64_t ibl = 0; ibl < nb; ++ibl, x += 4 * bs) {
        for (int r = 0; r < 4; ++r) {
            float * yr = y + r * k + ibl * qk;
            switch (type) {
            case T_IQ1_S_R4: {  // :8195-8216  block {u8 qs[16]; u16 qh[4]} = 32 weights of 4 rows
                const uint16_t qh = rd16(x + 16 + 2 * r);
                const float shift = qh & 0x8000 ? -0.125f : 0.125f, dl = d4[r] * (2 * ((qh >> 12) & 7) + 1);
                for (int i = 0; i < 4; ++i) { const int8_t * grid = g_iq1s_grid + 8 * (x[4 * i + r] | (((qh >> 3 * i) & 7) << 8)); for (int j = 0; j < 8; ++j) yr[8 * i + j] = dl * (grid[j] + shift); }
            } break;
            case T_IQ1_M_R4: {  // :8336-8362  block {u8 qs[16]; u8 qh[8]; u8 scales[4]}
                const uint8_t * qs = x; const uint8_t * qh = x + 16; const uint8_t sc = x[24 + r];
                const float dl[2] = { d4[r] * (sc & 0xf), d4[r] * (sc >> 4) };
                for (int i = 0; i < 2; ++i) {
                    const uint8_t h = qh[4 * i + r];
                    const int8_t * g1 = g_iq1s_grid + 8 * (qs[8 * i + r] | ((h & 0x07) << 8)); const int8_t * g2 = g_iq1s_grid + 8 * (qs[8 * i + r + 4] | ((h & 0x70) << 4));
                    const float e1 = h & 0x08 ? -0.125f : 0.125f, e2 = h & 0x80 ? -0.125f : 0.125f;
                    for (int j = 0; j < 8; ++j) { yr[16 * i + j] = dl[i] * (g1[j] + e1); yr[16 * i + j + 8] = dl[i] * (g2[j] + e2); }
                }
            } break;
            case T_IQ2_K_R4: case T_IQ3_K_R4: {  // :7586-7616 / :7460-7494
                const int q3 = type == T_IQ3_K_R4;
                const float d = h2f(rd16(x + 2 * r)); const uint8_t * extra = x + 8;
                const uint8_t * scales_h = x + 16; const uint8_t * scales_l = x + (q3 ? 24 : 16); const uint8_t * ql = x + (q3 ? 56 : 48); const uint8_t * qh = x + 56 + 256;
                static const int8_t v2[8] = {-31, -13, 1, 17, -26, -8, 6, 22};                             // iq2nl_values (ggml-common.h:2212)
                static const int8_t v3[16] = {-63, -40, -23, -10, 1, 13, 28, 47, -59, -36, -19, -6, 5, 17, 32, 51};   // iq3nl_values (:2222)
                for (int ib = 0; ib < 8; ++ib) {
                    float dl[2];
                    for (int h = 0; h < 2; ++h) {
                        const int is = 8 * ib + r + 4 * h; const int nib = (scales_l[is % 32] >> 4 * (is / 32)) & 0xf;
                        dl[h] = q3 ? d * (2 * nib + 1) * ((scales_h[is % 8] >> (is / 8)) & 1 ? -1 : 1) : d * (nib - 8);
                    }
                    const int e1 = extra[r] & (1 << ib) ? 1 : 0, e2 = extra[r + 4] & (1 << ib) ? 1 : 0;
                    for (int i = 0; i < 4; ++i) for (int f = 0; f < 4; ++f) {
                        const int a = (ql[4 * r + i] >> 2 * f) & 3, b = (ql[4 * r + i + 16] >> 2 * f) & 3;
                        if (q3) {
                            const int ha = (qh[4 * r + i] >> f) & 1, hb = (qh[4 * r + i] >> (4 + f)) & 1;
                            yr[32 * ib + i + 4 * f] = dl[0] * v3[8 * e1 + (a | (ha << 2))]; yr[32 * ib + i + 4 * f + 16] = dl[1] * v3[8 * e2 + (b | (hb << 2))];
                        } else { yr[32 * ib + i + 4 * f] = dl[0] * v2[4 * e1 + a]; yr[32 * ib + i + 4 * f + 16] = dl[1] * v2[4 * e2 + b]; }
                    }
                    ql += 32; qh += 16;
                }
            } break;
            case T_IQ4_K_R4: case T_IQ5_K_R4: case T_IQ4_KS_R4: case T_IQ5_KS_R4: {  // :6700-6729 / :6838-6867 / :5879-5904 / :6946-6985
                const int ks = type == T_IQ4_KS_R4 || type == T_IQ5_KS_R4, q5 = type == T_IQ5_K_R4 || type == T_IQ5_KS_R4;
                const float d = ks ? d4[r] : h2f(rd16(x + 2 * r));
                const uint8_t * extra = x + 8; const uint8_t * scales_h = x + 16; const uint8_t * scales_l = x + 32;
                const uint8_t * qs = x + (ks ? 32 : 64); const uint8_t * qh = qs + 512;
                for (int ib = 0; ib < 8; ++ib) {
                    float dl[2]; int e[2];
                    if (ks) { const uint8_t sc = x[4 * ib + r]; dl[0] = dl[1] = d * (float)((int)(sc & 254) - 127); e[0] = e[1] = sc & 1; }
                    else for (int h = 0; h < 2; ++h) {
                        const int is = 8 * ib + r + 4 * h;
                        dl[h] = d * (float)((int)(((scales_l[is % 32] >> 4 * (is / 32)) & 0xf) | (((scales_h[is % 16] >> 2 * (is / 16)) & 3) << 4)) - 32);
                        e[h] = extra[r + 4 * h] & (1 << ib) ? 1 : 0;
                    }
                    for (int i = 0; i < 4; ++i) {
                        // byte c = qs[64 ib + 4 r + i + 16 c'] (c' = 0..3) holds weights {i, i+8}, {i+16, i+24}, {i+4, i+12}, {i+20, i+28} (low, high nibble)
                        static const int pos[4][2] = {{0, 8}, {16, 24}, {4, 12}, {20, 28}};
                        static const int bit[4][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}};
                        for (int c = 0; c < 4; ++c) for (int hn = 0; hn < 2; ++hn) {
                            const uint8_t byte = qs[64 * ib + 4 * r + i + 16 * c]; int q = hn ? byte >> 4 : byte & 0xf;
                            const int h = pos[c][hn] >= 16 ? 1 : 0;
                            if (q5) { q |= ((qh[16 * ib + 4 * r + i] >> bit[c][hn]) & 1) << 4; yr[32 * ib + i + pos[c][hn]] = dl[h] * k_iq5nl[32 * e[h] + q]; }
                            else yr[32 * ib + i + pos[c][hn]] = dl[h] * k_iq4k[16 * e[h] + q];
                        }
                    }
                }
            } break;
            default: return -1;
            }
        }
    }
    return 0;
}
// Dequantize rows i0 .. i0+3 (a whole group) or one row, whatever the type addresses: W = tensor base, rs = per-row bytes.
// Returns a pointer to row i inside `buf` ([4][k] floats), refilling buf when i enters a new group.
static const float * dequant_row_i(int type, const uint8_t * W, int64_t i, int64_t k, int64_t rs, float * buf, int * rc) {
    const int R = oracle_rows_interleaved(type);
    if (R == 1) { *rc = oracle_dequantize_row(type, W + i * rs, buf, k); return buf; }
    if (i % 4 == 0 || *rc == 1) *rc = dequant_group4(type, W + (i / 4) * 4 * rs, buf, k);
    return buf + (i % 4) * k;
}
// whole matrix -> f32 [m][k] (m must be a multiple of the interleave)
ORACLE_API int oracle_dequantize_matrix(int type, const uint8_t * W, float * y, int64_t m, int64_t k) {
    const int64_t rs = oracle_row_size(type, k); if (rs < 0) return -1;
    const int R = oracle_rows_interleaved(type); if (m % R) return -1;
    for (int64_t i = 0; i < m; i += R) { const int rc = R == 1 ? oracle_dequantize_row(type, W + i * rs, y + i * k, k) : dequant_group4(type, W + i * rs, y + i * k, k); if (rc) return rc; }
    return 0;
}

// quantize_q8_1 of the reference's CUDA path (ggml-cuda/quantize.cu:13-47): per 32 values
//   d = amax/127 ; q = amax == 0 ? 0 : roundf(x/d) ; stored d -> half, s = sum(x) -> half.
// q: int8 [n][k]; d_bits, s_bits: half bit patterns [n][k/32].  k must be a multiple of 32
// (the reference pads K to a multiple of 512 with zeros, which quantize to q = 0 and do not change the result).
ORACLE_API int oracle_quantize_q8_1(const float * x, int64_t n, int64_t k, int8_t * q, uint16_t * d_bits, uint16_t * s_bits) {
    if (k % 32) return -1;
    for (int64_t r = 0; r < n; ++r) for (int64_t b = 0; b < k / 32; ++b) {
        const float * xb = x + r * k + b * 32; float amax = 0.0f, sum = 0.0f;
        for (int j = 0; j < 32; ++j) { const float a = fabsf(xb[j]); if (a > amax) amax = a; }
        // warp_reduce_sum butterfly order (xor 16,8,4,2,1) to reproduce the f32 sum bit-for-bit
        float t[32]; for (int j = 0; j < 32; ++j) t[j] = xb[j];
        for (int w = 16; w > 0; w >>= 1) { float u[32]; for (int j = 0; j < 32; ++j) u[j] = t[j] + t[j ^ w]; memcpy(t, u, sizeof t); }
        sum = t[0];
        const float d = amax / 127;
        for (int j = 0; j < 32; ++j) q[r * k + b * 32 + j] = amax == 0.0f ? 0 : (int8_t)roundf(xb[j] / d);
        d_bits[r * (k / 32) + b] = f2h(d); s_bits[r * (k / 32) + b] = f2h(sum);
    }
    return 0;
}

// The product's variant of the activation quantiser (ik_llama_cpp_b200/csrc/b200q_decode.cu quantize_x_to_smem):
//   d = amax/127 (stored as half), inv = 1/d (correctly rounded), q = clamp(rint(x*inv), +-127) (round-half-even), i.e. one division + one reciprocal per block
// instead of the reference's roundf(x/d) per element.  Identical except at rounding ties (p ~ 1e-5 per element).
// Restated so that kernel-vs-oracle checks can be held to f32 summation-order accuracy.
ORACLE_API int oracle_quantize_q8_1_b200(const float * x, int64_t n, int64_t k, int8_t * q, uint16_t * d_bits) {
    if (k % 32) return -1;
    for (int64_t r = 0; r < n; ++r) for (int64_t b = 0; b < k / 32; ++b) {
        const float * xb = x + r * k + b * 32; float amax = 0.0f;
        for (int j = 0; j < 32; ++j) { const float a = fabsf(xb[j]); if (a > amax) amax = a; }
        const float d = amax / 127.0f; const float inv = d > 0.0f ? 1.0f / d : 0.0f;      // correctly rounded reciprocal == __frcp_rn
        for (int j = 0; j < 32; ++j) { const float p = xb[j] * inv; long qi = lrintf(p); if (qi > 127) qi = 127; if (qi < -127) qi = -127; q[r * k + b * 32 + j] = (int8_t)qi; }
        d_bits[r * (k / 32) + b] = f2h(d);
    }
    return 0;
}

// dst[n][m] (f32) = exact f64 dot of dequant(W) rows with x columns.  W: m wire rows, x: f32 [n][k].
ORACLE_API int oracle_mul_mat_exact(int type, const uint8_t * W, const float * x, float * dst, int64_t m, int64_t k, int64_t n) {
    const int64_t rs = oracle_row_size(type, k); if (rs < 0) return -1;
    float * buf = (float *)malloc(sizeof(float) * 4 * k);
    for (int64_t i = 0; i < m; ++i) {
        int rc = 0; const float * w = dequant_row_i(type, W, i, k, rs, buf, &rc);
        if (rc) { free(buf); return -1; }
        for (int64_t j = 0; j < n; ++j) { double acc = 0; const float * xr = x + j * k; for (int64_t l = 0; l < k; ++l) acc += (double)w[l] * xr[l]; dst[j * m + i] = (float)acc; }
    }
    free(buf); return 0;
}

// dst[n][m] = the value the reference's MMVQ kernels compute up to f32 summation order (see header).
static int mul_mat_q8_impl(int type, const uint8_t * W, const float * x, float * dst, int64_t m, int64_t k, int64_t n, int variant);
ORACLE_API int oracle_mul_mat_q8_1(int type, const uint8_t * W, const float * x, float * dst, int64_t m, int64_t k, int64_t n) { return mul_mat_q8_impl(type, W, x, dst, m, k, n, 0); }
// same with the product's quantiser variant (see oracle_quantize_q8_1_b200)
ORACLE_API int oracle_mul_mat_q8_1_b200(int type, const uint8_t * W, const float * x, float * dst, int64_t m, int64_t k, int64_t n) { return mul_mat_q8_impl(type, W, x, dst, m, k, n, 1); }
static int mul_mat_q8_impl(int type, const uint8_t * W, const float * x, float * dst, int64_t m, int64_t k, int64_t n, int variant) {
    const int64_t rs = oracle_row_size(type, k); if (rs < 0 || k % 32) return -1;
    int8_t * q = (int8_t *)malloc((size_t)n * k); uint16_t * db = (uint16_t *)malloc(sizeof(uint16_t) * n * (k / 32)); uint16_t * sb = (uint16_t *)malloc(sizeof(uint16_t) * n * (k / 32));
    float * xq = (float *)malloc(sizeof(float) * n * k); float * buf = (float *)malloc(sizeof(float) * 4 * k);
    if (variant) oracle_quantize_q8_1_b200(x, n, k, q, db); else oracle_quantize_q8_1(x, n, k, q, db, sb);
    for (int64_t j = 0; j < n; ++j) for (int64_t l = 0; l < k; ++l) xq[j * k + l] = h2f(db[j * (k / 32) + l / 32]) * q[j * k + l];
    int rc = 0;
    for (int64_t i = 0; i < m && !rc; ++i) {
        const float * w = dequant_row_i(type, W, i, k, rs, buf, &rc);
        if (rc) break;
        for (int64_t j = 0; j < n; ++j) { double acc = 0; const float * xr = xq + j * k; for (int64_t l = 0; l < k; ++l) acc += (double)w[l] * xr[l]; dst[j * m + i] = (float)acc; }
    }
    free(q); free(db); free(sb); free(xq); free(buf); return rc;
}
