// b200q_decode.cu — HBM-bound kernels of the hot path for sm_100a:
//   * k_repack / k_unrepack   wire (GGUF) blocks <-> plane layout (b200q_types.cuh), run once per tensor upload/download
//   * k_mmvq                  decode mat-vec  dst[n][m] = sum_k W[m][k] x[n][k],  n <= 8   (replaces the reference's
//                             quantize_q8_1 + mul_mat_vec_q / iqk_mul_mat_vec_q / fused_mul_mat_vec_q:
//                             ggml/src/ggml-cuda/quantize.cu:13-47, mmvq-templates.cuh:68-330, iqk_mmvq_templates.cuh:21-300)
//   * k_dequant_bf16          planes -> bf16 [M][K] (generic feeder of the tcgen05 GEMM for types without a fused prefill kernel)
//
// Decode design (one launch per GGML_OP_MUL_MAT / FUSED_UP_GATE node, no tensor cores):
//   - prologue: every CTA quantises the activation column(s) to q8_1 semantics straight into shared memory
//     (int8 values in natural k order, d rounded to half like block_q8_1.ds.x, integer sums per 16) — no separate
//     quantize launch, no q8_1 round trip through HBM;
//   - main loop: one warp per output row; lane l owns items l, l+32, ... (item = 32 weights = one 16-byte LDG.128 of the
//     low-bit plane, perfectly coalesced: 512 contiguous bytes per warp-load), UNROLL independent loads in flight
//     before the first use; PRMT-LUT / mask decode into int8x4 lanes; dp4a against the smem activations;
//   - epilogue: warp-shuffle reduce, optional bias, optional act(gate)*up fusion, one 4-byte store per row.
#include "b200q_types.cuh"
#include "b200q_internal.h"
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

// ------------------------------------------------------------------------------------------------
// repack
// ------------------------------------------------------------------------------------------------
__global__ void k_repack(const uint8_t * __restrict__ wire, uint8_t * __restrict__ planes, b200q_layout L, int inverse) {
    const int64_t rs = (int64_t)L.row_meta + L.nb * L.wire_block;
    const int64_t total = L.M * L.nb;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / L.nb, blk = i % L.nb;
        const uint8_t * w = wire + row * rs + L.row_meta + blk * L.wire_block;
        b200q_repack_block(L, w, planes, row, blk, inverse != 0);
        if (blk == 0) b200q_repack_row_meta(L, wire + row * rs, planes, row, inverse != 0);
    }
}

// ------------------------------------------------------------------------------------------------
// dequantise planes -> bf16 [M][K]
// ------------------------------------------------------------------------------------------------
template <int TYPE>
__global__ void k_dequant_bf16(const uint8_t * __restrict__ W, b200q_layout L, __nv_bfloat16 * __restrict__ out) {
    const int64_t n32 = L.K / 32, total = L.M * n32;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / n32, it = i % n32;
        b200q_item I; b200q_canon C;
        b200q_load_item<TYPE>(I, W, L, row, it);
        b200q_decode_item<TYPE>(I, it, C);
        float f[32];
        b200q_canon_to_float<b200q_traits<TYPE>::HAS_B>(C, f);
        uint4 * o = reinterpret_cast<uint4 *>(out + row * L.K + it * 32);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(f[8 * v + 0], f[8 * v + 1]), p1 = __floats2bfloat162_rn(f[8 * v + 2], f[8 * v + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(f[8 * v + 4], f[8 * v + 5]), p3 = __floats2bfloat162_rn(f[8 * v + 6], f[8 * v + 7]);
            uint4 u; u.x = *reinterpret_cast<uint32_t *>(&p0); u.y = *reinterpret_cast<uint32_t *>(&p1); u.z = *reinterpret_cast<uint32_t *>(&p2); u.w = *reinterpret_cast<uint32_t *>(&p3);
            o[v] = u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// decode mat-vec
// ------------------------------------------------------------------------------------------------
struct mmvq_seg {               // one weight tensor of a multi-tensor launch (Q,K,V share the activation)
    const uint8_t * W;          // plane base
    const uint8_t * W2;         // second tensor (gate) for the fused up/gate mode, else nullptr
    float *         dst;        // [ncols][M] f32 (ggml: dst[j*M + i])
    const float *   bias;       // optional [M]
    int64_t         M;
    int64_t         row0;       // first global row index of this segment
};
struct mmvq_args {
    mmvq_seg     seg[B200Q_MAX_SEGS];
    int          n_seg;
    int64_t      M_total;
    int64_t      K;
    const float * x;            // [ncols][K] f32, row stride x_stride floats
    int64_t      x_stride;
    int          act;           // B200Q_ACT_* for the up/gate mode
    float        limit;         // clamp for swiglu variants (0 = none)
    b200q_layout L;             // geometry (M of the layout is per-segment; only offsets that do not depend on M are used here)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float act_apply(int act, float g) {
    switch (act) {
        case B200Q_ACT_SILU: return g / (1.0f + expf(-g));
        case B200Q_ACT_GELU: { const float c = 0.79788456080286535587989211986876f, a = 0.044715f; return 0.5f * g * (1.0f + tanhf(c * g * (1.0f + a * g * g))); }
        case B200Q_ACT_RELU: return fmaxf(g, 0.0f);
        default: return g;
    }
}

// Quantise ncols activation columns into shared memory (q8_1 semantics of ggml-cuda/quantize.cu:13-47):
//   d = amax/127 ; q = amax == 0 ? 0 : roundf(x/d) ; d kept as float(half(d)) ; isum = packed int16 sums of q over each 16.
template <int NCOLS>
__device__ __forceinline__ void quantize_x_to_smem(const float * __restrict__ x, int64_t x_stride, int64_t K,
                                                   int8_t * sq, float * sd, int * sis) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int n32 = (int)(K / 32);
    for (int b = warp; b < n32 * NCOLS; b += nwarps) {
        const int col = b / n32, blk = b % n32;
        const float v = __ldg(x + col * x_stride + (int64_t)blk * 32 + lane);
        float amax = fabsf(v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const float d = amax / 127.0f;
        const int q = amax == 0.0f ? 0 : (int)roundf(__fdiv_rn(v, d));
        int s = q;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);   // sum within each 16-lane half
        const int s_hi = __shfl_sync(0xffffffffu, s, 16);
        sq[(size_t)col * K + blk * 32 + lane] = (int8_t)q;
        if (lane == 0) {
            sd[col * n32 + blk]  = __half2float(__float2half_rn(d));
            sis[col * n32 + blk] = (s & 0xFFFF) | (s_hi << 16);
        }
    }
}

template <int TYPE, int NCOLS>
__device__ __forceinline__ void item_dot(const b200q_canon & C, const int8_t * sq, const float * sd, const int * sis,
                                         int64_t K, int n32, int it, float acc[NCOLS]) {
    constexpr bool HAS_B = b200q_traits<TYPE>::HAS_B;
#pragma unroll
    for (int c = 0; c < NCOLS; ++c) {
        const int4 * xp = reinterpret_cast<const int4 *>(sq + (size_t)c * K + (size_t)it * 32);
        const int4 x0 = xp[0], x1 = xp[1];
        int s0 = 0, s1 = 0;
        s0 = b200q_dp4a(C.va[0], x0.x, s0); s0 = b200q_dp4a(C.va[1], x0.y, s0); s0 = b200q_dp4a(C.va[2], x0.z, s0); s0 = b200q_dp4a(C.va[3], x0.w, s0);
        s1 = b200q_dp4a(C.va[4], x1.x, s1); s1 = b200q_dp4a(C.va[5], x1.y, s1); s1 = b200q_dp4a(C.va[6], x1.z, s1); s1 = b200q_dp4a(C.va[7], x1.w, s1);
        if (HAS_B) {
            s0 = b200q_dp4a(C.vb[0], x0.x, s0); s0 = b200q_dp4a(C.vb[1], x0.y, s0); s0 = b200q_dp4a(C.vb[2], x0.z, s0); s0 = b200q_dp4a(C.vb[3], x0.w, s0);
            s1 = b200q_dp4a(C.vb[4], x1.x, s1); s1 = b200q_dp4a(C.vb[5], x1.y, s1); s1 = b200q_dp4a(C.vb[6], x1.z, s1); s1 = b200q_dp4a(C.vb[7], x1.w, s1);
        }
        const float d8 = sd[c * n32 + it];
        float t = C.dl[0] * (float)s0 + C.dl[1] * (float)s1;
        if (b200q_mmvq_has_ml(TYPE)) {
            const int is = sis[c * n32 + it];
            t -= C.ml[0] * (float)(int)(short)(is & 0xFFFF) + C.ml[1] * (float)(is >> 16);
        }
        acc[c] = fmaf(d8, t, acc[c]);
    }
}

template <int TYPE, int NCOLS, bool UPGATE>
__global__ void __launch_bounds__(512, (NCOLS == 1 ? 2 : 1)) k_mmvq(const mmvq_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int64_t K = a.K; const int n32 = (int)(K / 32);
    int8_t * sq = reinterpret_cast<int8_t *>(smem_raw);
    float *  sd = reinterpret_cast<float *>(smem_raw + (size_t)NCOLS * K);
    int *    sis = reinterpret_cast<int *>(sd + NCOLS * n32);

    quantize_x_to_smem<NCOLS>(a.x, a.x_stride, K, sq, sd, sis);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, tw = (int64_t)gridDim.x * nwarps;
    constexpr int U = UPGATE ? 2 : 4;

    for (int64_t grow = gw; grow < a.M_total; grow += tw) {
        int s = 0;
#pragma unroll
        for (int i = 1; i < B200Q_MAX_SEGS; ++i) if (i < a.n_seg && grow >= a.seg[i].row0) s = i;
        const mmvq_seg & sg = a.seg[s];
        const int64_t row = grow - sg.row0;
        b200q_layout L = a.L; L.M = sg.M;
        // plane offsets depend on M: recompute (cheap integer math; identical to b200q_make_layout)
        { int64_t off = 0;
#pragma unroll
          for (int p = 0; p < B200Q_MAX_PLANES; ++p) if (p < L.n_planes) { L.plane_off[p] = off; const int64_t n = L.plane_per_row[p] ? L.M : L.M * L.nb; off = b200q_align_up(off + n * L.plane_bytes[p], 256); } }

        float acc[NCOLS], acc2[NCOLS];
#pragma unroll
        for (int c = 0; c < NCOLS; ++c) { acc[c] = 0.0f; acc2[c] = 0.0f; }

        for (int it0 = lane; it0 < n32; it0 += 32 * U) {
            b200q_item I[U], J[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int it = it0 + 32 * u;
                if (it < n32) { b200q_load_item<TYPE>(I[u], sg.W, L, row, it); if (UPGATE) b200q_load_item<TYPE>(J[u], sg.W2, L, row, it); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int it = it0 + 32 * u;
                if (it < n32) {
                    b200q_canon C;
                    b200q_decode_item<TYPE>(I[u], it, C);
                    item_dot<TYPE, NCOLS>(C, sq, sd, sis, K, n32, it, acc);
                    if (UPGATE) { b200q_decode_item<TYPE>(J[u], it, C); item_dot<TYPE, NCOLS>(C, sq, sd, sis, K, n32, it, acc2); }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NCOLS; ++c) {
            float v = warp_sum(acc[c]);
            if (UPGATE) {
                float g = warp_sum(acc2[c]);      // acc = up . x, acc2 = gate . x
                if (a.limit > 0.0f) { g = fminf(g, a.limit); v = fminf(fmaxf(v, -a.limit), a.limit); }
                v = act_apply(a.act, g) * v;
            } else if (sg.bias) v += sg.bias[row];
            if (lane == 0) sg.dst[(int64_t)c * sg.M + row] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
#define B200Q_FOR_TYPES(X) X(B200Q_TYPE_IQ4_NL) X(B200Q_TYPE_Q4_0) X(B200Q_TYPE_Q8_0) X(B200Q_TYPE_Q4_K) X(B200Q_TYPE_Q5_K) \
    X(B200Q_TYPE_Q6_K) X(B200Q_TYPE_IQ4_XS) X(B200Q_TYPE_IQ4_K) X(B200Q_TYPE_IQ4_KS) X(B200Q_TYPE_IQ5_K) X(B200Q_TYPE_IQ2_BN)

int b200q_launch_repack(const void * wire, void * planes, const b200q_layout & L, int inverse, cudaStream_t st) {
    const int64_t total = L.M * L.nb;
    const int bs = 128; const int64_t nb = (total + bs - 1) / bs;
    k_repack<<<(unsigned)(nb > 65535 * 16 ? 65535 * 16 : (nb < 1 ? 1 : nb)), bs, 0, st>>>((const uint8_t *)wire, (uint8_t *)planes, L, inverse);
    return (int)cudaGetLastError();
}

int b200q_launch_dequant_bf16(const void * W, const b200q_layout & L, void * out, cudaStream_t st) {
    const int64_t total = L.M * (L.K / 32);
    const int bs = 256; int64_t nb = (total + bs - 1) / bs; if (nb > 148 * 64) nb = 148 * 64; if (nb < 1) nb = 1;
    switch (L.type) {
#define X(T) case T: k_dequant_bf16<T><<<(unsigned)nb, bs, 0, st>>>((const uint8_t *)W, L, (__nv_bfloat16 *)out); break;
        B200Q_FOR_TYPES(X)
#undef X
        default: return -1;
    }
    return (int)cudaGetLastError();
}

template <int TYPE, int NCOLS, bool UPGATE>
static int launch_mmvq_t(const mmvq_args & a, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)NCOLS * a.K + (size_t)NCOLS * (a.K / 32) * 8;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        if (cudaFuncSetAttribute(k_mmvq<TYPE, NCOLS, UPGATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
        configured = smem;
    }
    // one warp per row; size the CTA so that the grid covers the SMs about twice (prologue cost is per CTA)
    int nwarps = 16;
    while (nwarps > 4 && a.M_total < (int64_t)sm_count * 2 * nwarps) nwarps >>= 1;
    int64_t grid = (a.M_total + nwarps - 1) / nwarps;
    const int64_t max_grid = (int64_t)sm_count * (smem > 100 * 1024 ? 1 : 2) * (16 / nwarps);
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    k_mmvq<TYPE, NCOLS, UPGATE><<<(unsigned)grid, nwarps * 32, smem, st>>>(a);
    return (int)cudaGetLastError();
}

template <int TYPE>
static int launch_mmvq_type(const mmvq_args & a, int ncols, bool upgate, int sm_count, cudaStream_t st) {
#define CASE(N) case N: return upgate ? launch_mmvq_t<TYPE, N, true>(a, sm_count, st) : launch_mmvq_t<TYPE, N, false>(a, sm_count, st);
    switch (ncols) { CASE(1) CASE(2) CASE(4) CASE(8) default: return -2; }
#undef CASE
}

int b200q_launch_mmvq(const b200q_mmvq_desc & d, cudaStream_t st) {
    mmvq_args a; memset(&a, 0, sizeof a);
    if (d.n_seg < 1 || d.n_seg > B200Q_MAX_SEGS || d.ncols < 1 || d.ncols > 8) return -2;
    if (b200q_make_layout(d.type, d.seg[0].M, d.K, &a.L)) return -1;
    int64_t r0 = 0;
    for (int i = 0; i < d.n_seg; ++i) {
        a.seg[i].W = (const uint8_t *)d.seg[i].W; a.seg[i].W2 = (const uint8_t *)d.seg[i].W2; a.seg[i].dst = d.seg[i].dst;
        a.seg[i].bias = d.seg[i].bias; a.seg[i].M = d.seg[i].M; a.seg[i].row0 = r0; r0 += d.seg[i].M;
    }
    a.n_seg = d.n_seg; a.M_total = r0; a.K = d.K; a.x = d.x; a.x_stride = d.x_stride ? d.x_stride : d.K; a.act = d.act; a.limit = d.limit;
    const bool upgate = d.seg[0].W2 != nullptr;
    switch (d.type) {
#define X(T) case T: return launch_mmvq_type<T>(a, d.ncols, upgate, d.sm_count, st);
        B200Q_FOR_TYPES(X)
#undef X
        default: return -1;
    }
}
