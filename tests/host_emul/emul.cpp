// tests/host_emul/emul.cpp — compiles the PRODUCT's __host__ __device__ format code (b200q_types.cuh)
// with g++ so that the exact repack / decode bit manipulation the CUDA kernels execute can be checked
// on a CPU-only box against the oracle.  Test infrastructure only; nothing here ships.
#include "../../ik_llama_cpp_b200/csrc/b200q_types.cuh"
#include "../../ik_llama_cpp_b200/csrc/b200q_wire.cuh"
#include <vector>
#include <cmath>
#define API extern "C" __attribute__((visibility("default")))

API long emul_layout_bytes(int type, long M, long K) { b200q_layout L; if (b200q_make_layout(type, M, K, &L)) return -1; return (long)L.total_bytes; }
API long emul_wire_row_size(int type, long K) { b200q_layout L; if (b200q_make_layout(type, 4, K, &L)) return -1; return (long)b200q_wire_row_size(L); }

API int emul_repack(int type, const uint8_t * wire, uint8_t * planes, long M, long K, int inverse) {
    b200q_layout L; if (b200q_make_layout(type, M, K, &L)) return -1;
    const long rs = b200q_wire_row_size(L);
    for (long r = 0; r < M; ++r) {
        const uint8_t * row = wire + r * rs;
        b200q_repack_row_meta(L, row, planes, r, inverse != 0);
        for (long b = 0; b < L.nb; ++b) b200q_repack_block(L, row + L.row_meta + b * L.wire_block, planes, r, b, inverse != 0);
    }
    return 0;
}

template <int T> static void deq(const uint8_t * planes, const b200q_layout & L, float * out) {
    for (long r = 0; r < L.M; ++r) for (long it = 0; it < L.K / 32; ++it) {
        b200q_item I; memset(&I, 0, sizeof I); b200q_canon C; memset(&C, 0, sizeof C);
        b200q_load_item<T>(I, planes, L, r, it); b200q_decode_item<T>(I, it, C);
        b200q_canon_to_float<b200q_traits<T>::HAS_B>(C, out + r * L.K + it * 32);
    }
}
// dst[n][m]: emulation of the mat-vec arithmetic: per item  acc += d8 * (dl0*sumi0 + dl1*sumi1) - d8*(ml0*is0 + ml1*is1)
template <int T> static void mmv(const uint8_t * planes, const b200q_layout & L, const int8_t * xq, const float * xd, const int * xis, long n, float * dst) {
    for (long r = 0; r < L.M; ++r) for (long j = 0; j < n; ++j) {
        float acc = 0.f;
        for (long it = 0; it < L.K / 32; ++it) {
            b200q_item I; memset(&I, 0, sizeof I); b200q_canon C; memset(&C, 0, sizeof C);
            b200q_load_item<T>(I, planes, L, r, it); b200q_decode_item<T>(I, it, C);
            const int * x = (const int *)(xq + j * L.K + it * 32);
            int s0 = 0, s1 = 0;
            for (int w = 0; w < 4; ++w) { s0 = b200q_dp4a(C.va[w], x[w], s0); s1 = b200q_dp4a(C.va[4 + w], x[4 + w], s1); }
            if (b200q_traits<T>::HAS_B) for (int w = 0; w < 4; ++w) { s0 = b200q_dp4a(C.vb[w], x[w], s0); s1 = b200q_dp4a(C.vb[4 + w], x[4 + w], s1); }
            const int is = xis[j * (L.K / 32) + it]; const int is0 = (int)(int16_t)(is & 0xFFFF), is1 = (int)(int16_t)(is >> 16);
            acc += xd[j * (L.K / 32) + it] * ((C.dl[0] * (float)s0 + C.dl[1] * (float)s1) - (C.ml[0] * (float)is0 + C.ml[1] * (float)is1));
        }
        dst[j * L.M + r] = acc;
    }
}
#define DISPATCH(F, ...) switch (type) { \
    case B200Q_TYPE_IQ4_NL: F<B200Q_TYPE_IQ4_NL>(__VA_ARGS__); break; case B200Q_TYPE_Q4_0: F<B200Q_TYPE_Q4_0>(__VA_ARGS__); break; \
    case B200Q_TYPE_Q8_0: F<B200Q_TYPE_Q8_0>(__VA_ARGS__); break; case B200Q_TYPE_Q4_K: F<B200Q_TYPE_Q4_K>(__VA_ARGS__); break; \
    case B200Q_TYPE_Q5_K: F<B200Q_TYPE_Q5_K>(__VA_ARGS__); break; case B200Q_TYPE_Q6_K: F<B200Q_TYPE_Q6_K>(__VA_ARGS__); break; \
    case B200Q_TYPE_IQ4_XS: F<B200Q_TYPE_IQ4_XS>(__VA_ARGS__); break; case B200Q_TYPE_IQ4_K: F<B200Q_TYPE_IQ4_K>(__VA_ARGS__); break; \
    case B200Q_TYPE_IQ4_KS: F<B200Q_TYPE_IQ4_KS>(__VA_ARGS__); break; case B200Q_TYPE_IQ5_K: F<B200Q_TYPE_IQ5_K>(__VA_ARGS__); break; \
    case B200Q_TYPE_IQ2_BN: F<B200Q_TYPE_IQ2_BN>(__VA_ARGS__); break; \
    case B200Q_TYPE_Q4_1: F<B200Q_TYPE_Q4_1>(__VA_ARGS__); break; case B200Q_TYPE_Q5_0: F<B200Q_TYPE_Q5_0>(__VA_ARGS__); break; \
    case B200Q_TYPE_Q5_1: F<B200Q_TYPE_Q5_1>(__VA_ARGS__); break; case B200Q_TYPE_Q6_0: F<B200Q_TYPE_Q6_0>(__VA_ARGS__); break; \
    case B200Q_TYPE_Q2_K: F<B200Q_TYPE_Q2_K>(__VA_ARGS__); break; case B200Q_TYPE_Q3_K: F<B200Q_TYPE_Q3_K>(__VA_ARGS__); break; \
    case B200Q_TYPE_IQ2_K: F<B200Q_TYPE_IQ2_K>(__VA_ARGS__); break; case B200Q_TYPE_IQ3_K: F<B200Q_TYPE_IQ3_K>(__VA_ARGS__); break; \
    case B200Q_TYPE_MXFP4: F<B200Q_TYPE_MXFP4>(__VA_ARGS__); break; case B200Q_TYPE_IQ5_KS: F<B200Q_TYPE_IQ5_KS>(__VA_ARGS__); break; \
    case B200Q_TYPE_IQ2_KS: F<B200Q_TYPE_IQ2_KS>(__VA_ARGS__); break; case B200Q_TYPE_IQ3_KS: F<B200Q_TYPE_IQ3_KS>(__VA_ARGS__); break; default: return -1; }

API int emul_dequant(int type, const uint8_t * planes, long M, long K, float * out) {
    b200q_layout L; if (b200q_make_layout(type, M, K, &L)) return -1;
    DISPATCH(deq, planes, L, out); return 0;
}
API int emul_mul_mat_vec(int type, const uint8_t * planes, long M, long K, const int8_t * xq, const float * xd, const int * xis, long n, float * dst) {
    b200q_layout L; if (b200q_make_layout(type, M, K, &L)) return -1;
    DISPATCH(mmv, planes, L, xq, xd, xis, n, dst); return 0;
}

// ---- wire-layout types (b200q_wire.cuh): the product's decode32 of every 32-weight group, straight from the GGUF bytes ----
template <int T> static void wire_deq(const uint8_t * W, long M, long K, float * out) {
    for (long r = 0; r < M; ++r) for (long it = 0; it < K / 32; ++it) b200q_wire_decode32<T>(W, K, r, it, out + r * K + it * 32);
}
API int emul_wire_dequant(int type, const uint8_t * W, long M, long K, float * out) {
    b200q_layout L; if (b200q_make_layout(type, M, K, &L) || !L.wire) return -1;
    switch (type) {
#define X(T) case T: wire_deq<T>(W, M, K, out); break;
        B200Q_FOR_WIRE_TYPES(X)
#undef X
        default: return -1;
    }
    return 0;
}
