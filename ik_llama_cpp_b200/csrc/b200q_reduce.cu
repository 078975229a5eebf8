// b200q_reduce.cu — GGML_OP_REDUCE (sum) over NVLink/NVSwitch for the row-parallel mat-muls of "split mode graph".
// Replaces ggml_cuda_op_reduce (ggml/src/ggml-cuda/reduce.cu:125-598: ncclAllReduce / copy-engine ring / k_reduce_add_T).
//
// One kernel per all-reduce, one process per GPU, buffers in symmetric memory with an NVLS multicast mapping:
//   (1) zero the dirty part of the OTHER parity buffer (it is used by the next all-reduce; peers may only start adding to it after they have
//       seen this rank's flag increment below, which is ordered after the zeroing);
//   (2) multimem.red.add.f32 of the local partial into the multicast address: the switch adds it into EVERY rank's copy;
//   (3) last CTA: multimem.red.add.u32 on the multicast flag (release.sys) -> every rank's flag += 1;
//   (4) every CTA spins (ld.acquire.sys) until the local flag reaches world * use_count, then copies its slice of the
//       local (now fully reduced) buffer to `out`.
// tg: n = n_embd floats (16 KiB) -> 1 CTA, pure latency (~2 NVLink hops); pp512: 8 MiB -> up to 148 CTAs.
// f32 adds are performed by the switch in arrival order (like NCCL's NVLS algorithm): run-to-run LSB differences are possible.
#include "b200q_internal.h"
#include <cuda_runtime.h>

namespace {
__device__ __forceinline__ void mm_red_add_f32x4(float * mc, float4 v) {
    asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mm_red_add_f32(float * mc, float v) {
    asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}
__device__ __forceinline__ void mm_red_add_u32_release(uint32_t * mc, uint32_t v) {
    asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t * p) {
    uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// `seq` (device-local counter of reduces issued on this communicator) selects the parity buffer and the flag target inside the
// kernel, so the launch parameters are constant and the kernel can be replayed from a CUDA graph.
__global__ void __launch_bounds__(512) k_allreduce_nvls(const float * __restrict__ in, float * __restrict__ out, int64_t n,
                                                         float * mc_base, float * local_base, int64_t stride,
                                                         uint32_t * mc_flag, const uint32_t * local_flag, uint32_t world,
                                                         uint32_t * seq, uint32_t * cta_counter) {
    const uint32_t s = *reinterpret_cast<volatile uint32_t *>(seq);      // read before this CTA's counter increment (see below)
    // seq[1 + p] = number of floats of parity buffer p that may be non-zero (left there by its last use): reduces of different
    // lengths share the buffers (tg: n_embd floats, pp512: 512 x n_embd), so the zeroing covers what was actually dirtied
    const int64_t n_dirty = (int64_t)reinterpret_cast<volatile uint32_t *>(seq)[1 + ((s & 1) ^ 1)];
    const uint32_t target = world * (s + 1);
    float * mc_buf = mc_base + (int64_t)(s & 1) * stride;
    const float * local_buf = local_base + (int64_t)(s & 1) * stride;
    float * local_zero = local_base + (int64_t)((s & 1) ^ 1) * stride;
    const int64_t n4 = n / 4;
    const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * per, i1 = min(n4, i0 + per);
    // (1)
    {
        const int64_t d4 = (n_dirty + 3) / 4, dper = (d4 + gridDim.x - 1) / gridDim.x;
        const int64_t z0 = (int64_t)blockIdx.x * dper, z1 = min(d4, z0 + dper);
        for (int64_t i = z0 + threadIdx.x; i < z1; i += blockDim.x) reinterpret_cast<float4 *>(local_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // (2)
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x)
        mm_red_add_f32x4(mc_buf + 4 * i, __ldg(reinterpret_cast<const float4 *>(in) + i));
    if (blockIdx.x == gridDim.x - 1)
        for (int64_t i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) mm_red_add_f32(mc_buf + i, in[i]);
    __threadfence_system();
    __syncthreads();
    // (3)
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(cta_counter, 1u);
        if (done == gridDim.x - 1) {                 // every CTA has read `seq` (it does so before its atomicAdd)
            *cta_counter = 0;
            reinterpret_cast<volatile uint32_t *>(seq)[1 + ((s & 1) ^ 1)] = 0; reinterpret_cast<volatile uint32_t *>(seq)[1 + (s & 1)] = (uint32_t)n;
            *reinterpret_cast<volatile uint32_t *>(seq) = s + 1; __threadfence();
            mm_red_add_u32_release(mc_flag, 1u);
        }
        // (4)
        while ((int32_t)(ld_acquire_sys(local_flag) - target) < 0) { __nanosleep(32); }
    }
    __syncthreads();
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x)
        reinterpret_cast<float4 *>(out)[i] = __ldcv(reinterpret_cast<const float4 *>(local_buf) + i);
    if (blockIdx.x == gridDim.x - 1)
        for (int64_t i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) out[i] = __ldcv(local_buf + i);
}
}  // namespace

int b200q_launch_allreduce_nvls(const float * in, float * out, int64_t n, void * mc_base, void * local_base, int64_t stride,
                                void * mc_flag, const void * local_flag, uint32_t world, void * seq, void * cta_counter, int sm_count, cudaStream_t st) {
    if (n <= 0 || n > stride || ((uintptr_t)in & 15) || ((uintptr_t)out & 15) || ((uintptr_t)mc_base & 15) || (stride & 3)) return -2;
    int64_t grid = (n + 16383) / 16384;            // >= 64 KiB per CTA
    if (grid > sm_count) grid = sm_count;          // all CTAs must be co-resident: they spin on the flag
    if (grid < 1) grid = 1;
    k_allreduce_nvls<<<(unsigned)grid, 512, 0, st>>>(in, out, n, (float *)mc_base, (float *)local_base, stride,
                                                    (uint32_t *)mc_flag, (const uint32_t *)local_flag, world, (uint32_t *)seq, (uint32_t *)cta_counter);
    return (int)cudaGetLastError();
}
