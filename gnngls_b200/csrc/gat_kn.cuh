// Shared device helpers and argument block of the two K_n aggregate kernels (gat_kn.cu: exact sorted-prefix kernel;
// gat_kn_tc.cu: tcgen05 indicator-matrix kernel).
#pragma once
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include "common.h"

namespace gnngls {
struct KnArgs {
    int n;
    const void *ft;                       // [M,128] FT
    const float *el, *er;                 // [M,8] log2 domain
    float *recV;                          // partial numerators: [M,128] of the lower star (gat_kn.cu) / [M,2,128] of both stars (gat_kn_tc.cu)
    float *recDM;                         // (denominator, reference max) per head: [M,8,2] / [M,2,8,2]
    int *flags;                           // [B*n*HG] "star has published" (gat_kn.cu) / [B] stars out per instance (gat_kn_tc.cu)
    const float *h, *bias, *bn_scale, *bn_shift;
    float *h1, *h1_tf32;
};
}  // namespace gnngls

namespace {   // (internal linkage: each translation unit gets its own copy)

constexpr int D_ = GNNGLS_EMBED_DIM;   // 128
constexpr int H_ = GNNGLS_HEADS;       // 8
constexpr int F_ = GNNGLS_HEAD_DIM;    // 16
constexpr float kSlope = 0.2f;
constexpr float kFixFrac = 0.9f;       // self weight / row total above which the arg-max row is redone exactly

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lrelu(float s) { return fmaxf(s, kSlope * s); }
__device__ __forceinline__ uint32_t tf32_bits(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float4 tf32_round4(float4 v) {
    return make_float4(__uint_as_float(tf32_bits(v.x)), __uint_as_float(tf32_bits(v.y)),
                       __uint_as_float(tf32_bits(v.z)), __uint_as_float(tf32_bits(v.w)));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// the 128-byte line at `p` will not be read again: a dirty copy in L2 need not be written back
__device__ __forceinline__ void discard_l2_128(const void *p) {
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int COUNT>
__device__ __forceinline__ void group_barrier(int id) {
    if (COUNT == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory");
}

// line-graph node id of the TSP edge {a,b}, a != b (sorted-tuple order, datasets.py:56-60)
__host__ __device__ __forceinline__ int kn_node(int a, int b, int n) {
    const int i = a < b ? a : b, j = a < b ? b : a;
    return i * (2 * n - i - 1) / 2 + (j - i - 1);
}

// 4 consecutive features from shared memory
__device__ __forceinline__ float4 lds_ft4(const unsigned char *p, float) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 lds_ft4(const unsigned char *p, __half) {
    const uint2 raw = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
    return make_float4(a.x, a.y, b.x, b.y);
}



using gnngls::KnArgs;
}  // namespace

namespace gnngls {
// tcgen05 variant (fp16 features, n <= 128); defined in gat_kn_tc.cu.  Lays out its own records in `workspace`.
size_t kn_tc_workspace_bytes(int B, int n);
int launch_kn_tc(const KnArgs &args, int B, void *workspace, cudaStream_t st);
}  // namespace gnngls
