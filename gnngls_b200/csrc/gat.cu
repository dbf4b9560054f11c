// GAT edge-softmax / aggregate for the edge-regret model (dgl.nn.GATConv reached from
// gnngls/models.py:23; semantics in SURVEY.md Appendix A), fused with the skip connection and the
// first BatchNorm of the layer (models.py:12-15,27):
//
//     h1[v] = BN1( h[v] + sum_u softmax_u( leaky_relu(el[u] + er[v], 0.2) ) * ft[u]  (+ bias) )
//
// This file: gat_csr_kernel, any destination-sorted CSR graph.  One warp per destination; per-head
// maxima, then tiles of 32 in-edges: each lane evaluates the 8 head weights of ONE edge (no
// redundant exp), parks them in shared memory, and the warp gathers the 32 source rows with
// 128-bit loads.  Pure fp32.  The line graph of K_n has its own kernel (gat_kn.cu).
#include <cuda_fp16.h>
#include <cstdint>
#include <cmath>
#include "common.h"

namespace {

constexpr int D_ = GNNGLS_EMBED_DIM;   // 128
constexpr int H_ = GNNGLS_HEADS;       // 8
constexpr float kSlope = 0.2f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t tf32_bits(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lrelu(float s) { return fmaxf(s, kSlope * s); }
__device__ __forceinline__ float4 tf32_round4(float4 v) {
    return make_float4(__uint_as_float(tf32_bits(v.x)), __uint_as_float(tf32_bits(v.y)),
                       __uint_as_float(tf32_bits(v.z)), __uint_as_float(tf32_bits(v.w)));
}

// ================================================================================================
// generic CSR path
// ================================================================================================
constexpr int CSR_WARPS = 8;

// 4 consecutive features of one ft row (fp32 or fp16 storage)
template <typename FT>
__device__ __forceinline__ float4 load_ft4(const FT *row, int lane);
template <>
__device__ __forceinline__ float4 load_ft4<float>(const float *row, int lane) {
    return *reinterpret_cast<const float4 *>(row + 4 * lane);
}
template <>
__device__ __forceinline__ float4 load_ft4<__half>(const __half *row, int lane) {
    const uint2 raw = *reinterpret_cast<const uint2 *>(row + 4 * lane);
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

template <typename FT>
__global__ void __launch_bounds__(CSR_WARPS * 32)
gat_csr_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, int64_t M,
               const FT *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
               const float *__restrict__ h, const float *__restrict__ bias, const float *__restrict__ bn_scale,
               const float *__restrict__ bn_shift, float *__restrict__ h1, float *__restrict__ h1_tf32) {
    __shared__ __align__(16) float ps[CSR_WARPS][32][H_];   // edge-tile weights
    __shared__ int us[CSR_WARPS][32];                      // edge-tile sources
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hh = lane >> 2;
    for (int64_t v = (int64_t)blockIdx.x * CSR_WARPS + warp; v < M; v += (int64_t)gridDim.x * CSR_WARPS) {
        const int e0 = indptr[v], e1 = indptr[v + 1];
        float erv[H_];
        {
            const float4 a = *reinterpret_cast<const float4 *>(er + v * H_);
            const float4 b = *reinterpret_cast<const float4 *>(er + v * H_ + 4);
            erv[0] = a.x; erv[1] = a.y; erv[2] = a.z; erv[3] = a.w;      // el/er arrive pre-multiplied by log2(e)
            erv[4] = b.x; erv[5] = b.y; erv[6] = b.z; erv[7] = b.w;
        }
        // pass A: per-head maximum of the (log2-scaled) scores over the in-edges
        float mx[H_];
#pragma unroll
        for (int k = 0; k < H_; ++k) mx[k] = -INFINITY;
        for (int e = e0 + lane; e < e1; e += 32) {
            const int u = indices[e];
            const float4 a = *reinterpret_cast<const float4 *>(el + (int64_t)u * H_);
            const float4 b = *reinterpret_cast<const float4 *>(el + (int64_t)u * H_ + 4);
            const float l8[H_] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < H_; ++k) mx[k] = fmaxf(mx[k], lrelu(l8[k] + erv[k]));
        }
#pragma unroll
        for (int k = 0; k < H_; ++k)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
        // pass B: weights per edge tile, then gather-accumulate
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float den = 0.f;
        for (int base = e0; base < e1; base += 32) {
            const int e = base + lane;
            int u = 0;
            float p[H_];
            if (e < e1) {
                u = indices[e];
                const float4 a = *reinterpret_cast<const float4 *>(el + (int64_t)u * H_);
                const float4 b = *reinterpret_cast<const float4 *>(el + (int64_t)u * H_ + 4);
                const float l8[H_] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < H_; ++k) p[k] = ex2(lrelu(l8[k] + erv[k]) - mx[k]);
            } else {
#pragma unroll
                for (int k = 0; k < H_; ++k) p[k] = 0.f;
            }
            *reinterpret_cast<float4 *>(&ps[warp][lane][0]) = make_float4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<float4 *>(&ps[warp][lane][4]) = make_float4(p[4], p[5], p[6], p[7]);
            us[warp][lane] = u;
            __syncwarp();
            const int cnt = min(32, e1 - base);
            int t = 0;
            for (; t + 4 <= cnt; t += 4) {
                float4 f[4];
                float w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    f[q] = load_ft4<FT>(ft + (int64_t)us[warp][t + q] * D_, lane);
                    w[q] = ps[warp][t + q][hh];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc.x = fmaf(w[q], f[q].x, acc.x); acc.y = fmaf(w[q], f[q].y, acc.y);
                    acc.z = fmaf(w[q], f[q].z, acc.z); acc.w = fmaf(w[q], f[q].w, acc.w);
                    den += w[q];
                }
            }
            for (; t < cnt; ++t) {
                const float4 f = load_ft4<FT>(ft + (int64_t)us[warp][t] * D_, lane);
                const float w = ps[warp][t][hh];
                acc.x = fmaf(w, f.x, acc.x); acc.y = fmaf(w, f.y, acc.y);
                acc.z = fmaf(w, f.z, acc.z); acc.w = fmaf(w, f.w, acc.w);
                den += w;
            }
            __syncwarp();
        }
        const float inv = den > 0.f ? 1.f / den : 0.f;
        float4 g = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        if (bias) {
            const float4 bb = *reinterpret_cast<const float4 *>(bias + 4 * lane);
            g.x += bb.x; g.y += bb.y; g.z += bb.z; g.w += bb.w;
        }
        const float4 hv = *reinterpret_cast<const float4 *>(h + v * D_ + 4 * lane);
        const float4 sc = *reinterpret_cast<const float4 *>(bn_scale + 4 * lane);
        const float4 sh = *reinterpret_cast<const float4 *>(bn_shift + 4 * lane);
        float4 o;
        o.x = (hv.x + g.x) * sc.x + sh.x; o.y = (hv.y + g.y) * sc.y + sh.y;
        o.z = (hv.z + g.z) * sc.z + sh.z; o.w = (hv.w + g.w) * sc.w + sh.w;
        *reinterpret_cast<float4 *>(h1 + v * D_ + 4 * lane) = o;
        if (h1_tf32) *reinterpret_cast<float4 *>(h1_tf32 + v * D_ + 4 * lane) = tf32_round4(o);
    }
}

}  // namespace

extern "C" int gnngls_gat_aggregate_csr(const int32_t *indptr, const int32_t *indices, int64_t M, const void *ft,
                                        int ft_dtype, const float *el, const float *er, const float *h,
                                        const float *gat_bias, const float *bn_scale, const float *bn_shift, float *h1,
                                        float *h1_tf32, void *stream) {
    GNNGLS_REQUIRE(indptr && indices && ft && el && er && h && bn_scale && bn_shift && h1, GNNGLS_ERR_BAD_ARG,
                   "null pointer argument");
    GNNGLS_REQUIRE(ft_dtype == GNNGLS_FT_F32 || ft_dtype == GNNGLS_FT_TF32 || ft_dtype == GNNGLS_FT_F16, GNNGLS_ERR_BAD_ARG,
                   "unknown ft_dtype %d", ft_dtype);
    if (M <= 0) return GNNGLS_OK;
    const int64_t blocks = (M + CSR_WARPS - 1) / CSR_WARPS;
    const int64_t cap = (int64_t)gnngls::device_sm_count() * 64;
    const int grid = (int)(blocks < cap ? blocks : cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ft_dtype == GNNGLS_FT_F16)
        gat_csr_kernel<__half><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, M, static_cast<const __half *>(ft), el, er, h,
                                                                gat_bias, bn_scale, bn_shift, h1, h1_tf32);
    else
        gat_csr_kernel<float><<<grid, CSR_WARPS * 32, 0, st>>>(indptr, indices, M, static_cast<const float *>(ft), el, er, h,
                                                               gat_bias, bn_scale, bn_shift, h1, h1_tf32);
    GNNGLS_LAUNCH_OK("gat_csr_kernel");
    return GNNGLS_OK;
}

