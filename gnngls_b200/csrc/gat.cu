// GAT edge-softmax / aggregate for the edge-regret model (dgl.nn.GATConv reached from
// gnngls/models.py:23; semantics in SURVEY.md Appendix A), fused with the skip connection and the
// first BatchNorm of the layer (models.py:12-15,27):
//
//     h1[v] = BN1( h[v] + sum_u softmax_u( leaky_relu(el[u] + er[v], 0.2) ) * ft[u]  (+ bias) )
//
// Two implementations behind the C ABI:
//  * gat_csr_kernel   — any destination-sorted CSR graph.  One warp per destination; per-head
//    maxima, then tiles of 32 in-edges: each lane evaluates the 8 head weights of ONE edge (no
//    redundant exp), parks them in shared memory, and the warp gathers the 32 source rows with
//    128-bit loads.  Pure fp32.
//  * gat_kn_*         — line graph of K_n with the adjacency computed arithmetically: node (i,j)
//    receives from the "stars" of vertex i and vertex j.  One CTA per (instance, vertex) stages
//    that vertex's star (n-1 rows) in shared memory ONCE and produces, for all n-1 destinations
//    that contain the vertex, the partial numerator/denominator/max of their softmax with
//    mma.sync TF32 (attention weights are generated directly in the A-fragment registers,
//    flash-attention style).  Each destination belongs to two stars: the CTA that finishes second
//    (arrival counter) merges the two partials and applies bias + skip + BatchNorm in its own
//    epilogue.  Every ft row is read from L2/HBM exactly twice per layer.
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <type_traits>
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
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
// the 128-byte line at `p` will not be read again: a dirty copy in L2 need not be written back
__device__ __forceinline__ void discard_l2_128(const void *p) {
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
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

// ================================================================================================
// K_n line-graph path
// ================================================================================================
// line-graph node id of the TSP edge {a,b}, a != b (sorted-tuple order, datasets.py:56-60)
__host__ __device__ __forceinline__ int kn_node(int a, int b, int n) {
    const int i = a < b ? a : b, j = a < b ? b : a;
    return i * (2 * n - i - 1) / 2 + (j - i - 1);
}

constexpr int FS_LD = D_ + 8;   // 136: B-fragment reads (k=t, n=g) hit bank 8t+g -> conflict-free
constexpr int STAR_THREADS = 256;

struct Top2 { float m1, m2; int a1; };
__device__ __forceinline__ Top2 top2_merge(Top2 x, Top2 y) {
    Top2 r;
    if (x.m1 >= y.m1) { r.m1 = x.m1; r.a1 = x.a1; r.m2 = fmaxf(x.m2, y.m1); }
    else { r.m1 = y.m1; r.a1 = y.a1; r.m2 = fmaxf(y.m2, x.m1); }
    return r;
}

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t star_smem_bytes(int n) {
    const int KP = round_up(n, 8), MP = round_up(n, 16);
    return sizeof(float) * ((size_t)KP * FS_LD + (size_t)KP * H_ + (size_t)MP * H_ + 3 * H_ + (size_t)KP);
}

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Partials: one 576-byte record per (destination, contributing star), destination-major:
//   part[((b*N + v)*2 + slot)*144 + ...]   v = line-graph node {i,j}, slot 0 = star of min(i,j), slot 1 = star of max(i,j)
//   [0,128): sum_k w*ft   [128,136): sum_k w   [136,144): log2-scaled max used for w
// so the two records of a destination, and the records of consecutive destinations, are contiguous.
constexpr int REC = D_ + 2 * H_;   // 144 floats
struct StarCtx {
    int n, i, b, hd, g, t;
    float m1, m2;
    int a1;
    const float *Fh;      // star features, offset to this warp's head
    const float *ELs, *ERs;
    int ksteps;
    float *part;          // partial records of instance b (part + b*N*2*REC)
    int skip_row;         // destination the tile loop must not publish (-1: none)
    const int *NODE;      // [KP] line-graph node of {i, k} (shared memory; -1 for dead slots)
};

// publish one destination row of this warp's head: lane t of the quad holds features {2t,2t+1} and
// {8+2t,9+2t}.  (.cg stores: the partials are consumed by another SM through L2.)
__device__ __forceinline__ void publish_row(const StarCtx &c, int j, float2 n0, float2 n1, float den, float mx) {
    if (j >= c.n || j == c.i || j == c.skip_row) return;
    float *rec = c.part + (size_t)(c.NODE[j] * 2 + (c.i > j)) * REC;
    float *o = rec + c.hd * 16 + 2 * c.t;
    __stcg(reinterpret_cast<float2 *>(o), n0);
    __stcg(reinterpret_cast<float2 *>(o + 8), n1);
    if (c.t == 0) { __stcg(rec + D_ + c.hd, den); __stcg(rec + D_ + H_ + c.hd, mx); }
}

// NT m-tiles (16 destinations each) processed together so that el and the B fragments of a k-step
// are loaded once for NT tiles.  Per attention weight: FADD + FFMA + FMNMX + MUFU.EX2; the weights
// go to the tensor core as raw fp32 bits (hardware keeps the top 19 bits) and the softmax
// denominator comes from a third MMA against a ones fragment, so numerator and denominator see
// identically truncated weights.
template <int NT>
__device__ __forceinline__ void star_tiles(const StarCtx &c, int mt0) {
    float c1a[NT][2], c2a[NT][2], mxa[NT][2];           // per tile, rows (lo,hi): er-mx, 0.2*er-mx, mx
    float acc0[NT][4], acc1[NT][4], accs[NT][4];
#pragma unroll
    for (int u = 0; u < NT; ++u) {
        const int j_lo = (mt0 + u) * 16 + c.g, j_hi = j_lo + 8;
        const float er_lo = c.ERs[j_lo * H_ + c.hd], er_hi = c.ERs[j_hi * H_ + c.hd];
        mxa[u][0] = lrelu(((c.a1 == j_lo) ? c.m2 : c.m1) + er_lo);
        mxa[u][1] = lrelu(((c.a1 == j_hi) ? c.m2 : c.m1) + er_hi);
        c1a[u][0] = er_lo - mxa[u][0]; c2a[u][0] = kSlope * er_lo - mxa[u][0];
        c1a[u][1] = er_hi - mxa[u][1]; c2a[u][1] = kSlope * er_hi - mxa[u][1];
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc0[u][q] = 0.f; acc1[u][q] = 0.f; accs[u][q] = 0.f; }
    }
    const uint32_t one = __float_as_uint(1.0f);
    // one k-step (8 star members) for the NT tiles; DIAG = this k-step may contain a destination's own slot
    auto kstep = [&](int ks, auto diag_tag) {
        constexpr bool DIAG = decltype(diag_tag)::value;
        const int k_lo = ks * 8 + c.t, k_hi = k_lo + 4;
        const float el_lo = c.ELs[k_lo * H_ + c.hd], el_hi = c.ELs[k_hi * H_ + c.hd];
        const float *r_lo = c.Fh + (size_t)k_lo * FS_LD + c.g, *r_hi = c.Fh + (size_t)k_hi * FS_LD + c.g;
        const uint32_t b00 = __float_as_uint(r_lo[0]), b01 = __float_as_uint(r_hi[0]);
        const uint32_t b10 = __float_as_uint(r_lo[8]), b11 = __float_as_uint(r_hi[8]);
#pragma unroll
        for (int u = 0; u < NT; ++u) {
            // leaky_relu(el+er) - mx == max(el + (er-mx), 0.2*el + (0.2*er-mx))
            float w0 = ex2(fmaxf(el_lo + c1a[u][0], fmaf(kSlope, el_lo, c2a[u][0])));   // (row lo, col k_lo)
            float w1 = ex2(fmaxf(el_lo + c1a[u][1], fmaf(kSlope, el_lo, c2a[u][1])));   // (row hi, col k_lo)
            float w2 = ex2(fmaxf(el_hi + c1a[u][0], fmaf(kSlope, el_hi, c2a[u][0])));   // (row lo, col k_hi)
            float w3 = ex2(fmaxf(el_hi + c1a[u][1], fmaf(kSlope, el_hi, c2a[u][1])));   // (row hi, col k_hi)
            if (DIAG) {                                  // a node is not its own neighbour
                const int j_lo = (mt0 + u) * 16 + c.g, j_hi = j_lo + 8;
                if (k_lo == j_lo) w0 = 0.f;
                if (k_lo == j_hi) w1 = 0.f;
                if (k_hi == j_lo) w2 = 0.f;
                if (k_hi == j_hi) w3 = 0.f;
            }
            const uint32_t a[4] = {__float_as_uint(w0), __float_as_uint(w1), __float_as_uint(w2), __float_as_uint(w3)};
            mma_tf32_16x8x8(acc0[u], a, b00, b01);
            mma_tf32_16x8x8(acc1[u], a, b10, b11);
            mma_tf32_16x8x8(accs[u], a, one, one);       // row sums of the (truncated) weights
        }
    };
    // destinations of tile mt sit on the diagonal of k-steps 2mt and 2mt+1 only: peel those so the bulk of
    // the loop carries no masking code
    const int d0 = min(2 * mt0, c.ksteps), d1 = min(2 * (mt0 + NT), c.ksteps);
    for (int ks = 0; ks < d0; ++ks) kstep(ks, std::false_type{});
    for (int ks = d0; ks < d1; ++ks) kstep(ks, std::true_type{});
    for (int ks = d1; ks < c.ksteps; ++ks) kstep(ks, std::false_type{});
#pragma unroll
    for (int u = 0; u < NT; ++u) {
        const int j_lo = (mt0 + u) * 16 + c.g;
        publish_row(c, j_lo, make_float2(acc0[u][0], acc0[u][1]), make_float2(acc1[u][0], acc1[u][1]), accs[u][0], mxa[u][0]);
        publish_row(c, j_lo + 8, make_float2(acc0[u][2], acc0[u][3]), make_float2(acc1[u][2], acc1[u][3]), accs[u][2], mxa[u][1]);
    }
}

// Second half of both star kernels.  `scratch` is shared memory for 2n+5 ints that the main loop no longer needs.
template <int FIN_U>
__device__ __forceinline__ void star_finish(int n, int i, int b, int64_t node0, int *scratch, float *__restrict__ part,
                                            int *__restrict__ arrive,
                                            const float *__restrict__ h, const float *__restrict__ bias,
                                            const float *__restrict__ bn_scale, const float *__restrict__ bn_shift,
                                            float *__restrict__ h1, float *__restrict__ h1_tf32) {
    // ---- every destination {i,j} belongs to the stars of i and of j.  Publish this star's partials,
    // then bump the destination's arrival counter: the star that arrives second merges the two
    // partials (flash-style rescale, always in (min(i,j), max(i,j)) order so the result does not
    // depend on arrival order) and applies bias + skip + BatchNorm1.  No separate combine pass.
    __threadfence();
    __syncthreads();
    int *second = scratch;                                    // [n] flags
    for (int j = threadIdx.x; j < n; j += STAR_THREADS)
        second[j] = (j != i) ? atomicAdd(arrive + node0 + kn_node(i, j, n), 1) : 0;
    __syncthreads();
    __threadfence();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hh = lane >> 2;
    const float4 sc = *reinterpret_cast<const float4 *>(bn_scale + 4 * lane);
    const float4 sh = *reinterpret_cast<const float4 *>(bn_shift + 4 * lane);
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bb = *reinterpret_cast<const float4 *>(bias + 4 * lane);
    // compact the destinations this CTA must finish, then FIN_U per warp iteration (more loads in flight)
    int *todo = second + round_up(n, 4);                      // [n] compacted list, count in todo[n]
    __syncthreads();
    if (warp == 0) {
        int cnt = 0;
        for (int base = 0; base < n; base += 32) {
            const int j = base + lane;
            const bool f = j < n && second[j] != 0;
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (f) todo[cnt + __popc(m & ((1u << lane) - 1))] = j;
            cnt += __popc(m);
        }
        if (lane == 0) todo[n] = cnt;
    }
    __syncthreads();
    const int cnt = todo[n];
    for (int q0 = warp * FIN_U; q0 < cnt; q0 += (STAR_THREADS / 32) * FIN_U) {
        float4 n1[FIN_U], n2[FIN_U], hv[FIN_U];
        float x1[FIN_U], x2[FIN_U], d1[FIN_U], d2[FIN_U];
        int64_t v[FIN_U];
        int jj[FIN_U];
        bool ok[FIN_U];
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            ok[u] = q0 + u < cnt;
            const int j = todo[ok[u] ? q0 + u : q0];
            v[u] = node0 + kn_node(i, j, n);
            const float *r1 = part + (size_t)v[u] * 2 * REC, *r2 = r1 + REC;       // records of the lower / higher vertex's star
            n1[u] = __ldcg(reinterpret_cast<const float4 *>(r1 + 4 * lane));
            n2[u] = __ldcg(reinterpret_cast<const float4 *>(r2 + 4 * lane));
            x1[u] = __ldcg(r1 + D_ + H_ + hh); x2[u] = __ldcg(r2 + D_ + H_ + hh);
            d1[u] = __ldcg(r1 + D_ + hh); d2[u] = __ldcg(r2 + D_ + hh);
            hv[u] = *reinterpret_cast<const float4 *>(h + v[u] * D_ + 4 * lane);
            jj[u] = j;
        }
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            if (!ok[u]) continue;
            // both partial rows are dead now (each is read exactly once): drop their dirty L2 lines instead of
            // writing 1 KB per destination back to HBM
            if (lane < 9) discard_l2_128(part + (size_t)v[u] * 2 * REC + lane * 32);   // 2 records = 1152 B = 9 lines
            const float mx = fmaxf(x1[u], x2[u]);
            const float s1 = ex2(x1[u] - mx), s2 = ex2(x2[u] - mx);
            const float inv = 1.f / fmaf(d1[u], s1, d2[u] * s2);
            const float a1 = s1 * inv, a2 = s2 * inv;
            float4 o;
            o.x = (hv[u].x + (fmaf(n1[u].x, a1, n2[u].x * a2) + bb.x)) * sc.x + sh.x;
            o.y = (hv[u].y + (fmaf(n1[u].y, a1, n2[u].y * a2) + bb.y)) * sc.y + sh.y;
            o.z = (hv[u].z + (fmaf(n1[u].z, a1, n2[u].z * a2) + bb.z)) * sc.z + sh.z;
            o.w = (hv[u].w + (fmaf(n1[u].w, a1, n2[u].w * a2) + bb.w)) * sc.w + sh.w;
            *reinterpret_cast<float4 *>(h1 + v[u] * D_ + 4 * lane) = o;
            if (h1_tf32) *reinterpret_cast<float4 *>(h1_tf32 + v[u] * D_ + 4 * lane) = tf32_round4(o);
        }
    }
}

__global__ void __launch_bounds__(STAR_THREADS, 3)
gat_kn_star_kernel(int n, const float *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
                   float *__restrict__ part,
                   int *__restrict__ arrive, const float *__restrict__ h, const float *__restrict__ bias,
                   const float *__restrict__ bn_scale, const float *__restrict__ bn_shift, float *__restrict__ h1,
                   float *__restrict__ h1_tf32, int ft_is_tf32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int KP = round_up(n, 8), MP = round_up(n, 16);
    float *Fs = reinterpret_cast<float *>(smem_raw);          // [KP][FS_LD] tf32-rounded ft rows of the star
    float *ELs = Fs + (size_t)KP * FS_LD;                     // [KP][8]  el (log2 domain; dead slots: -inf)
    float *ERs = ELs + (size_t)KP * H_;                       // [MP][8]  er (log2 domain; dead slots: 0)
    float *TM1 = ERs + (size_t)MP * H_;                       // [8] max over the star
    float *TM2 = TM1 + H_;                                    // [8] second max
    int *TA1 = reinterpret_cast<int *>(TM2 + H_);             // [8] arg of the max

    const int b = blockIdx.x / n, i = blockIdx.x - b * n;
    const int64_t N = (int64_t)n * (n - 1) / 2;
    const int64_t node0 = (int64_t)b * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- stage the star of vertex i (slot k <-> TSP edge {i,k}) with cp.async: every row of the star
    // is in flight at once, so the CTA pays one L2 latency instead of one per row
    int *NODE = TA1 + H_;                                     // [KP] line-graph node of slot k, -1 for dead slots
    for (int k = threadIdx.x; k < KP; k += STAR_THREADS) NODE[k] = (k < n && k != i) ? kn_node(i, k, n) : -1;
    __syncthreads();
    for (int idx = threadIdx.x; idx < KP * 32; idx += STAR_THREADS) {
        const int k = idx >> 5, q = idx & 31;                 // 16-byte piece q of row k
        const int node = NODE[k];
        float *dst = Fs + (size_t)k * FS_LD + 4 * q;
        if (node >= 0) cp_async16(dst, ft + (node0 + node) * D_ + 4 * q);
        else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int idx = threadIdx.x; idx < MP * 4; idx += STAR_THREADS) {
        const int k = idx >> 2, q = idx & 3;                  // el row = 2 pieces, er row = 2 pieces
        const int node = k < KP ? NODE[k] : -1;
        if (node >= 0) {
            if (q < 2) cp_async16(ELs + k * H_ + 4 * q, el + (node0 + node) * H_ + 4 * q);
            else cp_async16(ERs + k * H_ + 4 * (q - 2), er + (node0 + node) * H_ + 4 * (q - 2));
        } else {                                              // dead slots: weight 0 as a source, unused as a destination
            const float fill = q < 2 ? -INFINITY : 0.f;
            if (q >= 2 || k < KP)
                *reinterpret_cast<float4 *>((q < 2 ? ELs : ERs) + k * H_ + 4 * (q & 1)) = make_float4(fill, fill, fill, fill);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (!ft_is_tf32) {                                        // producer did not round: the tensor core would truncate
        for (int idx = threadIdx.x; idx < KP * 32; idx += STAR_THREADS) {
            float4 *ptr = reinterpret_cast<float4 *>(Fs + (size_t)(idx >> 5) * FS_LD + 4 * (idx & 31));
            *ptr = tf32_round4(*ptr);
        }
        __syncthreads();
    }

    // ---- per-head top-2 of el over the star (warp w <-> head w)
    {
        Top2 t2{-INFINITY, -INFINITY, -1};
        for (int k = lane; k < KP; k += 32) {
            const float x = ELs[k * H_ + warp];
            if (x > t2.m1) { t2.m2 = t2.m1; t2.m1 = x; t2.a1 = k; }
            else if (x > t2.m2) t2.m2 = x;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Top2 o;
            o.m1 = __shfl_xor_sync(0xffffffffu, t2.m1, off);
            o.m2 = __shfl_xor_sync(0xffffffffu, t2.m2, off);
            o.a1 = __shfl_xor_sync(0xffffffffu, t2.a1, off);
            t2 = top2_merge(t2, o);
        }
        if (lane == 0) { TM1[warp] = t2.m1; TM2[warp] = t2.m2; TA1[warp] = t2.a1; }
    }
    __syncthreads();

    // ---- main loop: warp <-> head; m-tiles of 16 destinations, k-steps of 8 star members
    StarCtx c;
    c.n = n; c.i = i; c.b = b; c.hd = warp; c.g = lane >> 2; c.t = lane & 3;
    c.m1 = TM1[warp]; c.m2 = TM2[warp]; c.a1 = TA1[warp];
    c.Fh = Fs + warp * 16; c.ELs = ELs; c.ERs = ERs; c.ksteps = KP / 8;
    {
        c.part = part + (size_t)node0 * 2 * REC;
        c.skip_row = -1;
        c.NODE = NODE;
    }
    const int MT = MP / 16;
    int mt = 0;
    for (; mt + 2 <= MT; mt += 2) star_tiles<2>(c, mt);
    if (mt < MT) star_tiles<1>(c, mt);

    star_finish<3>(n, i, b, node0, reinterpret_cast<int *>(ELs), part, arrive, h, bias, bn_scale, bn_shift, h1, h1_tf32);
}


// ================================================================================================
// fp16-operand star kernel (default tensor-core path).  fp16 has the same 10-bit mantissa as TF32, so
// storing ft and the attention weights (0 < w <= 1) as fp16 loses nothing against the TF32 kernel
// above, while the star needs half the bytes from HBM and half the shared memory, one
// mma.m16n8k16 does the work of two m16n8k8, and a single ldmatrix.x4.trans feeds the B fragments of
// a whole 16-member k-step.  Weights are rounded to nearest (cvt.rn.f16x2), not truncated.
// ================================================================================================
constexpr int FH_LD = D_ + 8;   // halves; 272-byte rows: the 8 rows of an ldmatrix tile fall in distinct 16-byte bank groups

__host__ __device__ inline size_t star16_smem_bytes(int n) {
    const int KP = round_up(n, 8), KE = round_up(n, 16);
    return (size_t)KP * FH_LD * sizeof(__half) + sizeof(float) * ((size_t)4 * KE * H_ + 3 * H_ + (size_t)KP);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void mma_f16_16x8x16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

struct Star16Ctx {
    StarCtx base;          // n, i, hd, g, t, m1, m2, a1, ERs and the partial-row pointers (Fh/ELs/ksteps unused)
    const float *ELt;      // this head's el over the star, [KE] (log2 domain; dead slots -inf)
    const float2 *EA;      // this head's (2^(el-m1), 2^(0.2(el-m1))) over the star, [KE] (dead slots 0)
    uint32_t b4_addr;      // this lane's ldmatrix.x4 row address for k-step 0
    uint32_t b2_addr;      // this lane's ldmatrix.x2 row address for the 8-member tail step
    int kfull;             // number of 16-member k-steps
    bool tail;             // an 8-member step follows (KP % 16 == 8)
};

// Attention weight without a per-edge exponential.  With s = el_k + er_j and mx_j the destination's maximum,
//   2^(leaky_relu(s) - mx_j) = max(2^(s - mx_j), 2^(0.2 s - mx_j))                    (2^x is monotone)
//                            = max(A_k * C1_j, A'_k * C2_j)
//   A_k = 2^(el_k - m1), A'_k = 2^(0.2 (el_k - m1))                one pair per star member and head
//   C1_j = 2^(m1 + er_j - mx_j), C2_j = 2^(0.2 (m1 + er_j) - mx_j)  one pair per destination and head
// m1 = max_k el_k and mx_j = leaky_relu(m1 + er_j) >= the row's true maximum (any upper bound is a valid softmax
// reference; the merge uses the published mx_j).  All four factors lie in [0, 1]: nothing overflows, and a factor
// only underflows when the weight does.  Only the arg-max member's own row can sit far below its reference (its
// sources exclude itself): when the runner-up m2 is more than 6 log2 units down, star16_fix_row redoes that row.
// Cost per weight: FMUL + FMUL + FMNMX instead of FADD + FFMA + FMNMX + MUFU.EX2 -- the SFU pipe (16/clk/SM),
// which bounded the loop, is out of it.
__device__ __forceinline__ float att_w(float a, float a5, float c1, float c2) { return fmaxf(a * c1, a5 * c2); }

template <int NT>
__device__ __forceinline__ void star16_tiles(const Star16Ctx &c, int mt0) {
    const StarCtx &b = c.base;
    float c1a[NT][2], c2a[NT][2], mxa[NT][2];
    float acc0[NT][4], acc1[NT][4], accs[NT][4];
#pragma unroll
    for (int u = 0; u < NT; ++u) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = (mt0 + u) * 16 + b.g + 8 * r;
            const float s = b.m1 + b.ERs[j * H_ + b.hd];
            mxa[u][r] = lrelu(s);
            c1a[u][r] = ex2(s - mxa[u][r]);
            c2a[u][r] = ex2(kSlope * s - mxa[u][r]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc0[u][q] = 0.f; acc1[u][q] = 0.f; accs[u][q] = 0.f; }
    }
    constexpr uint32_t ones = 0x3C003C00u;                    // half2(1, 1)
    auto kstep16 = [&](int ks, auto diag_tag) {
        constexpr bool DIAG = decltype(diag_tag)::value;
        const int ka = ks * 16 + 2 * b.t, kb = ka + 8;
        const float4 pa = *reinterpret_cast<const float4 *>(c.EA + ka);      // (A, A') of members ka, ka+1
        const float4 pb = *reinterpret_cast<const float4 *>(c.EA + kb);      //          ... of members kb, kb+1
        uint32_t B[4];
        ldmatrix_x4_trans(B, c.b4_addr + (uint32_t)ks * (16 * FH_LD * 2));
#pragma unroll
        for (int u = 0; u < NT; ++u) {
            float w00 = att_w(pa.x, pa.y, c1a[u][0], c2a[u][0]), w01 = att_w(pa.z, pa.w, c1a[u][0], c2a[u][0]);   // row lo, k = ka, ka+1
            float w10 = att_w(pa.x, pa.y, c1a[u][1], c2a[u][1]), w11 = att_w(pa.z, pa.w, c1a[u][1], c2a[u][1]);   // row hi
            float w20 = att_w(pb.x, pb.y, c1a[u][0], c2a[u][0]), w21 = att_w(pb.z, pb.w, c1a[u][0], c2a[u][0]);   // row lo, k = kb, kb+1
            float w30 = att_w(pb.x, pb.y, c1a[u][1], c2a[u][1]), w31 = att_w(pb.z, pb.w, c1a[u][1], c2a[u][1]);   // row hi
            if (DIAG) {                                   // a node is not its own neighbour
                const int j_lo = (mt0 + u) * 16 + b.g, j_hi = j_lo + 8;
                if (ka == j_lo) w00 = 0.f;
                if (ka + 1 == j_lo) w01 = 0.f;
                if (ka == j_hi) w10 = 0.f;
                if (ka + 1 == j_hi) w11 = 0.f;
                if (kb == j_lo) w20 = 0.f;
                if (kb + 1 == j_lo) w21 = 0.f;
                if (kb == j_hi) w30 = 0.f;
                if (kb + 1 == j_hi) w31 = 0.f;
            }
            const uint32_t a[4] = {pack_f16x2(w00, w01), pack_f16x2(w10, w11), pack_f16x2(w20, w21), pack_f16x2(w30, w31)};
            mma_f16_16x8x16(acc0[u], a, B[0], B[1]);
            mma_f16_16x8x16(acc1[u], a, B[2], B[3]);
            mma_f16_16x8x16(accs[u], a, ones, ones);     // row sums of the rounded weights
        }
    };
    // destinations of tile mt sit on the diagonal of k-step mt only (or of the tail step): peel those
    const int d0 = min(mt0, c.kfull), d1 = min(mt0 + NT, c.kfull);
    for (int ks = 0; ks < d0; ++ks) kstep16(ks, std::false_type{});
    for (int ks = d0; ks < d1; ++ks) kstep16(ks, std::true_type{});
    for (int ks = d1; ks < c.kfull; ++ks) kstep16(ks, std::false_type{});
    if (c.tail) {
        const int ka = c.kfull * 16 + 2 * b.t;
        const float4 pa = *reinterpret_cast<const float4 *>(c.EA + ka);
        uint32_t B[2];
        ldmatrix_x2_trans(B, c.b2_addr);
#pragma unroll
        for (int u = 0; u < NT; ++u) {
            const int j_lo = (mt0 + u) * 16 + b.g, j_hi = j_lo + 8;
            float w00 = att_w(pa.x, pa.y, c1a[u][0], c2a[u][0]), w01 = att_w(pa.z, pa.w, c1a[u][0], c2a[u][0]);
            float w10 = att_w(pa.x, pa.y, c1a[u][1], c2a[u][1]), w11 = att_w(pa.z, pa.w, c1a[u][1], c2a[u][1]);
            if (ka == j_lo) w00 = 0.f;
            if (ka + 1 == j_lo) w01 = 0.f;
            if (ka == j_hi) w10 = 0.f;
            if (ka + 1 == j_hi) w11 = 0.f;
            const uint32_t a0 = pack_f16x2(w00, w01), a1 = pack_f16x2(w10, w11);
            mma_f16_16x8x8(acc0[u], a0, a1, B[0]);
            mma_f16_16x8x8(acc1[u], a0, a1, B[1]);
            mma_f16_16x8x8(accs[u], a0, a1, ones);
        }
    }
#pragma unroll
    for (int u = 0; u < NT; ++u) {
        const int j_lo = (mt0 + u) * 16 + b.g;
        publish_row(b, j_lo, make_float2(acc0[u][0], acc0[u][1]), make_float2(acc1[u][0], acc1[u][1]), accs[u][0], mxa[u][0]);
        publish_row(b, j_lo + 8, make_float2(acc0[u][2], acc0[u][3]), make_float2(acc1[u][2], acc1[u][3]), accs[u][2], mxa[u][1]);
    }
}

// The destination that is itself the head's arg-max star member, when the other members are far below it: its
// weights relative to m1 would be < 2^-6 and lose fp16 precision (or flush to zero).  One row per head and star:
// the warp evaluates it directly in fp32 against m2 -- lanes over members, shuffle reduction -- and publishes it.
__device__ __forceinline__ void star16_fix_row(const Star16Ctx &c, const __half *Fh, int KP, int lane) {
    const StarCtx &b = c.base;
    const int j = b.a1;
    if (j < 0 || j >= b.n || j == b.i) return;            // (cannot happen: the arg-max is a live member)
    const float erj = b.ERs[j * H_ + b.hd];
    const float mx = lrelu(b.m2 + erj);
    float num[16], den = 0.f;
#pragma unroll
    for (int f = 0; f < 16; ++f) num[f] = 0.f;
    for (int k = lane; k < KP; k += 32) {
        const float w = (k == j) ? 0.f : ex2(lrelu(c.ELt[k] + erj) - mx);     // dead slots: el = -inf -> 0
        const uint4 r0 = *reinterpret_cast<const uint4 *>(Fh + (size_t)k * FH_LD + b.hd * 16);
        const uint4 r1 = *reinterpret_cast<const uint4 *>(Fh + (size_t)k * FH_LD + b.hd * 16 + 8);
        const uint32_t raw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float2 v = __half22float2(*reinterpret_cast<const __half2 *>(&raw[q]));
            num[2 * q] = fmaf(w, v.x, num[2 * q]);
            num[2 * q + 1] = fmaf(w, v.y, num[2 * q + 1]);
        }
        den += w;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        den += __shfl_xor_sync(0xffffffffu, den, off);
#pragma unroll
        for (int f = 0; f < 16; ++f) num[f] += __shfl_xor_sync(0xffffffffu, num[f], off);
    }
    if (lane == 0) {
        float *rec = b.part + ((size_t)kn_node(b.i, j, b.n) * 2 + (b.i > j)) * REC;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            __stcg(reinterpret_cast<float4 *>(rec + b.hd * 16 + 4 * q), make_float4(num[4 * q], num[4 * q + 1], num[4 * q + 2], num[4 * q + 3]));
        __stcg(rec + D_ + b.hd, den);
        __stcg(rec + D_ + H_ + b.hd, mx);
    }
}

template <bool FUSED>
__device__ __forceinline__ void star_f16_body(int n, const __half *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
                       float *__restrict__ part,
                       int *__restrict__ arrive, const float *__restrict__ h, const float *__restrict__ bias,
                       const float *__restrict__ bn_scale, const float *__restrict__ bn_shift, float *__restrict__ h1,
                       float *__restrict__ h1_tf32, int group) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (n <= 0) return;                                                       // warm-up launch (module loading)
    const int KP = round_up(n, 8), KE = round_up(n, 16);
    __half *Fh = reinterpret_cast<__half *>(smem_raw);                        // [KP][FH_LD] fp16 ft rows of the star
    float *ELt = reinterpret_cast<float *>(Fh + (size_t)KP * FH_LD);          // [8][KE] el, head-major (dead slots: -inf)
    float *ERs = ELt + (size_t)H_ * KE;                                       // [KE][8] er (dead slots: 0)
    float2 *EA = reinterpret_cast<float2 *>(ERs + (size_t)KE * H_);           // [8][KE] (2^(el-m1), 2^(0.2(el-m1))), head-major
    float *TM1 = reinterpret_cast<float *>(EA + (size_t)H_ * KE);
    float *TM2 = TM1 + H_;
    int *TA1 = reinterpret_cast<int *>(TM2 + H_);
    int *NODE = TA1 + H_;                                                     // [KP]

    const int b = blockIdx.x / n, i = blockIdx.x - b * n;
    const int64_t N = (int64_t)n * (n - 1) / 2;
    const int64_t node0 = (int64_t)b * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int k = threadIdx.x; k < KP; k += STAR_THREADS) NODE[k] = (k < n && k != i) ? kn_node(i, k, n) : -1;
    __syncthreads();
    for (int idx = threadIdx.x; idx < KP * 16; idx += STAR_THREADS) {
        const int k = idx >> 4, q = idx & 15;                 // 16-byte piece q (8 halves) of row k
        const int node = NODE[k];
        __half *dst = Fh + (size_t)k * FH_LD + 8 * q;
        if (node >= 0) cp_async16(dst, ft + (node0 + node) * D_ + 8 * q);
        else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int idx = threadIdx.x; idx < KE * 2; idx += STAR_THREADS) {
        const int k = idx >> 1, q = idx & 1;
        const int node = k < KP ? NODE[k] : -1;
        if (node >= 0) cp_async16(ERs + k * H_ + 4 * q, er + (node0 + node) * H_ + 4 * q);
        else *reinterpret_cast<float4 *>(ERs + k * H_ + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int k = threadIdx.x; k < KE; k += STAR_THREADS) {    // el, transposed to head-major
        const int node = k < KP ? NODE[k] : -1;
        float4 a = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), c4 = a;
        if (node >= 0) {
            a = *reinterpret_cast<const float4 *>(el + (node0 + node) * H_);
            c4 = *reinterpret_cast<const float4 *>(el + (node0 + node) * H_ + 4);
        }
        ELt[0 * KE + k] = a.x; ELt[1 * KE + k] = a.y; ELt[2 * KE + k] = a.z; ELt[3 * KE + k] = a.w;
        ELt[4 * KE + k] = c4.x; ELt[5 * KE + k] = c4.y; ELt[6 * KE + k] = c4.z; ELt[7 * KE + k] = c4.w;
    }
    cp_async_wait_all();
    __syncthreads();

    {   // per-head top-2 of el over the star (warp w <-> head w)
        Top2 t2{-INFINITY, -INFINITY, -1};
        for (int k = lane; k < KP; k += 32) {
            const float x = ELt[warp * KE + k];
            if (x > t2.m1) { t2.m2 = t2.m1; t2.m1 = x; t2.a1 = k; }
            else if (x > t2.m2) t2.m2 = x;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Top2 o;
            o.m1 = __shfl_xor_sync(0xffffffffu, t2.m1, off);
            o.m2 = __shfl_xor_sync(0xffffffffu, t2.m2, off);
            o.a1 = __shfl_xor_sync(0xffffffffu, t2.a1, off);
            t2 = top2_merge(t2, o);
        }
        // every lane now holds the warp's result; no shared-memory round trip needed
        Star16Ctx c;
        c.base.n = n; c.base.i = i; c.base.b = b; c.base.hd = warp; c.base.g = lane >> 2; c.base.t = lane & 3;
        c.base.m1 = t2.m1; c.base.m2 = t2.m2; c.base.a1 = t2.a1;
        c.base.Fh = nullptr; c.base.ELs = nullptr; c.base.ERs = ERs; c.base.ksteps = 0;
        c.base.part = part + (size_t)node0 * 2 * REC;
        // the arg-max member's own row: fine in the shared factorisation unless the runner-up is far below
        // (its weights would sink towards fp16 subnormals); then it is redone exactly
        const bool fix = t2.m1 - t2.m2 > 6.f;
        c.base.skip_row = fix ? t2.a1 : -1;
        c.base.NODE = NODE;
        c.ELt = ELt + warp * KE;
        c.EA = EA + warp * KE;
        for (int k = lane; k < KE; k += 32) {                                 // per-member factors of this warp's head
            const float d = ELt[warp * KE + k] - t2.m1;
            EA[warp * KE + k] = make_float2(ex2(d), ex2(kSlope * d));
        }
        __syncwarp();
        const uint32_t fh = (uint32_t)__cvta_generic_to_shared(Fh);
        const int q = lane >> 3, r = lane & 7;
        c.b4_addr = fh + (uint32_t)(((r + 8 * (q & 1)) * FH_LD + warp * 16 + 8 * (q >> 1)) * 2);
        c.kfull = KP / 16;
        c.tail = (KP & 8) != 0;
        c.b2_addr = fh + (uint32_t)(((c.kfull * 16 + r) * FH_LD + warp * 16 + 8 * (q & 1)) * 2);
        const int MT = KE / 16;
        int mt = 0;
        for (; mt + 2 <= MT; mt += 2) star16_tiles<2>(c, mt);
        if (mt < MT) star16_tiles<1>(c, mt);
        if (fix) star16_fix_row(c, Fh, KP, lane);
    }
    if (FUSED) {
        star_finish<3>(n, i, b, node0, reinterpret_cast<int *>(ELt), part, arrive, h, bias, bn_scale, bn_shift, h1, h1_tf32);
    } else {
        // split pipeline: this star is done once its partials are visible; the concurrently running combine
        // kernel waits for the per-group count (arrive[g] == stars of the group) and finishes the destinations
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(arrive + b / group, 1);
    }
}

// fused variant: 80 registers (the finalize keeps many loads in flight), 3 CTAs per SM
__global__ void __launch_bounds__(STAR_THREADS, 3)
gat_kn_star_f16_fused_kernel(int n, const __half *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
                             float *__restrict__ part,
                             int *__restrict__ arrive, const float *__restrict__ h, const float *__restrict__ bias,
                             const float *__restrict__ bn_scale, const float *__restrict__ bn_shift, float *__restrict__ h1,
                             float *__restrict__ h1_tf32) {
    star_f16_body<true>(n, ft, el, er, part, arrive, h, bias, bn_scale, bn_shift, h1, h1_tf32, 1);
}
// split variant: 72 registers, so that three of these CTAs leave 10240 registers per SM for one combine CTA
__global__ void __maxnreg__(72)
gat_kn_star_f16_split_kernel(int n, const __half *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
                             float *__restrict__ part,
                             int *__restrict__ done, int group) {
    star_f16_body<false>(n, ft, el, er, part, done, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, group);
}

// ================================================================================================
// Combine kernel of the split pipeline.  A few small persistent CTAs (one per SM, 128 threads, <= 32
// registers: they fit beside the three resident star CTAs) run CONCURRENTLY with the star kernel on a
// second stream.  For each group of instances they wait until all of its stars have published, then
// every CTA takes a slice of the group's destinations: both partial rows and the skip row are pulled into
// shared memory with cp.async (memory-level parallelism without registers), merged in the canonical
// (min(i,j), max(i,j)) order, and bias + skip + BatchNorm1 are applied.  The star CTAs therefore spend
// their whole life in load -> main loop -> publish, and the partials are consumed while still in L2.
// ================================================================================================
constexpr int CMB_THREADS = 256;          // 1 producer warp + 7 merging warps; 256 x 40 registers = what three star CTAs leave free
constexpr int CMB_WARPS = CMB_THREADS / 32;
constexpr int CMB_CONSUMERS = CMB_THREADS - 32;
constexpr int CMB_BATCH = 32;             // consecutive destinations per batch
constexpr int CMB_STAGES = 2;
constexpr int CMB_PART_FLOATS = CMB_BATCH * 2 * REC;     // both records of every destination: one contiguous 36 KB piece
constexpr int CMB_H_FLOATS = CMB_BATCH * D_;             // the skip rows: one contiguous 16 KB piece
constexpr int CMB_BUF_FLOATS = CMB_PART_FLOATS + CMB_H_FLOATS;
// batch buffers + bn_scale/bn_shift/bias + per-(destination, head) merge factors + full/empty mbarriers
constexpr size_t CMB_SMEM = sizeof(float) * (CMB_STAGES * CMB_BUF_FLOATS + 3 * D_ + 2 * CMB_BATCH * H_) + 2 * CMB_STAGES * sizeof(uint64_t);

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bulk global -> shared copy (TMA engine: no registers, one instruction per contiguous piece) signalling `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// named barriers over the whole combine CTA: arrive only / arrive and wait
__device__ __forceinline__ void ready_arrive(int slot) {
    if (slot == 0) asm volatile("bar.arrive 1, 256;" ::: "memory");
    else asm volatile("bar.arrive 2, 256;" ::: "memory");
}
__device__ __forceinline__ void ready_sync(int slot) {
    if (slot == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
    else asm volatile("bar.sync 2, 256;" ::: "memory");
}
static_assert(CMB_THREADS == 256 && CMB_CONSUMERS == 224 && CMB_STAGES == 2, "named barriers hard-code the CTA shape");

// The batches a combine CTA owns, in order: groups ascending, inside a group every gridDim-th batch starting
// at a per-group rotated offset (so the ragged last round does not always hit the same CTAs).
struct CmbIter {
    int g;            // current group (== groups when exhausted)
    int64_t base;     // first destination (group-relative) of the current batch
};
__device__ __forceinline__ int64_t cmb_group_total(int g, int B, int group, int64_t N) { return (int64_t)min(group, B - g * group) * N; }
__device__ __forceinline__ int64_t cmb_first(int g) { return (int64_t)((blockIdx.x + (unsigned)g * 37u) % gridDim.x) * CMB_BATCH; }
__device__ __forceinline__ void cmb_settle(CmbIter &it, int B, int group, int groups, int64_t N) {
    while (it.g < groups && it.base >= cmb_group_total(it.g, B, group, N)) {
        ++it.g;
        if (it.g < groups) it.base = cmb_first(it.g);
    }
}
__device__ __forceinline__ void cmb_next(CmbIter &it, int B, int group, int groups, int64_t N) {
    it.base += (int64_t)gridDim.x * CMB_BATCH;
    cmb_settle(it, B, group, groups, N);
}

__global__ void __maxnreg__(40)
gat_kn_combine_kernel(int B, int n, int group, const float *__restrict__ part, const int *__restrict__ done,
                      const float *__restrict__ h, const float *__restrict__ bias, const float *__restrict__ bn_scale,
                      const float *__restrict__ bn_shift, float *__restrict__ h1, float *__restrict__ h1_tf32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *bufs = reinterpret_cast<float *>(smem_raw);                               // [STAGES][part records | h rows]
    float *affine = bufs + CMB_STAGES * CMB_BUF_FLOATS;                              // bn_scale | bn_shift | bias
    float2 *factor = reinterpret_cast<float2 *>(affine + 3 * D_);                    // [CMB_BATCH][8]: weights of the two partial numerators
    uint64_t *full = reinterpret_cast<uint64_t *>(factor + CMB_BATCH * H_);          // [STAGES] bytes landed
    uint64_t *empty = full + CMB_STAGES;                                             // [STAGES] merging warps are done with the buffer
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t N = (int64_t)n * (n - 1) / 2;
    const int groups = group > 0 ? (B + group - 1) / group : 0;
    if (groups == 0) return;                                                         // warm-up launch (module loading)
    for (int c = threadIdx.x; c < D_; c += CMB_THREADS) {
        affine[c] = bn_scale[c];
        affine[D_ + c] = bn_shift[c];
        affine[2 * D_ + c] = bias ? bias[c] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < CMB_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CMB_WARPS - 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ---- producer.  A small polling state machine (one lane decides, the warp follows): issue the next batch
        // as soon as its group is complete and its buffer is free (up to two in flight), release a batch to the
        // merging warps as soon as its bytes have landed.  It only ever sleeps, it never blocks on one condition
        // while the other could make progress.
        CmbIter it{0, cmb_first(0)};
        cmb_settle(it, B, group, groups, N);
        int k_issue = 0, k_ready = 0, known_done = -1;
        uint32_t fphase = 0, ephase = 0;                   // bit s = parity to wait for on full[s] / empty[s]
        unsigned long long t_last = 0;
        constexpr int poll_every = 8;                      // idle rounds (64 ns each) between two polls of a group counter
        int idle = poll_every;
        while (it.g < groups || k_ready < k_issue) {
            int act = 0;                                   // 1 = issue, 2 = release
            if (lane == 0) {
                if (it.g < groups && k_issue - k_ready < CMB_STAGES) {
                    const int s = k_issue % CMB_STAGES;
                    const bool buffer_free = k_issue < CMB_STAGES || mbar_try_wait(&empty[s], (ephase >> s) & 1);
                    if (buffer_free) {
                        // the group counter lives in one L2 line that every combine CTA polls: keep that to one
                        // poll per ~0.5 us per CTA, or the pollers saturate the line's L2 slice
                        if (it.g <= known_done) act = 1;
                        else if (idle >= poll_every) {
                            idle = 0;
                            if (ld_acquire_gpu(done + it.g) >= min(group, B - it.g * group) * n) act = 1;
                        }
                    }
                }
                if (!act && k_ready < k_issue && mbar_try_wait(&full[k_ready % CMB_STAGES], (fphase >> (k_ready % CMB_STAGES)) & 1)) act = 2;
            }
            act = __shfl_sync(0xffffffffu, act, 0);
            if (act == 1) {
                const int s = k_issue % CMB_STAGES;
                if (k_issue >= CMB_STAGES) ephase ^= 1u << s;
                known_done = it.g;
                const int64_t total = cmb_group_total(it.g, B, group, N);
                const int cnt = (int)min((int64_t)CMB_BATCH, total - it.base);
                const int64_t d0 = (int64_t)it.g * group * N + it.base;              // first destination (global node index)
                float *buf = bufs + s * CMB_BUF_FLOATS;
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[s], (uint32_t)cnt * (2 * REC + D_) * 4);
                    bulk_g2s(buf, part + (size_t)d0 * 2 * REC, (uint32_t)cnt * 2 * REC * 4, &full[s]);
                    bulk_g2s(buf + CMB_PART_FLOATS, h + (size_t)d0 * D_, (uint32_t)cnt * D_ * 4, &full[s]);
                }
                cmb_next(it, B, group, groups, N);
                ++k_issue;
                t_last = 0;
                idle = poll_every;
            } else if (act == 2) {
                const int s = k_ready % CMB_STAGES;
                fphase ^= 1u << s;
                ready_arrive(s);
                ++k_ready;
                t_last = 0;
            } else {
                __nanosleep(64);
                ++idle;
                if (lane == 0) {                           // fail loudly instead of hanging the device if nothing moves for 20 s
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t_last == 0) t_last = t1;
                    else if (t1 - t_last > 20000000000ull) asm volatile("trap;");
                }
            }
        }
    } else {
        // ---- merging warps
        CmbIter it{0, cmb_first(0)};
        cmb_settle(it, B, group, groups, N);
        const int hh = lane >> 2, ctid = threadIdx.x - 32;
        uint32_t fphase = 0;
        for (int k = 0; it.g < groups; ++k) {
            const int s = k % CMB_STAGES;
            const int cnt = (int)min((int64_t)CMB_BATCH, cmb_group_total(it.g, B, group, N) - it.base);
            const int64_t d0 = (int64_t)it.g * group * N + it.base;
            ready_sync(s);
            while (!mbar_try_wait(&full[s], (fphase >> s) & 1)) {}     // already complete: acquires the bulk-copied bytes for this thread
            fphase ^= 1u << s;
            const float *buf = bufs + s * CMB_BUF_FLOATS;
            // pass 1: one thread per (destination, head) evaluates the flash-style merge factors, so the three
            // dependent MUFU operations are paid once per batch instead of once per destination row
            for (int e = ctid; e < cnt * H_; e += CMB_CONSUMERS) {
                const float *r = buf + (e >> 3) * 2 * REC + D_ + (e & 7);
                const float d1 = r[0], x1 = r[H_], d2 = r[REC], x2 = r[REC + H_];
                const float mx = fmaxf(x1, x2);
                const float s1 = ex2(x1 - mx), s2 = ex2(x2 - mx);
                const float inv = 1.f / fmaf(d1, s1, d2 * s2);
                factor[e] = make_float2(s1 * inv, s2 * inv);
            }
            asm volatile("bar.sync 3, 224;" ::: "memory");
            // pass 2: warp per destination, lane per 4 features
            for (int q = warp - 1; q < cnt; q += CMB_WARPS - 1) {
                const float *r = buf + q * 2 * REC;
                const float4 n1 = *reinterpret_cast<const float4 *>(r + 4 * lane);
                const float4 n2 = *reinterpret_cast<const float4 *>(r + REC + 4 * lane);
                const float4 hv = *reinterpret_cast<const float4 *>(buf + CMB_PART_FLOATS + q * D_ + 4 * lane);
                const float2 a = factor[q * H_ + hh];
                const float4 bb = *reinterpret_cast<const float4 *>(affine + 2 * D_ + 4 * lane);
                float4 o;
                o.x = hv.x + (fmaf(n1.x, a.x, n2.x * a.y) + bb.x);
                o.y = hv.y + (fmaf(n1.y, a.x, n2.y * a.y) + bb.y);
                o.z = hv.z + (fmaf(n1.z, a.x, n2.z * a.y) + bb.z);
                o.w = hv.w + (fmaf(n1.w, a.x, n2.w * a.y) + bb.w);
                const float4 sc = *reinterpret_cast<const float4 *>(affine + 4 * lane);
                const float4 sh = *reinterpret_cast<const float4 *>(affine + D_ + 4 * lane);
                o.x = o.x * sc.x + sh.x; o.y = o.y * sc.y + sh.y; o.z = o.z * sc.z + sh.z; o.w = o.w * sc.w + sh.w;
                *reinterpret_cast<float4 *>(h1 + (d0 + q) * D_ + 4 * lane) = o;
                if (h1_tf32) *reinterpret_cast<float4 *>(h1_tf32 + (d0 + q) * D_ + 4 * lane) = tf32_round4(o);
                // both records are dead now (each is read exactly once): drop their 9 dirty L2 lines
                if (lane < 9) discard_l2_128(part + (size_t)(d0 + q) * 2 * REC + lane * 32);
            }
            asm volatile("bar.sync 3, 224;" ::: "memory");             // `factor` is rewritten by the next batch
            if (lane == 0) mbar_arrive(&empty[s]);                     // this warp has left the buffer
            cmb_next(it, B, group, groups, N);
        }
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

namespace {
// second stream for the combine kernel (non-blocking: it must not serialise against the legacy default stream,
// the star kernel it waits for may have been launched there) + fork/join events
struct SidePipe {
    cudaStream_t side;
    cudaEvent_t fork, join;
};
SidePipe *side_pipe() {
    static SidePipe pipe;
    static bool ok = [] {
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return false;
        if (cudaStreamCreateWithPriority(&pipe.side, cudaStreamNonBlocking, greatest) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&pipe.fork, cudaEventDisableTiming) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&pipe.join, cudaEventDisableTiming) != cudaSuccess) return false;
        // With lazy module loading the first launch of a kernel may synchronise the context, which would deadlock
        // against the spinning combine kernel: force both kernels to be resident before the first fork.
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, gat_kn_star_f16_split_kernel) != cudaSuccess) return false;
        if (cudaFuncGetAttributes(&fa, gat_kn_combine_kernel) != cudaSuccess) return false;
        if (cudaFuncSetAttribute(gat_kn_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CMB_SMEM) != cudaSuccess) return false;
        // The two kernels share every SM: three star CTAs beside one combine CTA only fit under the largest
        // shared-memory carve-out, and the carve-out cannot change while the (first-launched) combine CTA is resident.
        if (cudaFuncSetAttribute(gat_kn_combine_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) return false;
        if (cudaFuncSetAttribute(gat_kn_star_f16_split_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) return false;
        // ... and run each once as a no-op (n = 0 / B = 0) so that nothing is left to load on the first real launch
        gat_kn_star_f16_split_kernel<<<1, STAR_THREADS, 0, pipe.side>>>(0, nullptr, nullptr, nullptr, nullptr, nullptr, 1);
        gat_kn_combine_kernel<<<1, CMB_THREADS, CMB_SMEM, pipe.side>>>(0, 3, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                        nullptr);
        if (cudaStreamSynchronize(pipe.side) != cudaSuccess) return false;
        return true;
    }();
    return ok ? &pipe : nullptr;
}
// GNNGLS_STAR_PIPELINE=split selects the two-kernel pipeline (star kernel + concurrently running combine kernel).
// Measured on B200 it only ties the fused single-kernel variant, and two kernels that must
// be co-resident cannot run under tools that serialise kernels (ncu, compute-sanitizer), so fused is the default.
bool split_pipeline_enabled() {
    static const bool split = [] {
        const char *e = getenv("GNNGLS_STAR_PIPELINE");
        return e && (e[0] == 's' || e[0] == 'S');
    }();
    return split;
}
}  // namespace

extern "C" size_t gnngls_gat_kn_workspace_bytes(int B, int n) {
    if (B <= 0 || n <= 0) return 0;
    const size_t N = (size_t)n * (n - 1) / 2;
    return sizeof(float) * (size_t)B * N * 2 * REC + sizeof(int) * (size_t)B * N;
}

extern "C" int gnngls_gat_aggregate_kn(int B, int n, const void *ft, int ft_dtype, const float *el, const float *er,
                                       const float *h, const float *gat_bias, const float *bn_scale,
                                       const float *bn_shift, float *h1, float *h1_tf32, void *workspace,
                                       size_t workspace_bytes, void *stream) {
    GNNGLS_REQUIRE(ft && el && er && h && bn_scale && bn_shift && h1, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(ft_dtype == GNNGLS_FT_F32 || ft_dtype == GNNGLS_FT_TF32 || ft_dtype == GNNGLS_FT_F16, GNNGLS_ERR_BAD_ARG,
                   "unknown ft_dtype %d", ft_dtype);
    GNNGLS_REQUIRE(n >= 3, GNNGLS_ERR_UNSUPPORTED, "line graph of K_n needs n >= 3 (got %d)", n);
    if (B <= 0) return GNNGLS_OK;
    const bool f16 = ft_dtype == GNNGLS_FT_F16;
    const size_t smem = f16 ? star16_smem_bytes(n) : star_smem_bytes(n);
    GNNGLS_REQUIRE(smem <= (size_t)gnngls::device_max_optin_smem(), GNNGLS_ERR_UNSUPPORTED,
                   "n=%d: a vertex star (%zu B) does not fit shared memory; use the CSR path", n, smem);
    GNNGLS_REQUIRE((int64_t)B * n < (int64_t)1 << 31, GNNGLS_ERR_UNSUPPORTED, "B*n too large for one launch");
    GNNGLS_REQUIRE(workspace && workspace_bytes >= gnngls_gat_kn_workspace_bytes(B, n), GNNGLS_ERR_WORKSPACE,
                   "gat_kn workspace too small: need %zu bytes", gnngls_gat_kn_workspace_bytes(B, n));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t N = (size_t)n * (n - 1) / 2;
    float *part = static_cast<float *>(workspace);
    int *arrive = reinterpret_cast<int *>(part + (size_t)B * N * 2 * REC);
    if (f16 && split_pipeline_enabled()) {
        // destinations per group >= one full round of the combine CTAs
        const int sms = gnngls::device_sm_count();
        int group = (int)(((size_t)sms * CMB_BATCH + N - 1) / N);
        if (group > B) group = B;
        const int groups = (B + group - 1) / group;
        SidePipe *sp = side_pipe();
        GNNGLS_REQUIRE(sp != nullptr, GNNGLS_ERR_CUDA, "could not create the combine stream");
        if (smem > 48 * 1024)      // before the fork: nothing but the two launches may happen while the combine kernel spins
            GNNGLS_CUDA_OK(cudaFuncSetAttribute(gat_kn_star_f16_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GNNGLS_CUDA_OK(cudaMemsetAsync(arrive, 0, sizeof(int) * (size_t)groups, st));
        GNNGLS_CUDA_OK(cudaEventRecord(sp->fork, st));
        GNNGLS_CUDA_OK(cudaStreamWaitEvent(sp->side, sp->fork, 0));
        gat_kn_combine_kernel<<<sms, CMB_THREADS, CMB_SMEM, sp->side>>>(B, n, group, part, arrive, h, gat_bias, bn_scale,
                                                                        bn_shift, h1, h1_tf32);
        GNNGLS_LAUNCH_OK("gat_kn_combine_kernel");
        gat_kn_star_f16_split_kernel<<<B * n, STAR_THREADS, smem, st>>>(n, static_cast<const __half *>(ft), el, er, part, arrive, group);
        GNNGLS_LAUNCH_OK("gat_kn_star_f16_kernel");
        GNNGLS_CUDA_OK(cudaEventRecord(sp->join, sp->side));
        GNNGLS_CUDA_OK(cudaStreamWaitEvent(st, sp->join, 0));
        return GNNGLS_OK;
    }
    GNNGLS_CUDA_OK(cudaMemsetAsync(arrive, 0, sizeof(int) * (size_t)B * N, st));
    if (f16) {
        if (smem > 48 * 1024)
            GNNGLS_CUDA_OK(cudaFuncSetAttribute(gat_kn_star_f16_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gat_kn_star_f16_fused_kernel<<<B * n, STAR_THREADS, smem, st>>>(n, static_cast<const __half *>(ft), el, er, part, arrive,
                                                                        h, gat_bias, bn_scale, bn_shift, h1, h1_tf32);
        GNNGLS_LAUNCH_OK("gat_kn_star_f16_kernel");
    } else {
        if (smem > 48 * 1024)
            GNNGLS_CUDA_OK(cudaFuncSetAttribute(gat_kn_star_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gat_kn_star_kernel<<<B * n, STAR_THREADS, smem, st>>>(n, static_cast<const float *>(ft), el, er, part, arrive,
                                                              h, gat_bias, bn_scale, bn_shift, h1, h1_tf32,
                                                              ft_dtype == GNNGLS_FT_TF32);
        GNNGLS_LAUNCH_OK("gat_kn_star_kernel");
    }
    return GNNGLS_OK;
}
