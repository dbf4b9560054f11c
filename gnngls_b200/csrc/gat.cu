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
//    flash-attention style).  A combine kernel merges the two partials of each destination and
//    applies skip + BatchNorm.  Every ft row is read from L2/HBM exactly twice per layer.
#include <cstdint>
#include <cmath>
#include "common.h"

namespace {

constexpr int D_ = GNNGLS_EMBED_DIM;   // 128
constexpr int H_ = GNNGLS_HEADS;       // 8
constexpr float kSlope = 0.2f;
constexpr float kLog2e = 1.4426950408889634f;

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

__global__ void __launch_bounds__(CSR_WARPS * 32)
gat_csr_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, int64_t M,
               const float *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
               const float *__restrict__ h, const float *__restrict__ bias, const float *__restrict__ bn_scale,
               const float *__restrict__ bn_shift, float *__restrict__ h1, int round_tf32) {
    __shared__ __align__(16) float ps[CSR_WARPS][32][H_];   // edge-tile weights
    __shared__ int us[CSR_WARPS][32];                      // edge-tile sources
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hh = lane >> 2;
    for (int64_t v = (int64_t)blockIdx.x * CSR_WARPS + warp; v < M; v += (int64_t)gridDim.x * CSR_WARPS) {
        const int e0 = indptr[v], e1 = indptr[v + 1];
        float erv[H_];
        {
            const float4 a = *reinterpret_cast<const float4 *>(er + v * H_);
            const float4 b = *reinterpret_cast<const float4 *>(er + v * H_ + 4);
            erv[0] = a.x * kLog2e; erv[1] = a.y * kLog2e; erv[2] = a.z * kLog2e; erv[3] = a.w * kLog2e;
            erv[4] = b.x * kLog2e; erv[5] = b.y * kLog2e; erv[6] = b.z * kLog2e; erv[7] = b.w * kLog2e;
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
            for (int k = 0; k < H_; ++k) mx[k] = fmaxf(mx[k], lrelu(fmaf(l8[k], kLog2e, erv[k])));
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
                for (int k = 0; k < H_; ++k) p[k] = ex2(lrelu(fmaf(l8[k], kLog2e, erv[k])) - mx[k]);
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
                    f[q] = *reinterpret_cast<const float4 *>(ft + (int64_t)us[warp][t + q] * D_ + 4 * lane);
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
                const float4 f = *reinterpret_cast<const float4 *>(ft + (int64_t)us[warp][t] * D_ + 4 * lane);
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
        if (round_tf32) o = tf32_round4(o);
        *reinterpret_cast<float4 *>(h1 + v * D_ + 4 * lane) = o;
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
    return sizeof(float) * ((size_t)KP * FS_LD + (size_t)KP * H_ + (size_t)MP * H_ + 3 * H_);
}

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Partials, indexed [(b*n + i)*n + j] for star vertex i and destination {i,j}:
//   pnum[.,128]: sum_k w*ft   pden[.,8]: sum_k w   pmax[.,8]: log2-scaled max used for w
__global__ void __launch_bounds__(STAR_THREADS)
gat_kn_star_kernel(int n, const float *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
                   float *__restrict__ pnum, float *__restrict__ pden, float *__restrict__ pmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int KP = round_up(n, 8), MP = round_up(n, 16);
    float *Fs = reinterpret_cast<float *>(smem_raw);          // [KP][FS_LD] tf32-rounded ft rows of the star
    float *ELs = Fs + (size_t)KP * FS_LD;                     // [KP][8]  el*log2e   (dead slots: -inf)
    float *ERs = ELs + (size_t)KP * H_;                       // [MP][8]  er*log2e   (dead slots: 0)
    float *TM1 = ERs + (size_t)MP * H_;                       // [8] max over the star
    float *TM2 = TM1 + H_;                                    // [8] second max
    int *TA1 = reinterpret_cast<int *>(TM2 + H_);             // [8] arg of the max

    const int b = blockIdx.x / n, i = blockIdx.x - b * n;
    const int64_t N = (int64_t)n * (n - 1) / 2;
    const int64_t node0 = (int64_t)b * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- stage the star of vertex i: slot k <-> TSP edge {i,k}
    for (int k = warp; k < KP; k += STAR_THREADS / 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool live = (k < n) && (k != i);
        int64_t node = 0;
        if (live) {
            node = node0 + kn_node(i, k, n);
            v = *reinterpret_cast<const float4 *>(ft + node * D_ + 4 * lane);
        }
        uint4 t;
        t.x = tf32_bits(v.x); t.y = tf32_bits(v.y); t.z = tf32_bits(v.z); t.w = tf32_bits(v.w);
        *reinterpret_cast<uint4 *>(Fs + (size_t)k * FS_LD + 4 * lane) = t;
        if (lane < H_) ELs[k * H_ + lane] = live ? el[node * H_ + lane] * kLog2e : -INFINITY;
        else if (lane < 2 * H_) ERs[k * H_ + (lane - H_)] = live ? er[node * H_ + (lane - H_)] * kLog2e : 0.f;
    }
    for (int idx = KP * H_ + threadIdx.x; idx < MP * H_; idx += STAR_THREADS) ERs[idx] = 0.f;
    __syncthreads();

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
    const int hd = warp;
    const int g = lane >> 2, t = lane & 3;
    const float m1 = TM1[hd], m2 = TM2[hd];
    const int a1 = TA1[hd];
    const float *Fh = Fs + hd * 16;
    const int ksteps = KP / 8;
    for (int mt = 0; mt < MP / 16; ++mt) {
        const int j_lo = mt * 16 + g, j_hi = j_lo + 8;
        const float er_lo = ERs[j_lo * H_ + hd], er_hi = ERs[j_hi * H_ + hd];
        // softmax shift of this partial: leaky_relu is monotone, so the max score over the star
        // minus the destination itself is leaky_relu(max_k el + er)
        const float mx_lo = lrelu(((a1 == j_lo) ? m2 : m1) + er_lo);
        const float mx_hi = lrelu(((a1 == j_hi) ? m2 : m1) + er_hi);
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
        float sum_lo = 0.f, sum_hi = 0.f;
        for (int ks = 0; ks < ksteps; ++ks) {
            const int k_lo = ks * 8 + t, k_hi = k_lo + 4;
            const float el_lo = ELs[k_lo * H_ + hd], el_hi = ELs[k_hi * H_ + hd];
            float w0 = ex2(lrelu(el_lo + er_lo) - mx_lo);   // (row j_lo, col k_lo)
            float w1 = ex2(lrelu(el_lo + er_hi) - mx_hi);   // (row j_hi, col k_lo)
            float w2 = ex2(lrelu(el_hi + er_lo) - mx_lo);   // (row j_lo, col k_hi)
            float w3 = ex2(lrelu(el_hi + er_hi) - mx_hi);   // (row j_hi, col k_hi)
            if ((ks >> 1) == mt) {                          // tile touches the diagonal: a node is not its own neighbour
                if (k_lo == j_lo) w0 = 0.f;
                if (k_lo == j_hi) w1 = 0.f;
                if (k_hi == j_lo) w2 = 0.f;
                if (k_hi == j_hi) w3 = 0.f;
            }
            uint32_t a[4];
            a[0] = tf32_bits(w0); a[1] = tf32_bits(w1); a[2] = tf32_bits(w2); a[3] = tf32_bits(w3);
            sum_lo += __uint_as_float(a[0]) + __uint_as_float(a[2]);
            sum_hi += __uint_as_float(a[1]) + __uint_as_float(a[3]);
            const float *r_lo = Fh + (size_t)k_lo * FS_LD + g, *r_hi = Fh + (size_t)k_hi * FS_LD + g;
            mma_tf32_16x8x8(c0, a, __float_as_uint(r_lo[0]), __float_as_uint(r_hi[0]));
            mma_tf32_16x8x8(c1, a, __float_as_uint(r_lo[8]), __float_as_uint(r_hi[8]));
        }
        sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1); sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
        sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1); sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
        const int64_t prow = ((int64_t)b * n + i) * n;
        if (j_lo < n && j_lo != i) {
            float *o = pnum + (prow + j_lo) * D_ + hd * 16 + 2 * t;
            *reinterpret_cast<float2 *>(o) = make_float2(c0[0], c0[1]);
            *reinterpret_cast<float2 *>(o + 8) = make_float2(c1[0], c1[1]);
            if (t == 0) { pden[(prow + j_lo) * H_ + hd] = sum_lo; pmax[(prow + j_lo) * H_ + hd] = mx_lo; }
        }
        if (j_hi < n && j_hi != i) {
            float *o = pnum + (prow + j_hi) * D_ + hd * 16 + 2 * t;
            *reinterpret_cast<float2 *>(o) = make_float2(c0[2], c0[3]);
            *reinterpret_cast<float2 *>(o + 8) = make_float2(c1[2], c1[3]);
            if (t == 0) { pden[(prow + j_hi) * H_ + hd] = sum_hi; pmax[(prow + j_hi) * H_ + hd] = mx_hi; }
        }
    }
}

// merge the two partials of destination {i,j} (flash-style rescale), add bias + skip, BatchNorm1
__global__ void __launch_bounds__(128)
gat_kn_combine_kernel(int n, const float *__restrict__ pnum, const float *__restrict__ pden,
                      const float *__restrict__ pmax, const float *__restrict__ h, const float *__restrict__ bias,
                      const float *__restrict__ bn_scale, const float *__restrict__ bn_shift, float *__restrict__ h1,
                      int round_tf32) {
    const int b = blockIdx.x / n, i = blockIdx.x - b * n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hh = lane >> 2;
    const int64_t N = (int64_t)n * (n - 1) / 2;
    const float4 sc = *reinterpret_cast<const float4 *>(bn_scale + 4 * lane);
    const float4 sh = *reinterpret_cast<const float4 *>(bn_shift + 4 * lane);
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bb = *reinterpret_cast<const float4 *>(bias + 4 * lane);
    for (int j = i + 1 + warp; j < n; j += 4) {
        const int64_t p1 = ((int64_t)b * n + i) * n + j, p2 = ((int64_t)b * n + j) * n + i;
        const int64_t v = (int64_t)b * N + kn_node(i, j, n);
        const float4 n1 = *reinterpret_cast<const float4 *>(pnum + p1 * D_ + 4 * lane);
        const float4 n2 = *reinterpret_cast<const float4 *>(pnum + p2 * D_ + 4 * lane);
        const float x1 = pmax[p1 * H_ + hh], x2 = pmax[p2 * H_ + hh];
        const float d1 = pden[p1 * H_ + hh], d2 = pden[p2 * H_ + hh];
        const float mx = fmaxf(x1, x2);
        const float s1 = ex2(x1 - mx), s2 = ex2(x2 - mx);
        const float inv = 1.f / fmaf(d1, s1, d2 * s2);
        const float a1 = s1 * inv, a2 = s2 * inv;
        const float4 hv = *reinterpret_cast<const float4 *>(h + v * D_ + 4 * lane);
        float4 o;
        o.x = (hv.x + (fmaf(n1.x, a1, n2.x * a2) + bb.x)) * sc.x + sh.x;
        o.y = (hv.y + (fmaf(n1.y, a1, n2.y * a2) + bb.y)) * sc.y + sh.y;
        o.z = (hv.z + (fmaf(n1.z, a1, n2.z * a2) + bb.z)) * sc.z + sh.z;
        o.w = (hv.w + (fmaf(n1.w, a1, n2.w * a2) + bb.w)) * sc.w + sh.w;
        if (round_tf32) o = tf32_round4(o);
        *reinterpret_cast<float4 *>(h1 + v * D_ + 4 * lane) = o;
    }
}

}  // namespace

extern "C" int gnngls_gat_aggregate_csr(const int32_t *indptr, const int32_t *indices, int64_t M, const float *ft,
                                        const float *el, const float *er, const float *h, const float *gat_bias,
                                        const float *bn_scale, const float *bn_shift, float *h1, int round_tf32,
                                        void *stream) {
    GNNGLS_REQUIRE(indptr && indices && ft && el && er && h && bn_scale && bn_shift && h1, GNNGLS_ERR_BAD_ARG,
                   "null pointer argument");
    if (M <= 0) return GNNGLS_OK;
    const int64_t blocks = (M + CSR_WARPS - 1) / CSR_WARPS;
    const int64_t cap = (int64_t)gnngls::device_sm_count() * 64;
    gat_csr_kernel<<<(int)(blocks < cap ? blocks : cap), CSR_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        indptr, indices, M, ft, el, er, h, gat_bias, bn_scale, bn_shift, h1, round_tf32);
    GNNGLS_LAUNCH_OK("gat_csr_kernel");
    return GNNGLS_OK;
}

extern "C" size_t gnngls_gat_kn_workspace_bytes(int B, int n) {
    if (B <= 0 || n <= 0) return 0;
    return sizeof(float) * (size_t)B * n * n * (D_ + 2 * H_);
}

extern "C" int gnngls_gat_aggregate_kn(int B, int n, const float *ft, const float *el, const float *er,
                                       const float *h, const float *gat_bias, const float *bn_scale,
                                       const float *bn_shift, float *h1, int round_tf32, void *workspace,
                                       size_t workspace_bytes, void *stream) {
    GNNGLS_REQUIRE(ft && el && er && h && bn_scale && bn_shift && h1, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(n >= 3, GNNGLS_ERR_UNSUPPORTED, "line graph of K_n needs n >= 3 (got %d)", n);
    if (B <= 0) return GNNGLS_OK;
    const size_t smem = star_smem_bytes(n);
    GNNGLS_REQUIRE(smem <= (size_t)gnngls::device_max_optin_smem(), GNNGLS_ERR_UNSUPPORTED,
                   "n=%d: a vertex star (%zu B) does not fit shared memory; use the CSR path", n, smem);
    GNNGLS_REQUIRE((int64_t)B * n < (int64_t)1 << 31, GNNGLS_ERR_UNSUPPORTED, "B*n too large for one launch");
    GNNGLS_REQUIRE(workspace && workspace_bytes >= gnngls_gat_kn_workspace_bytes(B, n), GNNGLS_ERR_WORKSPACE,
                   "gat_kn workspace too small: need %zu bytes", gnngls_gat_kn_workspace_bytes(B, n));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *pnum = static_cast<float *>(workspace);
    float *pden = pnum + (size_t)B * n * n * D_;
    float *pmax = pden + (size_t)B * n * n * H_;
    if (smem > 48 * 1024)
        GNNGLS_CUDA_OK(cudaFuncSetAttribute(gat_kn_star_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gat_kn_star_kernel<<<B * n, STAR_THREADS, smem, st>>>(n, ft, el, er, pnum, pden, pmax);
    GNNGLS_LAUNCH_OK("gat_kn_star_kernel");
    gat_kn_combine_kernel<<<B * n, 128, 0, st>>>(n, pnum, pden, pmax, h, gat_bias, bn_scale, bn_shift, h1, round_tf32);
    GNNGLS_LAUNCH_OK("gat_kn_combine_kernel");
    return GNNGLS_OK;
}
