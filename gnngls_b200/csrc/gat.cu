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
#include <cstdint>
#include <cmath>
#include <type_traits>
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

__global__ void __launch_bounds__(CSR_WARPS * 32)
gat_csr_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, int64_t M,
               const float *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
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

// Partials, indexed [(b*n + i)*n + j] for star vertex i and destination {i,j}:
//   pnum[.,128]: sum_k w*ft   pden[.,8]: sum_k w   pmax[.,8]: log2-scaled max used for w
struct StarCtx {
    int n, i, b, hd, g, t;
    float m1, m2;
    int a1;
    const float *Fh;      // star features, offset to this warp's head
    const float *ELs, *ERs;
    int ksteps;
    float *pn, *pd, *pm;  // this star's partial rows: pnum/pden/pmax + ((b*n+i)*n) rows, plus the lane's column offset
};

// publish one destination row of this warp's head: lane t of the quad holds features {2t,2t+1} and
// {8+2t,9+2t}.  (.cg stores: the partials are consumed by another SM through L2.)
__device__ __forceinline__ void publish_row(const StarCtx &c, int j, float2 n0, float2 n1, float den, float mx) {
    if (j >= c.n || j == c.i) return;
    float *o = c.pn + j * D_;
    __stcg(reinterpret_cast<float2 *>(o), n0);
    __stcg(reinterpret_cast<float2 *>(o + 8), n1);
    if (c.t == 0) { __stcg(c.pd + j * H_, den); __stcg(c.pm + j * H_, mx); }
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

__global__ void __launch_bounds__(STAR_THREADS, 3)
gat_kn_star_kernel(int n, const float *__restrict__ ft, const float *__restrict__ el, const float *__restrict__ er,
                   float *__restrict__ pnum, float *__restrict__ pden, float *__restrict__ pmax,
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
        const int64_t prow = ((int64_t)b * n + i) * n;
        c.pn = pnum + prow * D_ + warp * 16 + 2 * c.t;
        c.pd = pden + prow * H_ + warp;
        c.pm = pmax + prow * H_ + warp;
    }
    const int MT = MP / 16;
    int mt = 0;
    for (; mt + 2 <= MT; mt += 2) star_tiles<2>(c, mt);
    if (mt < MT) star_tiles<1>(c, mt);

    // ---- every destination {i,j} belongs to the stars of i and of j.  Publish this star's partials,
    // then bump the destination's arrival counter: the star that arrives second merges the two
    // partials (flash-style rescale, always in (min(i,j), max(i,j)) order so the result does not
    // depend on arrival order) and applies bias + skip + BatchNorm1.  No separate combine pass.
    __threadfence();
    __syncthreads();
    int *second = reinterpret_cast<int *>(ELs);               // reuse: [n] flags
    for (int j = threadIdx.x; j < n; j += STAR_THREADS)
        second[j] = (j != i) ? atomicAdd(arrive + node0 + kn_node(i, j, n), 1) : 0;
    __syncthreads();
    __threadfence();
    const int hh = lane >> 2;
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
    constexpr int FIN_U = 3;
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
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            const int64_t p1 = ((int64_t)b * n + lo) * n + hi, p2 = ((int64_t)b * n + hi) * n + lo;
            v[u] = node0 + kn_node(lo, hi, n);
            n1[u] = __ldcg(reinterpret_cast<const float4 *>(pnum + p1 * D_ + 4 * lane));
            n2[u] = __ldcg(reinterpret_cast<const float4 *>(pnum + p2 * D_ + 4 * lane));
            x1[u] = __ldcg(pmax + p1 * H_ + hh); x2[u] = __ldcg(pmax + p2 * H_ + hh);
            d1[u] = __ldcg(pden + p1 * H_ + hh); d2[u] = __ldcg(pden + p2 * H_ + hh);
            hv[u] = *reinterpret_cast<const float4 *>(h + v[u] * D_ + 4 * lane);
            jj[u] = j;
        }
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) {
            if (!ok[u]) continue;
            // both partial rows are dead now (each is read exactly once): drop their dirty L2 lines instead of
            // writing 1 KB per destination back to HBM
            if (lane < 8) {
                const int a = lane < 4 ? i : jj[u], c2 = lane < 4 ? jj[u] : i;      // rows (star a, destination c2) of pnum
                discard_l2_128(pnum + (((int64_t)b * n + a) * n + c2) * D_ + (lane & 3) * 32);
            }
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

}  // namespace

extern "C" int gnngls_gat_aggregate_csr(const int32_t *indptr, const int32_t *indices, int64_t M, const float *ft,
                                        const float *el, const float *er, const float *h, const float *gat_bias,
                                        const float *bn_scale, const float *bn_shift, float *h1, float *h1_tf32,
                                        void *stream) {
    GNNGLS_REQUIRE(indptr && indices && ft && el && er && h && bn_scale && bn_shift && h1, GNNGLS_ERR_BAD_ARG,
                   "null pointer argument");
    if (M <= 0) return GNNGLS_OK;
    const int64_t blocks = (M + CSR_WARPS - 1) / CSR_WARPS;
    const int64_t cap = (int64_t)gnngls::device_sm_count() * 64;
    gat_csr_kernel<<<(int)(blocks < cap ? blocks : cap), CSR_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        indptr, indices, M, ft, el, er, h, gat_bias, bn_scale, bn_shift, h1, h1_tf32);
    GNNGLS_LAUNCH_OK("gat_csr_kernel");
    return GNNGLS_OK;
}

extern "C" size_t gnngls_gat_kn_workspace_bytes(int B, int n) {
    if (B <= 0 || n <= 0) return 0;
    const size_t N = (size_t)n * (n - 1) / 2;
    return sizeof(float) * (size_t)B * n * n * (D_ + 2 * H_) + sizeof(int) * (size_t)B * N;
}

extern "C" int gnngls_gat_aggregate_kn(int B, int n, const float *ft, const float *el, const float *er,
                                       const float *h, const float *gat_bias, const float *bn_scale,
                                       const float *bn_shift, float *h1, float *h1_tf32, int ft_is_tf32,
                                       void *workspace, size_t workspace_bytes, void *stream) {
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
    const size_t N = (size_t)n * (n - 1) / 2;
    float *pnum = static_cast<float *>(workspace);
    float *pden = pnum + (size_t)B * n * n * D_;
    float *pmax = pden + (size_t)B * n * n * H_;
    int *arrive = reinterpret_cast<int *>(pmax + (size_t)B * n * n * H_);
    GNNGLS_CUDA_OK(cudaMemsetAsync(arrive, 0, sizeof(int) * (size_t)B * N, st));
    if (smem > 48 * 1024)
        GNNGLS_CUDA_OK(cudaFuncSetAttribute(gat_kn_star_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gat_kn_star_kernel<<<B * n, STAR_THREADS, smem, st>>>(n, ft, el, er, pnum, pden, pmax, arrive, h, gat_bias,
                                                          bn_scale, bn_shift, h1, h1_tf32, ft_is_tf32);
    GNNGLS_LAUNCH_OK("gat_kn_star_kernel");
    return GNNGLS_OK;
}
