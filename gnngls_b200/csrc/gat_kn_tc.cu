// Tensor-core (tcgen05) variant of the K_n GAT aggregate for fp16 features and n <= 128 (TSP20/50/100).
// Same operation as gat_kn.cu (dgl.nn.GATConv reached from gnngls/models.py:23; skip +
// BatchNorm1 of models.py:12-15,27 fused), different way of forming a star's partial sums.
//
// With s = el_k + er_j (log2 domain), 2^leaky_relu(s) is  A_k * C1_j  when el_k >= -er_j  and  A'_k * C2_j  otherwise
// (A_k = 2^(el_k-ref), A'_k = 2^(.2(el_k-ref)), ref = max_k el_k; C1_j, C2_j as in gat_kn.cu).  So for one head of one
// star, with the 0/1 indicator matrix I[j][k] = [el_k >= -er_j]  (destinations j x members k),
//
//     [ SA | SB | dA dB ] = I  x  [ A_k ft_k | A'_k ft_k | A_k  A'_k ]            one tcgen05.mma chain, M=128 N=48
//     num_j = C1_j SA_j + C2_j (TotB - SB_j) - self,     TotB = row of the all-ones indicator (the vertex's own, unused row)
//
// The per-(j,k) work is ONE half2 compare per PAIR of weights: `set.ge.f16x2` yields the fp16 values 1.0 / 0.0 that
// are the MMA's A operand, written straight into tensor memory (tcgen05.st, lane = destination row).  The B operand
// is built once per member in shared memory (MN-major canonical layout, no swizzle: core matrix = 8 members x 8
// columns), accumulators live in tensor memory and come back one destination row per thread (tcgen05.ld).
// No per-edge exponential, no per-edge multiply, no sort.
// Rounding: A_k ft_k is rounded to fp16 (the features already are fp16; same error class as fp16 attention weights).
// The branch decision compares fp16 roundings of the CENTRED scores el_k - ref and -er_j - ref: a member within one
// fp16 ulp of the threshold may take the other branch, which changes its weight by 2^(0.8|s|) with |s| below that ulp;
// because the ulp is relative to the distance d from the star's maximum and the weight is <= 2^-d of the row's
// maximum, the error is below 3e-4 of the row's largest weight whatever the scores' magnitude.
// When the star's largest score leads the runner-up by more than 6 (log2 units) its member is taken out of the MMA
// (ref = runner-up) and added in fp32, so the self-exclusion of its own row never cancels a dominant term.
//
// Persistent kernel, 128 threads per CTA (thread = member k when building operands, = destination row j afterwards), four
// CTAs per SM, each with its own 128 tensor-memory columns; CTA c takes the stars (instance b, vertex i) number c, c + G, ...
// Destination {i,j} belongs to two stars.  Each star writes its partial (16 numerators, denominator, reference max per
// head) for every row to the record buffer -- slot 0 if it is the star of the lower vertex, slot 1 otherwise -- and
// nothing in a star's work depends on any other star.  An instance whose n stars are all out (a counter per instance)
// is merged D instances later, a slice of (n-1)/2 consecutive nodes per star slot: the two records are combined in fixed
// (lower, higher) order -- deterministic and batching-invariant bitwise -- bias + skip + BN1 are applied and h1 is written
// with full lines; the consumed records are dropped from L2 so they never travel to HBM.  The merge rows of a CTA are
// spread over the shadows of its eight MMA chains (one per head), the only places where its warps would otherwise wait.
#include "gat_kn.cuh"

// Phase timing (debug builds only: GNNGLS_KN_STAMPS=1 python -m gnngls_b200.build --force): lane 0 of every warp accumulates
// clock64() differences per phase; read back with gnngls_debug_kn_stamps() (tools/kn_stamps.py).
#ifdef KN_STAMPS
__device__ unsigned long long g_kn_stamps[148 * 4 * 4 * 16];
#define KN_STAMP(slot)                                              \
    do {                                                            \
        if (lane == 0) {                                            \
            const long long now__ = clock64();                      \
            stamp_acc[slot] += (unsigned long long)(now__ - stamp_last); \
            stamp_last = now__;                                     \
        }                                                           \
    } while (0)
#else
#define KN_STAMP(slot) do { } while (0)
#endif

namespace {

constexpr int T_THREADS = 128, T_WARPS = 4, KPAD = 128;
constexpr int XN = 48;                                   // MMA N: 16 (A-branch) + 16 (B-branch) + 2 denominators, padded to 16s
constexpr int X_KB = (XN / 8) * 128;                     // bytes of one 8-member block of the B operand (6 core matrices)
constexpr int X_BYTES = (KPAD / 8) * X_KB;               // 12 KB per team
constexpr float kLeadGap = 6.f;                          // lead (log2 units) of the largest score above which its member is handled in fp32
constexpr int TMEM_COLS = 128;                           // 64 columns indicator (K=128 as fp16 pairs) + 48 accumulator
// Partial records.  The 16 numerators of a (row, head) travel as int16 with one shared scale (round 2: fp32):
// a record shrinks from 576 to 320 bytes per node and star, so the records in flight between a star and its merge (two to three
// rounds of the grid) are half as likely to be evicted from L2 to HBM, and the epilogue writes one 32-byte store per head instead
// of two.  The scale costs no storage: a partial (v, den, M) -- numerators, denominator, reference exponent -- means the same as
// (v c, den c, M - log2 c) for any c > 0, so the epilogue picks c = 2^-e (1 - 2^-15) with max|v| c in [2^14, 2^15), rounds v c to
// integers and stores den c and M - log2 c in the fp32 part of the record; the merge is unchanged.  Absolute error <= 2^-15 of the
// head's largest numerator (fp16 would give 2^-11 of every value: 4 x the kernel's whole error, measured).
// One record per node holds BOTH stars' partials, so that the merge fetches it (and the next node's) with one bulk copy:
//   [slot 0: 128 int16][slot 1: 128 int16][slot 0: 8 x (den, exponent) fp32][slot 1: 8 x (den, exponent) fp32]  = 640 bytes
constexpr int REC_BYTES = 2 * 256 + 2 * 64;
constexpr int MROW = (REC_BYTES + 512) / 4;              // floats per merge row in flight: the node's record + its skip row (128 fp32)
constexpr int ESTR = 20;                                 // floats per member in the score buffer: el (8), er (8), pad (bank spread)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "KN_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra.uni KN_WAIT_DONE;\n"
        "bra.uni KN_WAIT_LOOP;\n"
        "KN_WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
// (no wait: the caller issues tcgen05.wait::ld once after its last load)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t &r0, uint32_t &r1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
}
// D[tmem] (+)= A[tmem] * B[smem], kind::f16: A lane = row, one 32-bit column per pair of k; B through its descriptor
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// MN-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): element (k, n) of B at
// (k/8)*LBO + (n/8)*SBO + (k%8)*16 + (n%8)*2 bytes -- checked on B200 by tools/umma_mn_test.cu
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
// F32 accumulate, fp16 x fp16, A from tensor memory (K-major), B MN-major (bit 16), N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 16) | ((uint32_t)(XN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// L2 cache-policy hints: the records must survive in L2 until they are merged, everything that streams must not push them out
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint4(float4 *p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// 256-bit global accesses: one full 32-byte sector per lane and instruction (a lane's 64-byte chunk of a record or of a feature
// row costs half the LSU wavefronts of 128-bit accesses -- the L1 data pipe is this kernel's scarcest resource)
__device__ __forceinline__ void st_keep8(float *p, float4 a, float4 b) {
    asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)),
                 "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)),
                 "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w)) : "memory");
}
__device__ __forceinline__ void st_keep8u(void *p, const uint32_t (&w)[8]) {
    asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void ldg8(const void *p, uint4 &a, uint4 &b) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
// one piece of a merge row through the bulk-copy engine (global -> shared): no LSU wavefronts, completion counted in bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

struct Top2 { float m1, m2; int a1; };
__device__ __forceinline__ Top2 top2_merge(Top2 x, Top2 y) {
    Top2 r;
    if (x.m1 >= y.m1) { r.m1 = x.m1; r.a1 = x.a1; r.m2 = fmaxf(x.m2, y.m1); }
    else { r.m1 = y.m1; r.a1 = y.a1; r.m2 = fmaxf(y.m2, x.m1); }
    return r;
}
__device__ __forceinline__ void unpack8(const uint4 q, float (&f)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&w[u]));
        f[2 * u] = a.x; f[2 * u + 1] = a.y;
    }
}

struct TLayout {
    unsigned eler_off, x_off, elh_off, hd_off, tot_off, mland_off, si_off, dm_off, bar_off, total;
    __host__ __device__ TLayout(int n, bool pair) {
        eler_off = 0;                                        // [2][1 or 2 stars][n][ESTR] fp32 scores of the current / next star(s)
        x_off = (2u * (pair ? 2u : 1u) * (unsigned)n * ESTR * 4u + 127u) & ~127u;   // [X_BYTES] B operand
        elh_off = x_off + X_BYTES;                           // [8 heads][128] fp16 centred scores el - ref
        hd_off = elh_off + 8u * 256u;                        // [8 heads][1 or 2 stars][4] words: ref, m1, leading member, "leading member handled in fp32"
        tot_off = hd_off + 16u * 16u;                         // [52] floats: TotB (16), total dB, pad, fp32 features of the leading member (2 x 16, head parity)
        mland_off = (tot_off + 2u * 56u * 4u + 15u) & ~15u;       // [4 warps][2 rows][MROW] floats: records + skip row of the merge rows in flight
        si_off = mland_off + T_WARPS * 2u * MROW * 4u;       // slot handed out for the star after next
        dm_off = si_off + 16u;                               // [8 heads][128 rows] (denominator, max) of the star, written out once per star
        bar_off = dm_off + 8u * 128u * 8u;                   // MMA mbarrier + tmem slot + 4 merge-row mbarriers
        total = bar_off + 16u + 4u * 8u;
    }
};

// first node of vertex i's run in the sorted-tuple order: node {i,k}, i < k, is tri(i) + k - i - 1
__device__ __forceinline__ int tri(int i, int n) { return (i * (2 * n - i - 1)) >> 1; }
__device__ __forceinline__ int kn_local(int i, int k, int n) { return i < k ? tri(i, n) + k - i - 1 : tri(k, n) + i - k - 1; }

// PAIR (n <= 64): two consecutive stars share a CTA iteration -- rows / members 0..63 belong to the first, 64..127 to the second;
// the indicator is block diagonal (a row's two off-diagonal chunks are zeroed once and never written again), so one MMA chain serves both.
template <bool PAIR>
__global__ void __launch_bounds__(T_THREADS, 4) gat_kn_tc_kernel(const KnArgs a, const int B) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = a.n;
    const TLayout L(n, PAIR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, tt = tid;
    const int sub = PAIR ? tid >> 6 : 0, lt = PAIR ? tid & 63 : tid;      // which star of the pair; member / row within it
    constexpr int NS = PAIR ? 2 : 1;
    const size_t N = (size_t)n * (n - 1) / 2;

    float *ELER = reinterpret_cast<float *>(smem + L.eler_off);
    unsigned char *Xs = smem + L.x_off;
    __half *ELHall = reinterpret_cast<__half *>(smem + L.elh_off);
    float *HD = reinterpret_cast<float *>(smem + L.hd_off);
    float *TOT = reinterpret_cast<float *>(smem + L.tot_off);
    float *MLAND = reinterpret_cast<float *>(smem + L.mland_off) + warp * 2 * MROW;
    int *SI = reinterpret_cast<int *>(smem + L.si_off);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.bar_off);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bar + 1);
    uint64_t *mbar = bar + 2 + warp;                                   // this warp's merge-row barrier
    float2 *DMS = reinterpret_cast<float2 *>(smem + L.dm_off);

    // ---------------------------------------------------------------- setup
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        for (int w = 0; w < T_WARPS; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 2 + w)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tslot;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;            // this warp's lane quarter
    const int nk = PAIR ? KPAD / 16 : (n + 15) >> 4, nch_full = nk >> 1;   // MMAs (16 members each) / whole indicator chunks (32 members each)
    unsigned char *xrow = Xs + (tt >> 3) * X_KB + (tt & 7) * 16;       // this member's row of the B operand (pieces 128 B apart)
    const int total = B * n;                                           // stars = merge slices
    const int slots = PAIR ? (total + 1) >> 1 : total;                 // CTA iterations with a star (pair)
    if (PAIR) {                                                        // the off-diagonal indicator blocks: zero, once
        const uint32_t z[16] = {};
        for (int c = 0; c < 4; ++c) tmem_st16(tbase + lane_sel + c * 16, z);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

    // scores of star (sb, si) -> buffer `buf` (cp.async; the vertex's own slot is zero-filled)
    auto issue_scores = [&](int slot, int buf) {
#pragma unroll
        for (int sp = 0; sp < NS; ++sp) {
            const int st = NS * slot + sp;
            if (st >= total) break;
            const int sb = st / n, si = st - sb * n;
            for (int idx = tid; idx < 4 * n; idx += T_THREADS) {
                const int k = idx >> 2, p = idx & 3;
                float *dst = ELER + ((buf * NS + sp) * n + k) * ESTR + p * 4;
                if (k != si) cp_async16(dst, (p < 2 ? a.el : a.er) + ((size_t)sb * N + kn_local(si, k, n)) * H_ + (p & 1) * 4);
                else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };

    const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    int *ctr = a.flags + B, *mctr = ctr + 1;                           // star-slot / merge-slice counters (zeroed with the per-instance counters)
    if (tid == 0) {
        SI[0] = atomicAdd(ctr, 1);
        SI[1] = atomicAdd(ctr, 1);
    }
    __syncthreads();
    int cur = SI[0], nxt = SI[1];                                      // star slots are handed out in order, one star ahead
    int cand = -1;                                                     // merge slice taken by this CTA, not merged yet (-1: none)
    __syncthreads();
    if (cur < slots) issue_scores(cur, 0);
    cp_async_commit();
#ifdef KN_STAMPS
    unsigned long long stamp_acc[16] = {};
    long long stamp_last = clock64();
#endif
    uint32_t parity = 0, mparity = 0;
    // BN1 scale / shift and the GAT bias of this lane's four features (merge rows: one warp per row)
    const float4 sc4 = __ldg(reinterpret_cast<const float4 *>(a.bn_scale) + lane), sh4 = __ldg(reinterpret_cast<const float4 *>(a.bn_shift) + lane);
    const float4 bb4 = a.bias ? __ldg(reinterpret_cast<const float4 *>(a.bias) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);

    for (int it = 0; cur < slots || cand >= 0 || it == 0; ++it) {
        const int cbuf = it & 1;
        const bool havestar = cur < slots;                             // (the CTA has a star, or a pair, in this iteration)
        const int star = NS * cur + sub;                               // this thread's star
        const bool mystar = havestar && star < total;
        const int b = mystar ? star / n : 0, i = mystar ? star - b * n : 0;
        int grabbed = 0;
        if (tid == 0) {
            grabbed = atomicAdd(ctr, 1);                               // the slot after next (needed at the end of the iteration)
            // Merge slices (instance bm, nodes [mlo, mhi)) are handed out in order too, but a CTA only merges its slice once all n
            // stars of that instance are out: a look at the top of every iteration, never a wait while a star is pending.
            int rdy = 0;
            if (cand >= 0) {
                rdy = ld_acquire_gpu(a.flags + cand / n) >= n;
                if (!rdy && !havestar) __nanosleep(1000);              // (only merging left: do not hammer the counter)
            }
            SI[2] = rdy;
        }
        // this thread as member / destination row tt of the current star; its features for the first head
        const bool live = mystar && lt < n && lt != i;
        const int my_local = live ? kn_local(i, lt, n) : 0;
        const size_t my_node = (size_t)b * N + my_local;
        const uint4 *ftrow = reinterpret_cast<const uint4 *>(static_cast<const unsigned char *>(a.ft) + my_node * 256);
        uint4 f0 = make_uint4(0u, 0u, 0u, 0u), f1 = f0;
        if (live) ldg8(ftrow, f0, f1);
        KN_STAMP(0);                                                   // top: bookkeeping, feature loads, readiness look
        cp_async_wait_group<0>();                                      // the current star's scores (requested a star ago)
        __syncthreads();
        KN_STAMP(1);                                                   // scores + barrier
        const int mcur = SI[2] ? cand : -1;                            // the slice merged in this iteration
        const bool havemerge = mcur >= 0;
        const int bm = havemerge ? mcur / n : 0, im = mcur - bm * n;
        const int mlo = (im * (n - 1)) >> 1, mhi = havemerge ? ((im + 1) * (n - 1)) >> 1 : 0;
        const size_t mnode0 = (size_t)bm * N;
        int mtake = cand;
        if (tid == 0 && (havemerge || cand < 0)) mtake = atomicAdd(mctr, 1);   // the next slice (needed at the end of the iteration)
        if (havemerge) {
            // skip rows of the merge slice: start them on their way into L2, they are fetched one MMA shadow before they are used
            for (int idx = tid; idx < (mhi - mlo) * 4; idx += T_THREADS)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.h + (mnode0 + mlo) * D_ + 32 * idx));
        }
        if (nxt < slots) issue_scores(nxt, cbuf ^ 1);
        // merge rows of MMA shadow `hs` (this warp: the two CONSECUTIVE rows mlo + 8 hs + 2 warp, + 1): their records are adjacent in
        // memory and so are their skip rows -- two bulk copies per warp and shadow land [record, record, skip, skip] in shared memory
        auto fetch_merge_rows = [&](int hs) {
            const int r0 = mlo + hs * 2 * T_WARPS + 2 * warp;
            if (lane == 0 && r0 < mhi) {
                const uint32_t rows = (r0 + 1 < mhi) ? 2u : 1u;
                const uint32_t mb = smem_u32(mbar);
                asm volatile("fence.proxy.async.global;" ::: "memory");    // the records were made visible to the generic proxy
                mbar_arrive_expect_tx(mb, rows * (REC_BYTES + 512));
                const size_t node = mnode0 + r0;
                const uint32_t dst = smem_u32(MLAND);
                bulk_g2s(dst, reinterpret_cast<const unsigned char *>(a.recV) + node * REC_BYTES, rows * REC_BYTES, mb, pol_stream);
                bulk_g2s(dst + 2 * REC_BYTES, a.h + node * D_, rows * 512, mb, pol_stream);
            }
        };
        cp_async_commit();
        if (havemerge) fetch_merge_rows(0);
        const float *E = ELER + (cbuf * NS + sub) * n * ESTR;          // this thread's star
        if (havestar) {
            // ---- top-2 of el per (head, star): each warp takes 8 NS / 4 of them; the fp16 copy of the CENTRED scores that the branch
            // decision uses, at member position 64 sp + k
#pragma unroll
            for (int hq = 0; hq < H_ * NS / T_WARPS; ++hq) {
                const int combo = warp * (H_ * NS / T_WARPS) + hq, head = combo % H_, sp = combo / H_;
                const int st = NS * cur + sp, ip = st < total ? st % n : -1;    // (ip < 0: no such star -- an odd star count)
                const float *Ep = ELER + (cbuf * NS + sp) * n * ESTR;
                constexpr int NQ = PAIR ? 2 : KPAD / 32;
                float ev[NQ];
                Top2 t2{-INFINITY, -INFINITY, 0};
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int k = lane + 32 * q;
                    const bool ok = ip >= 0 && k < n && k != ip;
                    ev[q] = ok ? Ep[k * ESTR + head] : -INFINITY;
                    if (ok) t2 = top2_merge(t2, Top2{ev[q], -INFINITY, k});
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    Top2 o;
                    o.m1 = __shfl_xor_sync(0xffffffffu, t2.m1, off);
                    o.m2 = __shfl_xor_sync(0xffffffffu, t2.m2, off);
                    o.a1 = __shfl_xor_sync(0xffffffffu, t2.a1, off);
                    t2 = top2_merge(t2, o);
                }
                const bool fix = ip >= 0 && t2.m1 - t2.m2 > kLeadGap;  // (n >= 3: the runner-up exists)
                const float ref = ip < 0 ? 0.f : (fix ? t2.m2 : t2.m1);
#pragma unroll
                for (int q = 0; q < NQ; ++q) ELHall[head * KPAD + (PAIR ? 64 * sp : 0) + lane + 32 * q] = __float2half_rn(ev[q] - ref);
                if (lane == 0)
                    *reinterpret_cast<float4 *>(HD + (head * NS + sp) * 4) = make_float4(ref, t2.m1, __int_as_float(t2.a1), __int_as_float(fix ? 1 : 0));
            }
            KN_STAMP(2);                                               // slice setup, top-2
            __syncthreads();
            KN_STAMP(3);                                               // barrier
        }

        // the merge rows fetched one shadow ago: combine the two stars' records in fixed (lower, higher) order, apply bias + skip +
        // BN1, write h1; then fetch the rows of the next shadow
        auto merge_rows = [&](int hs) {
            if (mlo + hs * 2 * T_WARPS + 2 * warp < mhi) {             // (warp-uniform) rows were requested for this shadow
                mbar_wait(mbar, mparity);
                mparity ^= 1;
            }
            KN_STAMP(15);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = mlo + hs * 2 * T_WARPS + 2 * warp + q;
                if (r >= mhi) break;
                const size_t node = mnode0 + r;
                const unsigned char *Lr = reinterpret_cast<const unsigned char *>(MLAND) + q * REC_BYTES;
                const uint2 lr = *reinterpret_cast<const uint2 *>(Lr + 8 * lane), ur = *reinterpret_cast<const uint2 *>(Lr + 256 + 8 * lane);
                auto lo16 = [](uint32_t w) { return (float)((int)(w << 16) >> 16); };
                auto hi16 = [](uint32_t w) { return (float)((int)w >> 16); };
                const float4 lv = make_float4(lo16(lr.x), hi16(lr.x), lo16(lr.y), hi16(lr.y));
                const float4 uv = make_float4(lo16(ur.x), hi16(ur.x), lo16(ur.y), hi16(ur.y));
                const float2 l2 = *reinterpret_cast<const float2 *>(Lr + 512 + 8 * (lane >> 2)), u2 = *reinterpret_cast<const float2 *>(Lr + 576 + 8 * (lane >> 2));
                const float4 hv = *reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(MLAND) + 2 * REC_BYTES + q * 512 + 16 * lane);
                const float mx = fmaxf(l2.y, u2.y);
                const float s1 = ex2(l2.y - mx), s2 = ex2(u2.y - mx);
                const float inv = rcp_approx(fmaf(l2.x, s1, u2.x * s2));
                const float a1 = s1 * inv, a2 = s2 * inv;
                // (packed fp32 pairs; same operations and order as  (h + ((l a1 + u a2) + bias)) scale + shift -- the product-sum of the last
                //  step is a fused multiply-add exactly as nvcc contracts the scalar expression)
                const float2 a1p = make_float2(a1, a1), a2p = make_float2(a2, a2);
                const float2 g01 = __ffma2_rn(make_float2(lv.x, lv.y), a1p, __fmul2_rn(make_float2(uv.x, uv.y), a2p));
                const float2 g23 = __ffma2_rn(make_float2(lv.z, lv.w), a1p, __fmul2_rn(make_float2(uv.z, uv.w), a2p));
                const float2 y01 = __fadd2_rn(make_float2(hv.x, hv.y), __fadd2_rn(g01, make_float2(bb4.x, bb4.y)));
                const float2 y23 = __fadd2_rn(make_float2(hv.z, hv.w), __fadd2_rn(g23, make_float2(bb4.z, bb4.w)));
                const float2 o01 = __ffma2_rn(y01, make_float2(sc4.x, sc4.y), make_float2(sh4.x, sh4.y));
                const float2 o23 = __ffma2_rn(y23, make_float2(sc4.z, sc4.w), make_float2(sh4.z, sh4.w));
                const float4 o = make_float4(o01.x, o01.y, o23.x, o23.y);
                st_hint4(reinterpret_cast<float4 *>(a.h1 + node * D_) + lane, o, pol_stream);
                if (a.h1_tf32) st_hint4(reinterpret_cast<float4 *>(a.h1_tf32 + node * D_) + lane, tf32_round4(o), pol_stream);
                // the consumed records are dead (read exactly once): drop their dirty L2 lines instead of writing them back
                if (lane < REC_BYTES / 128) discard_l2_128(reinterpret_cast<const unsigned char *>(a.recV) + node * REC_BYTES + 128 * lane);
            }
            __syncwarp();
            if (hs + 1 < H_) fetch_merge_rows(hs + 1);
        };

#pragma unroll 1
        for (int head = 0; head < H_; ++head) {
            // (per head and star: ref, m1, leading member, "leading member in fp32"; meaningless without a star)
            const float4 hd = havestar ? *reinterpret_cast<const float4 *>(HD + (head * NS + sub) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float ref = hd.x;
            const bool fix = __float_as_int(hd.w) != 0;                // uniform over the star's threads (whole warps)
            const bool lead = fix && lt == __float_as_int(hd.z);
            const float erh = live ? E[lt * ESTR + 8 + head] : 0.f;
            const __half th16 = live ? __float2half_rn(-erh - ref) : __float2half_rn((mystar && lt == i) ? -INFINITY : INFINITY);
            if (havestar) {
                float *TOTs = TOT + sub * 56;                          // this star's column totals / leading-member features
                float *TOTl = TOTs + 20 + (head & 1) * 16;             // fp32 features of the leading member (parity buffer)
                // ---- B operand row of member tt: [A ft | A' ft | A A' 0...]
                if (tt < 16 * nk) {
                    uint4 xa0 = make_uint4(0u, 0u, 0u, 0u), xa1 = xa0, xb0 = xa0, xb1 = xa0, xd = xa0;
                    if (live && !lead) {
                        const float d = E[lt * ESTR + head] - ref;
                        const __half A16 = __float2half_rn(ex2(d)), A516 = __float2half_rn(ex2(kSlope * d));
                        const __half2 hA = __half2half2(A16), hB = __half2half2(A516);
                        auto mul4 = [](uint4 f, __half2 s) {
                            uint4 r;
                            __half2 t;
                            t = __hmul2(*reinterpret_cast<const __half2 *>(&f.x), s); r.x = *reinterpret_cast<uint32_t *>(&t);
                            t = __hmul2(*reinterpret_cast<const __half2 *>(&f.y), s); r.y = *reinterpret_cast<uint32_t *>(&t);
                            t = __hmul2(*reinterpret_cast<const __half2 *>(&f.z), s); r.z = *reinterpret_cast<uint32_t *>(&t);
                            t = __hmul2(*reinterpret_cast<const __half2 *>(&f.w), s); r.w = *reinterpret_cast<uint32_t *>(&t);
                            return r;
                        };
                        xa0 = mul4(f0, hA); xa1 = mul4(f1, hA); xb0 = mul4(f0, hB); xb1 = mul4(f1, hB);
                        const __half2 dd = __halves2half2(A16, A516);
                        xd.x = *reinterpret_cast<const uint32_t *>(&dd);
                    }
                    *reinterpret_cast<uint4 *>(xrow) = xa0;
                    *reinterpret_cast<uint4 *>(xrow + 128) = xa1;
                    *reinterpret_cast<uint4 *>(xrow + 256) = xb0;
                    *reinterpret_cast<uint4 *>(xrow + 384) = xb1;
                    *reinterpret_cast<uint32_t *>(xrow + 512) = xd.x;  // (columns 34..47 of the operand are never read back: left as they are)
                }
                if (fix) {                                             // (rare) the leading member of this head: fp32 features for the epilogue
                    if (lead) {
                        float ff[16];
                        unpack8(f0, reinterpret_cast<float(&)[8]>(ff[0]));
                        unpack8(f1, reinterpret_cast<float(&)[8]>(ff[8]));
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<float4 *>(TOTl + 4 * q) = make_float4(ff[4 * q], ff[4 * q + 1], ff[4 * q + 2], ff[4 * q + 3]);
                    }
                    __syncwarp();
                }
                KN_STAMP(14);
                if (live && head + 1 < H_) ldg8(ftrow + 2 * (head + 1), f0, f1);   // next head's features
                // ---- indicator row of destination tt: I[tt][k] = [el_k - ref >= -er_tt - ref] as fp16 1.0 / 0.0, straight into tensor
                // memory.  Row i (the vertex itself, not a destination) takes threshold -inf: its accumulator row is the column totals.
                {
                    const __half2 th2 = __half2half2(th16);
                    const __half *ELH = ELHall + head * KPAD;
                    const int c_lo = PAIR ? 2 * sub : 0, c_hi = PAIR ? 2 * sub + 2 : nch_full;   // (PAIR: own star's chunks only)
                    for (int c = c_lo; c < c_hi; ++c) {
                        uint32_t v[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 e = *reinterpret_cast<const uint4 *>(ELH + c * 32 + q * 8);     // 4 pairs of scores (broadcast)
                            const uint32_t w[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const __half2 r = __hge2(*reinterpret_cast<const __half2 *>(&w[u]), th2);
                                v[q * 4 + u] = *reinterpret_cast<const uint32_t *>(&r);
                            }
                        }
                        tmem_st16(tbase + lane_sel + c * 16, v);
                    }
                    if (!PAIR && (nk & 1)) {                           // an odd number of 16-member MMA steps: half a chunk more
                        uint32_t v[8];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint4 e = *reinterpret_cast<const uint4 *>(ELH + nch_full * 32 + q * 8);
                            const uint32_t w[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const __half2 r = __hge2(*reinterpret_cast<const __half2 *>(&w[u]), th2);
                                v[q * 4 + u] = *reinterpret_cast<const uint32_t *>(&r);
                            }
                        }
                        tmem_st8(tbase + lane_sel + nch_full * 16, v);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // B operand written through the generic proxy
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                KN_STAMP(5);                                           // wait::st + fences
                __syncthreads();                                       // operands complete
                KN_STAMP(6);                                           // barrier A
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tid == 0) {
                    for (int ks = 0; ks < nk; ++ks)
                        umma_f16_ts(tbase + 64, tbase + ks * 8, make_mn_desc(smem_u32(Xs) + ks * 2 * X_KB, X_KB, 128), kIdesc, ks != 0);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
                }
                __syncwarp();
            }
            KN_STAMP(7);                                               // MMA issue
            if (havemerge) merge_rows(head);                           // in the shadow of the MMA chain
            KN_STAMP(8);                                               // merge rows
            if (havestar) {
                // ---- accumulators
                uint32_t SA[16], SB[16], SD0, SD1;
                mbar_wait(bar, parity);
                parity ^= 1;
                KN_STAMP(9);                                           // rest of the MMA
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_ld16(tbase + lane_sel + 64, SA);
                tmem_ld16(tbase + lane_sel + 80, SB);
                tmem_ld2(tbase + lane_sel + 96, SD0, SD1);             // sum A, sum A' over the A-branch members
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                float *TOTs = TOT + sub * 56;
                if (mystar && lt == i) {                               // the all-ones row of this star: column totals
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4 *>(TOTs + 4 * q) = make_uint4(SB[4 * q], SB[4 * q + 1], SB[4 * q + 2], SB[4 * q + 3]);
                    TOTs[16] = __uint_as_float(SD1);
                }
                KN_STAMP(10);                                          // accumulators -> registers
                __syncthreads();                                       // totals visible; every row has its accumulators
                KN_STAMP(11);                                          // barrier B
                if (live) {
                    // ---- this star's partial for destination tt (fp32): v = C1 SA + C2 (Tot - SB) - self (+ leading member)
                    const float s = ref + erh;
                    const float c = ex2(-0.8f * fabsf(s));
                    const float C1 = s >= 0.f ? 1.f : c, C2 = s >= 0.f ? c : 1.f;
                    float M = lrelu(s);
                    const bool self_a = __hge(ELHall[head * KPAD + tt], th16);     // the branch the MMA put this row's own member in
                    const uint32_t xdw = *reinterpret_cast<const uint32_t *>(xrow + 512);
                    const float2 aa = __half22float2(*reinterpret_cast<const __half2 *>(&xdw));   // (A, A') as the MMA saw them
                    float den = fmaf(C1, __uint_as_float(SD0), C2 * (TOTs[16] - __uint_as_float(SD1))) - (self_a ? C1 * aa.x : C2 * aa.y);
                    float scl = 1.f;
                    const bool addlead = fix && !lead;
                    if (addlead) {                                     // the leading member joins in fp32 with weight exactly 1
                        const float Mrow = lrelu(hd.y + erh);
                        scl = ex2(M - Mrow);
                        den = fmaf(den, scl, 1.f);
                        M = Mrow;
                    }
                    const float k1 = C1 * scl, k2 = C2 * scl, ks = (self_a ? C1 : C2) * scl;
                    const unsigned char *xself = xrow + (self_a ? 0 : 256);        // this member's own products in the branch it was counted in
                    const int sl = i < lt ? 0 : 1;                     // slot 0: written by the star of the lower vertex
                    const float *TOTl = TOTs + 20 + (head & 1) * 16;
                    const uint4 xq0 = *reinterpret_cast<const uint4 *>(xself), xq1 = *reinterpret_cast<const uint4 *>(xself + 128);
                    const uint32_t xsw[8] = {xq0.x, xq0.y, xq0.z, xq0.w, xq1.x, xq1.y, xq1.z, xq1.w};
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {                      // 4 features at a time
                        const float4 t4 = *reinterpret_cast<const float4 *>(TOTs + 4 * q);
                        const uint32_t xs0 = xsw[2 * q], xs1 = xsw[2 * q + 1];
                        const float2 x01 = __half22float2(*reinterpret_cast<const __half2 *>(&xs0)), x23 = __half22float2(*reinterpret_cast<const __half2 *>(&xs1));
                        // packed fp32 pairs (FFMA2: two fmas per issue slot -- the kernel is issue-bound, not FMA-bound); same operations,
                        // same order, same rounding as the scalar form  k1 SA + (-k2 SB + (-ks x + k2 t))
                        const float2 k1p = make_float2(k1, k1), nk2p = make_float2(-k2, -k2), nksp = make_float2(-ks, -ks), k2p = make_float2(k2, k2);
                        const float2 sa01 = make_float2(__uint_as_float(SA[4 * q]), __uint_as_float(SA[4 * q + 1]));
                        const float2 sa23 = make_float2(__uint_as_float(SA[4 * q + 2]), __uint_as_float(SA[4 * q + 3]));
                        const float2 sb01 = make_float2(__uint_as_float(SB[4 * q]), __uint_as_float(SB[4 * q + 1]));
                        const float2 sb23 = make_float2(__uint_as_float(SB[4 * q + 2]), __uint_as_float(SB[4 * q + 3]));
                        const float2 r01 = __ffma2_rn(k1p, sa01, __ffma2_rn(nk2p, sb01, __ffma2_rn(nksp, x01, __fmul2_rn(k2p, make_float2(t4.x, t4.y)))));
                        const float2 r23 = __ffma2_rn(k1p, sa23, __ffma2_rn(nk2p, sb23, __ffma2_rn(nksp, x23, __fmul2_rn(k2p, make_float2(t4.z, t4.w)))));
                        v[q] = make_float4(r01.x, r01.y, r23.x, r23.y);
                    }
                    if (fix) {                                         // (rare; a branch, not predicated instructions)
                        if (addlead) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 l4 = *reinterpret_cast<const float4 *>(TOTl + 4 * q);
                                v[q].x += l4.x; v[q].y += l4.y; v[q].z += l4.z; v[q].w += l4.w;
                            }
                        }
                        __syncwarp(__activemask());
                    }
                    {
                        float mx = 0.f;
#pragma unroll
                        for (int q = 0; q < 4; ++q) mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[q].x), fabsf(v[q].y))), fmaxf(fabsf(v[q].z), fabsf(v[q].w)));
                        int E = __float_as_int(mx) >> 23;                          // biased exponent of the largest numerator
                        E = E < 40 ? 40 : (E > 230 ? 230 : E);                     // (keeps 2^-e and den 2^-e finite; |v| < 2^-87 rounds to 0)
                        // c = 2^-e (1 - 2^-15), e = (E - 127) - 14: max|v| c < 32767.5, so the rounded value always fits an int16
                        const float scale = __int_as_float((268 - E) << 23) * 0.999969482421875f;
                        den *= scale;
                        M += (float)(E - 141) + 4.4028e-5f;                        // M - log2(c)
                        constexpr float kMagic = 12582912.f;                       // 1.5 * 2^23: the low mantissa bits of x + kMagic are rint(x)
                        uint32_t w[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 scp = make_float2(scale, scale), mgp = make_float2(kMagic, kMagic);
                            const float2 m01 = __ffma2_rn(make_float2(v[q].x, v[q].y), scp, mgp), m23 = __ffma2_rn(make_float2(v[q].z, v[q].w), scp, mgp);
                            w[2 * q] = __byte_perm(__float_as_uint(m01.x), __float_as_uint(m01.y), 0x5410);
                            w[2 * q + 1] = __byte_perm(__float_as_uint(m23.x), __float_as_uint(m23.y), 0x5410);
                        }
                        st_keep8u(reinterpret_cast<unsigned char *>(a.recV) + my_node * REC_BYTES + sl * 256 + head * 32, w);
                    }
                    DMS[head * 128 + tt] = make_float2(den, M);            // written out with the other heads' at the end of the star
                }
                KN_STAMP(12);                                          // partial
            }
        }
        KN_STAMP(12);                                                  // (last) partial
        if (live) {                                                    // this row's 8 x (denominator, max): one 64-byte chunk of the record
            const float2 d0 = DMS[tt], d1 = DMS[128 + tt], d2 = DMS[256 + tt], d3 = DMS[384 + tt];
            const float2 d4 = DMS[512 + tt], d5 = DMS[640 + tt], d6 = DMS[768 + tt], d7 = DMS[896 + tt];
            float *rd = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(a.recV) + my_node * REC_BYTES + 512 + (i < lt ? 0 : 64));
            st_keep8(rd, make_float4(d0.x, d0.y, d1.x, d1.y), make_float4(d2.x, d2.y, d3.x, d3.y));
            st_keep8(rd + 8, make_float4(d4.x, d4.y, d5.x, d5.y), make_float4(d6.x, d6.y, d7.x, d7.y));
        }
        if (tid == 0) {
            SI[0] = grabbed;
            SI[1] = mtake >= total ? -1 : mtake;                       // (no slices left)
        }
        __syncthreads();                                               // this star's records are all written; next slot / slice known
        if (havestar && tid == 0) {
            __threadfence();
            atomicAdd(a.flags + b, 1);                                 // one more star of instance b is out
            if (PAIR && NS * cur + 1 < total) atomicAdd(a.flags + (NS * cur + 1) / n, 1);   // ... and the second of the pair
        }
        cur = nxt;
        nxt = SI[0];
        cand = SI[1];
        // (no barrier here: SI[0] and SI[1] are only rewritten at the end of the next iteration, many barriers from now; the top of the
        //  iteration writes SI[2] alone)
        KN_STAMP(13);                                                  // end of star: barriers, publish
    }
#ifdef KN_STAMPS
    if (lane == 0 && blockIdx.x < 148 * 4)
        for (int k = 0; k < 16; ++k) g_kn_stamps[(blockIdx.x * 4 + warp) * 16 + k] = stamp_acc[k];
#endif
    // ---------------------------------------------------------------- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tslot), "n"(TMEM_COLS) : "memory");
}

}  // namespace

namespace gnngls {
size_t kn_tc_workspace_bytes(int B, int n) {
    const size_t M = (size_t)B * ((size_t)n * (n - 1) / 2);
    return M * (size_t)REC_BYTES + sizeof(int) * ((size_t)B + 4);   // one record (both stars) per node, one counter per instance, slot + slice counters
}

int launch_kn_tc(const KnArgs &args_in, int B, void *workspace, cudaStream_t st) {
    KnArgs args = args_in;
    const int n = args.n;
    const bool pair = n <= 64;                                         // two stars per CTA iteration
    const TLayout L(n, pair);
    GNNGLS_REQUIRE(n <= KPAD, GNNGLS_ERR_UNSUPPORTED, "the tcgen05 K_n kernel handles n <= %d", KPAD);
    const size_t M = (size_t)B * ((size_t)n * (n - 1) / 2);
    args.recV = static_cast<float *>(workspace);
    args.recDM = nullptr;
    args.flags = reinterpret_cast<int *>(static_cast<unsigned char *>(workspace) + M * REC_BYTES);
    GNNGLS_CUDA_OK(cudaMemsetAsync(args.flags, 0, sizeof(int) * ((size_t)B + 4), st));
    auto kernel = pair ? gat_kn_tc_kernel<true> : gat_kn_tc_kernel<false>;
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t stars = (int64_t)B * n;
    GNNGLS_REQUIRE(stars < ((int64_t)1 << 30), GNNGLS_ERR_UNSUPPORTED, "B*n too large for one launch");
    const int grid_max = gnngls::device_sm_count() * 4;
    const int64_t slots = pair ? (stars + 1) / 2 : stars;
    const unsigned grid = (unsigned)(slots < grid_max ? slots : grid_max);
    kernel<<<grid, T_THREADS, L.total, st>>>(args, B);
    GNNGLS_LAUNCH_OK("gat_kn_tc_kernel");
    return GNNGLS_OK;
}
}  // namespace gnngls

// debug: per-(CTA, warp) phase cycle totals of the last launch (error unless built with -DKN_STAMPS)
extern "C" int gnngls_debug_kn_stamps(unsigned long long *out, int count) {
#ifdef KN_STAMPS
    if (count > 148 * 4 * 4 * 16) count = 148 * 4 * 4 * 16;
    GNNGLS_CUDA_OK(cudaMemcpyFromSymbol(out, g_kn_stamps, sizeof(unsigned long long) * count));
    return GNNGLS_OK;
#else
    (void)out; (void)count;
    return GNNGLS_ERR_UNSUPPORTED;
#endif
}
