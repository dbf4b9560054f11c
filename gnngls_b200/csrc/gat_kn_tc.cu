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
// Persistent kernel, one CTA of 4 teams x 128 threads per SM; CTA c takes the stars (instance b, vertex i) number c, c + G,
// c + 2G, ... (G = grid size).  Destination {i,j} belongs to two stars; the one that is "behind" on the circle of vertices
// ((i - j) mod n < n/2) finalises it, the other one publishes its partial (numerators, denominator, reference max) to
// the record buffer -- so every star finalises at most n/2 rows and keeps them in a shared-memory stash.  Per star:
//   1. (thread = member k when building operands, = destination row j afterwards) each team walks 2 heads: operand rows
//      + indicator -> MMA -> accumulators; the MMA of the second head runs under the epilogue of the first.  Rows this
//      star does not finalise go straight to the record buffer; ONE flag per star is released at the end.
//   2. the rows a star finalises are finished TWO iterations later, in the shadow of that iteration's first MMA: its partner
//      stars (numbers within n of it, so at most one iteration younger as long as G >= n) published a whole iteration ago,
//      so their flags are up, and their records were fetched by cp.async at the top of the iteration (skip rows: L2
//      prefetch); one warp per row merges in fixed (lower vertex, higher vertex) order -- deterministic and
//      batching-invariant bitwise -- applies bias + skip + BN1 and writes h1 with full lines.
// Nothing waits for global memory between the top of an iteration and its end.  A partner that is late all the same is
// never waited for before this CTA's own star of the iteration is published: blocking waits (end of the iteration) then
// only ever depend on stars of strictly older iterations, so they cannot form a cycle.
#include "gat_kn.cuh"

// Phase timing (debug builds only: -DKN_STAMPS): thread 0 of every team accumulates clock64() differences per phase;
// read back with gnngls_debug_kn_stamps().
#ifdef KN_STAMPS
__device__ unsigned long long g_kn_stamps[148 * 4 * 16];
#define KN_STAMP(slot)                                              \
    do {                                                            \
        if (tt == 0) {                                              \
            const long long now__ = clock64();                      \
            stamp_acc[slot] += (unsigned long long)(now__ - stamp_last); \
            stamp_last = now__;                                     \
        }                                                           \
    } while (0)
#else
#define KN_STAMP(slot) do { } while (0)
#endif

namespace {

constexpr int P_THREADS = 512, P_WARPS = 16, TEAMS = 4, TEAM = 128, HPT = H_ / TEAMS, KPAD = 128;
constexpr int XN = 48;                                   // MMA N: 16 (A-branch) + 16 (B-branch) + 2 denominators, padded to 16s
constexpr int X_KB = (XN / 8) * 128;                     // bytes of one 8-member block of the B operand (6 core matrices)
constexpr int X_BYTES = (KPAD / 8) * X_KB;               // 12 KB per team
constexpr float kLeadGap = 6.f;                          // lead (log2 units) of the largest score above which its member is handled in fp32
constexpr int SROW = 148;                                // floats per stash row: 128 numerators + 8 x (denominator, max) + pad (148 % 32 = 20: conflict-free float4 rows)
constexpr int LREC = 144;                                // floats of a partner record: 128 numerators + 8 x (denominator, max)
constexpr int ESTR = 20;                                 // floats per member in the score buffer: el (8), er (8), pad (bank spread)
constexpr int TMEM_COLS = 512;                           // per team 128: 64 columns indicator (K=128 as fp16 pairs) + 48 accumulator

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void team_barrier(int team) { asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory"); }
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
// one row through the bulk-copy engine (global -> shared), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
__device__ __forceinline__ void wait_flag(const int *f) {
    unsigned long long t0 = 0;
    while (ld_acquire_gpu(f) == 0) {
        __nanosleep(64);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t0 == 0) t0 = t1;
        else if (t1 - t0 > 4000000000ull) asm volatile("trap;");       // fail loudly instead of hanging
    }
}

struct PLayout {
    int R;                                                   // most rows one star finalises
    int lrow;                                                // floats per landing row: record, + the skip row when shared memory allows
    unsigned eler_off, x_off, elh_off, hd_off, tot_off, stash_off, land_off, rows_off, bar_off, total;
    __host__ __device__ explicit PLayout(int n) {
        R = n / 2;
        eler_off = 0;                                        // [2][n][ESTR] fp32 scores of the current / next star
        x_off = (2u * (unsigned)n * ESTR * 4u + 127u) & ~127u;   // [4 teams][X_BYTES] B operand
        elh_off = x_off + TEAMS * X_BYTES;                   // [8 heads][128] fp16 centred scores el - ref
        hd_off = elh_off + 8u * 256u;                        // [8][4] words: ref, m1, arg-max member, "arg-max handled in fp32"
        tot_off = hd_off + 8u * 16u;                         // [4 teams][2][36] floats: TotB (16), total dB, pad, fp32 features of the leading member (16)
        stash_off = tot_off + TEAMS * 2u * 36u * 4u;         // [3][R][SROW] floats: partials of the rows this CTA's last three stars finalise
        land_off = stash_off + 3u * (unsigned)R * SROW * 4u; // [R][lrow] floats: partner records (+ skip rows) of the star being finalised
        lrow = LREC + 128;
        if (land_off + (unsigned)R * (lrow * 4u + 8u) + 64u > 227u * 1024u) lrow = LREC;   // n > 104: skip rows come through L2 instead
        rows_off = land_off + (unsigned)R * lrow * 4u;       // [R] (node within the instance, partner vertex or ~vertex if its record is late)
        bar_off = rows_off + (unsigned)R * 8u;               // 4 MMA mbarriers + landing mbarrier + tmem slot
        total = bar_off + (TEAMS + 1) * 8u + 16u;
    }
};

// first node of vertex i's run in the sorted-tuple order: node {i,k}, i < k, is tri(i) + k - i - 1
__device__ __forceinline__ int tri(int i, int n) { return (i * (2 * n - i - 1)) >> 1; }
__device__ __forceinline__ int kn_local(int i, int k, int n) { return i < k ? tri(i, n) + k - i - 1 : tri(k, n) + i - k - 1; }

// the rows of a star whose partner star had not published when the iteration started (rare): wait, fetch, finish.
// Only called after this CTA's own star of the iteration is out.
__device__ __noinline__ void finalize_late(const KnArgs &a, int n, int pb, int pi, unsigned late, const float *stash) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t pnode0 = (size_t)pb * ((size_t)n * (n - 1) / 2);
    const float4 sc4 = __ldg(reinterpret_cast<const float4 *>(a.bn_scale) + lane), sh4 = __ldg(reinterpret_cast<const float4 *>(a.bn_shift) + lane);
    const float4 bb4 = a.bias ? __ldg(reinterpret_cast<const float4 *>(a.bias) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < 4; ++q) {
        if (!((late >> q) & 1u)) continue;
        const int r = warp + P_WARPS * q;
        int j = pi - 1 - r;
        if (j < 0) j += n;
        if (r >= (n - 1) / 2) j = pi + n / 2;
        const size_t node = pnode0 + kn_local(pi, j, n);
        wait_flag(a.flags + (size_t)pb * n + j);
        const float4 pv = __ldcg(reinterpret_cast<const float4 *>(a.recV + node * D_) + lane);
        const float2 pdm = __ldcg(reinterpret_cast<const float2 *>(a.recDM + node * 2 * H_) + (lane >> 2));
        const float4 hv = __ldg(reinterpret_cast<const float4 *>(a.h + node * D_) + lane);
        const float *Sr = stash + r * SROW;
        const float4 v = *reinterpret_cast<const float4 *>(Sr + 4 * lane);
        const float2 dm = *reinterpret_cast<const float2 *>(Sr + D_ + 2 * (lane >> 2));
        const float mx = fmaxf(dm.y, pdm.y);
        const float s1 = ex2(dm.y - mx), s2 = ex2(pdm.y - mx);
        const float inv = 1.f / fmaf(dm.x, s1, pdm.x * s2);
        const float a1 = s1 * inv, a2 = s2 * inv;
        float4 o;
        o.x = (hv.x + (fmaf(v.x, a1, pv.x * a2) + bb4.x)) * sc4.x + sh4.x;
        o.y = (hv.y + (fmaf(v.y, a1, pv.y * a2) + bb4.y)) * sc4.y + sh4.y;
        o.z = (hv.z + (fmaf(v.z, a1, pv.z * a2) + bb4.z)) * sc4.z + sh4.z;
        o.w = (hv.w + (fmaf(v.w, a1, pv.w * a2) + bb4.w)) * sc4.w + sh4.w;
        reinterpret_cast<float4 *>(a.h1 + node * D_)[lane] = o;
        if (a.h1_tf32) reinterpret_cast<float4 *>(a.h1_tf32 + node * D_)[lane] = tf32_round4(o);
        if ((lane & 7) == 0) discard_l2_128(a.recV + node * D_ + 4 * lane);
    }
}

__global__ void __launch_bounds__(P_THREADS, 1) gat_kn_tc_kernel(const KnArgs a, const int total_stars) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = a.n;
    const PLayout L(n);
    const int R = L.R;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int team = warp >> 2, wq = warp & 3, tt = tid & (TEAM - 1);
    const size_t N = (size_t)n * (n - 1) / 2;

    float *ELER = reinterpret_cast<float *>(smem + L.eler_off);
    unsigned char *Xs = smem + L.x_off + team * X_BYTES;
    __half *ELHall = reinterpret_cast<__half *>(smem + L.elh_off);
    float *HD = reinterpret_cast<float *>(smem + L.hd_off);
    float *TOT = reinterpret_cast<float *>(smem + L.tot_off) + team * 72;   // [0,17): column totals; [20,36), [36,52): leading member's features (head parity)
    float *STASH = reinterpret_cast<float *>(smem + L.stash_off);
    float *LAND = reinterpret_cast<float *>(smem + L.land_off);
    int2 *ROWS = reinterpret_cast<int2 *>(smem + L.rows_off);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bar_off);
    uint64_t *land_bar = bars + TEAMS;
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + TEAMS + 1);
    const int LROW = L.lrow;
    const bool land_h = LROW > LREC;

    // ---------------------------------------------------------------- setup
    if (tid == 0) {
#pragma unroll
        for (int t = 0; t < TEAMS; ++t) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[t])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(land_bar)), "n"(P_THREADS / 2));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tslot + team * 128;                        // this team's columns
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;               // this warp's lane quarter
    const int nk = (n + 15) >> 4, nch = (n + 31) >> 5;                 // MMAs (16 members each) / indicator chunks (32 members each)
    unsigned char *xrow = Xs + (tt >> 3) * X_KB + (tt & 7) * 16;       // this member's row of the B operand (6 pieces, 128 B apart)
    const int G = gridDim.x, step_b = G / n, step_i = G - step_b * n;  // the next star of this CTA is G stars further
    const int half_rows = (n - 1) >> 1;

    // scores of star (sb, si) -> buffer `buf` (cp.async; the vertex's own slot is zero-filled)
    auto issue_scores = [&](int sb, int si, int buf, int t0, int nthreads) {
        for (int idx = t0; idx < 4 * n; idx += nthreads) {
            const int k = idx >> 2, p = idx & 3;
            float *dst = ELER + (buf * n + k) * ESTR + p * 4;
            if (k != si) cp_async16(dst, (p < 2 ? a.el : a.er) + ((size_t)sb * N + kn_local(si, k, n)) * H_ + (p & 1) * 4);
            else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    // (b, i) of this CTA's star of the iteration, of the next one, and of the last two (b < 0: none)
    int cur = blockIdx.x, cb = cur / n, ci = cur - cb * n;
    int b1 = -1, i1 = 0, b2 = -1, i2 = 0;
    if (cur < total_stars) issue_scores(cb, ci, 0, tid, P_THREADS);
    cp_async_commit();
#ifdef KN_STAMPS
    unsigned long long stamp_acc[16] = {};
    long long stamp_last = clock64();
#endif
    uint32_t parity = 0, land_parity = 0;
    unsigned late = 0u;                                                // rows of the previous iteration's prev still to finish (this warp)
    int late_b = 0, late_i = 0, late_buf = 0;

    for (int it = 0;; ++it) {
        const bool havecur = cur < total_stars, haveprev = b2 >= 0;    // prev = (b2, i2): the star finalised in this iteration
        if (!havecur && b1 < 0 && b2 < 0) break;
        const int cbuf = it & 1, sbuf = it % 3, pbuf = (it + 1) % 3;  // score buffer / stash of the current star / stash of prev
        const int b = cb, i = ci;
        int nb = cb + step_b, ni = ci + step_i;                        // next star
        if (ni >= n) { ni -= n; ++nb; }
        // ---- top of the iteration, part A: requests whose answers are needed later on
        // this thread as member / destination row tt of the current star; its features for the team's first head
        const bool live = havecur && tt < n && tt != i;
        const int my_local = live ? kn_local(i, tt, n) : 0;
        const uint4 *ftrow = reinterpret_cast<const uint4 *>(static_cast<const unsigned char *>(a.ft) + ((size_t)b * N + my_local) * 256);
        uint4 f0 = make_uint4(0u, 0u, 0u, 0u), f1 = f0;
        if (live) { f0 = __ldg(ftrow + 2 * team * HPT); f1 = __ldg(ftrow + 2 * team * HPT + 1); }
        const int Rp = haveprev ? half_rows + (((n & 1) == 0 && i2 < (n >> 1)) ? 1 : 0) : 0;
        KN_STAMP(0);                                                   // top A
        cp_async_wait_group<0>();                                      // the current star's scores (requested an iteration ago)
        __syncthreads();                                               // ... and every warp has finished the previous iteration
        KN_STAMP(1);                                                   // wait scores + barrier 1
#ifdef KN_STAMPS
        if (tt == 0) stamp_acc[12] += __popc(late);
#endif
        if (late) finalize_late(a, n, late_b, late_i, late, STASH + late_buf * R * SROW);   // (per warp; rare)
        KN_STAMP(2);                                                   // late rows
        const float *E = ELER + cbuf * n * ESTR;
        if (warp < H_) {
            // ---- warps 0..7: top-2 of el per head (warp w = head w); the fp16 copy of the CENTRED scores that the branch decision uses
            if (havecur) {
                float ev[KPAD / 32];
                Top2 t2{-INFINITY, -INFINITY, 0};
#pragma unroll
                for (int q = 0; q < KPAD / 32; ++q) {
                    const int k = lane + 32 * q;
                    const bool ok = k < n && k != i;
                    ev[q] = ok ? E[k * ESTR + warp] : -INFINITY;
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
                const bool fix = t2.m1 - t2.m2 > kLeadGap;             // (n >= 3: the runner-up exists)
                const float ref = fix ? t2.m2 : t2.m1;
#pragma unroll
                for (int q = 0; q < KPAD / 32; ++q) ELHall[warp * KPAD + lane + 32 * q] = __float2half_rn(ev[q] - ref);
                if (lane == 0) *reinterpret_cast<float4 *>(HD + warp * 4) = make_float4(ref, t2.m1, __int_as_float(t2.a1), __int_as_float(fix ? 1 : 0));
            }
        } else {
            // ---- warps 8..15: everything this and the next iteration will read from global memory is requested here
            const int lw = warp - H_;                                  // rows lw + 8 q, q = 0..7
            if (lw == 0 && lane == 0 && b1 >= 0) {                     // the previous star's records are all written (barrier above): publish
                __threadfence();
                st_release_gpu(a.flags + b1 * n + i1, 1);
            }
            bool arrived = false;
            if (haveprev) {
                const size_t pnode0 = (size_t)b2 * N;
                const int r = lw + 8 * lane;                           // lanes 0..7: one row each, fetched by the bulk-copy engine
                if (lane < 8 && r < Rp) {
                    int j = i2 - 1 - r;
                    if (j < 0) j += n;
                    if (r >= half_rows) j = i2 + (n >> 1);
                    const int nl = kn_local(i2, j, n);
                    const uint32_t dst = smem_u32(LAND + r * LROW), bar = smem_u32(land_bar);
                    const float *hrow = a.h + (pnode0 + nl) * D_;
                    if (land_h) bulk_g2s(dst + LREC * 4, hrow, 512, bar);  // skip row: no dependency
                    else asm volatile("cp.async.bulk.prefetch.L2.global [%0], 512;" ::"l"(hrow) : "memory");
                    // Partner stars published a whole iteration ago, normally.  This is a look, not a wait: this CTA's current star is
                    // not out yet, and a partner's CTA may in turn need it.  Late rows are finished after the next iteration's barrier.
                    const int fl = ld_acquire_gpu(a.flags + (size_t)b2 * n + j);
                    ROWS[r] = make_int2(nl, fl != 0 ? j : ~j);
                    mbar_arrive_expect_tx(bar, (land_h ? 512u : 0u) + (fl != 0 ? (uint32_t)LREC * 4u : 0u));
                    if (fl != 0) {
                        asm volatile("fence.proxy.async.global;" ::: "memory");    // the copies below read what the flag guards
                        bulk_g2s(dst, a.recV + (pnode0 + nl) * D_, 512, bar);
                        bulk_g2s(dst + 512, a.recDM + (pnode0 + nl) * 2 * H_, 64, bar);
                    }
                    arrived = true;
                }
            }
            // the landing mbarrier completes its phase when all 256 loader threads have arrived and every byte announced has landed
            if (!arrived) mbar_arrive(smem_u32(land_bar));
            if (cur + G < total_stars) issue_scores(nb, ni, cbuf ^ 1, tid - P_THREADS / 2, P_THREADS / 2);
            cp_async_commit();
        }
        KN_STAMP(3);                                                   // top-2 | requests
        __syncthreads();                                               // heads' references and centred scores, row table visible
        KN_STAMP(4);                                                   // barrier 2

        // two rows of prev (q0, q0 + 1 of this warp's four): merge the landed partner record with the stashed partial and finish.
        // This star's partial first, the partner's second -- who finalises is a function of (i, j, n) only, so the result is
        // deterministic.  Runs in the shadow of an MMA.
        auto finalize_prev = [&](int q0) {
            if (warp + P_WARPS * q0 >= Rp) return;
            const size_t pnode0 = (size_t)b2 * N;
            const float *hb = a.h + pnode0 * D_;
            float *h1b = a.h1 + pnode0 * D_;
            const float *stash = STASH + pbuf * R * SROW;
            const float4 sc4 = __ldg(reinterpret_cast<const float4 *>(a.bn_scale) + lane), sh4 = __ldg(reinterpret_cast<const float4 *>(a.bn_shift) + lane);
            const float4 bb4 = a.bias ? __ldg(reinterpret_cast<const float4 *>(a.bias) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            int2 row[2];
            float4 hv[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = warp + P_WARPS * (q0 + q);
                row[q] = (r < Rp) ? ROWS[r] : make_int2(0, -1);
                if (!land_h) hv[q] = __ldg(reinterpret_cast<const float4 *>(hb + row[q].x * D_) + lane);   // (prefetched into L2)
            }
            mbar_wait(land_bar, land_parity);                          // the loader warps' copies have landed
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = warp + P_WARPS * (q0 + q);
                if (r >= Rp || row[q].y < 0) continue;                 // (late partner: finished after the next iteration's barrier)
                const float *Lr = LAND + r * LROW, *Sr = stash + r * SROW;
                if (land_h) hv[q] = *reinterpret_cast<const float4 *>(Lr + LREC + 4 * lane);
                const float4 pv = *reinterpret_cast<const float4 *>(Lr + 4 * lane), v = *reinterpret_cast<const float4 *>(Sr + 4 * lane);
                const float2 pdm = *reinterpret_cast<const float2 *>(Lr + D_ + 2 * (lane >> 2)), dm = *reinterpret_cast<const float2 *>(Sr + D_ + 2 * (lane >> 2));
                const float mx = fmaxf(dm.y, pdm.y);
                const float s1 = ex2(dm.y - mx), s2 = ex2(pdm.y - mx);
                const float inv = rcp_approx(fmaf(dm.x, s1, pdm.x * s2));
                const float a1 = s1 * inv, a2 = s2 * inv;
                float4 o;
                o.x = (hv[q].x + (fmaf(v.x, a1, pv.x * a2) + bb4.x)) * sc4.x + sh4.x;
                o.y = (hv[q].y + (fmaf(v.y, a1, pv.y * a2) + bb4.y)) * sc4.y + sh4.y;
                o.z = (hv[q].z + (fmaf(v.z, a1, pv.z * a2) + bb4.z)) * sc4.z + sh4.z;
                o.w = (hv[q].w + (fmaf(v.w, a1, pv.w * a2) + bb4.w)) * sc4.w + sh4.w;
                reinterpret_cast<float4 *>(h1b + row[q].x * D_)[lane] = o;
                if (a.h1_tf32) reinterpret_cast<float4 *>(a.h1_tf32 + (pnode0 + row[q].x) * D_)[lane] = tf32_round4(o);
                // the consumed numerator record is dead (read exactly once): drop its dirty L2 lines instead of writing them back
                if ((lane & 7) == 0) discard_l2_128(a.recV + (pnode0 + row[q].x) * D_ + 4 * lane);
            }
        };

        if (havecur) {
            // which of the two stars of destination {i, tt} finalises it
            int dist = i - tt;
            if (dist < 0) dist += n;
            const bool fin = live && (2 * dist < n || (2 * dist == n && i < tt));
            float *srow = STASH + (sbuf * R + ((2 * dist == n) ? half_rows : dist - 1)) * SROW;
#pragma unroll
            for (int hh = 0; hh < HPT; ++hh) {
                const int head = team * HPT + hh;
                const float4 hd = *reinterpret_cast<const float4 *>(HD + head * 4);   // ref, m1, leading member, "leading member in fp32"
                const float ref = hd.x;
                const bool fix = __float_as_int(hd.w) != 0;            // team-uniform
                const bool lead = fix && tt == __float_as_int(hd.z);
                const float erh = live ? E[tt * ESTR + 8 + head] : 0.f;
                const __half th16 = live ? __float2half_rn(-erh - ref) : __float2half_rn(tt == i ? -INFINITY : INFINITY);
                float *TOTl = TOT + 20 + (hh & 1) * 16;                // fp32 features of the leading member (parity buffer)
                // ---- B operand row of member tt: [A ft | A' ft | A A' 0...]
                if (tt < 16 * nk) {
                    uint4 xa0 = make_uint4(0u, 0u, 0u, 0u), xa1 = xa0, xb0 = xa0, xb1 = xa0, xd = xa0;
                    if (live && !lead) {
                        const float d = E[tt * ESTR + head] - ref;
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
                    } else if (live) {                                 // the leading member of this head: fp32 features for the epilogue
                        float ff[16];
                        unpack8(f0, reinterpret_cast<float(&)[8]>(ff[0]));
                        unpack8(f1, reinterpret_cast<float(&)[8]>(ff[8]));
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<float4 *>(TOTl + 4 * q) = make_float4(ff[4 * q], ff[4 * q + 1], ff[4 * q + 2], ff[4 * q + 3]);
                    }
                    *reinterpret_cast<uint4 *>(xrow) = xa0;
                    *reinterpret_cast<uint4 *>(xrow + 128) = xa1;
                    *reinterpret_cast<uint4 *>(xrow + 256) = xb0;
                    *reinterpret_cast<uint4 *>(xrow + 384) = xb1;
                    *reinterpret_cast<uint4 *>(xrow + 512) = xd;
                }
                if (live && hh + 1 < HPT) { f0 = __ldg(ftrow + 2 * (head + 1)); f1 = __ldg(ftrow + 2 * (head + 1) + 1); }   // next head's features
                // ---- indicator row of destination tt: I[tt][k] = [el_k - ref >= -er_tt - ref] as fp16 1.0 / 0.0, straight into tensor
                // memory.  Row i (the vertex itself, not a destination) takes threshold -inf: its accumulator row is the column totals.
                {
                    const __half2 th2 = __half2half2(th16);
                    const __half *ELH = ELHall + head * KPAD;
                    for (int c = 0; c < nch; ++c) {
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
                }
                KN_STAMP(5);                                           // operand row + indicator
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // B operand written through the generic proxy
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                team_barrier(team);                                    // operands complete
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tt == 0) {
                    for (int ks = 0; ks < nk; ++ks)
                        umma_f16_ts(tbase + 64, tbase + ks * 8, make_mn_desc(smem_u32(Xs) + ks * 2 * X_KB, X_KB, 128), kIdesc, ks != 0);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[team])) : "memory");
                }
                __syncwarp();
                KN_STAMP(6);                                           // barrier A + MMA issue
                if (haveprev) finalize_prev(2 * hh);                   // in the shadow of the MMA
                KN_STAMP(7);                                           // two rows of prev
                // ---- accumulators
                uint32_t SA[16], SB[16], SD0, SD1;
                mbar_wait(&bars[team], parity);
                parity ^= 1;
                KN_STAMP(8);                                           // rest of the MMA
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_ld16(tbase + lane_sel + 64, SA);
                tmem_ld16(tbase + lane_sel + 80, SB);
                tmem_ld2(tbase + lane_sel + 96, SD0, SD1);             // sum A, sum A' over the A-branch members
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (tt == i) {                                         // the all-ones row: column totals
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4 *>(TOT + 4 * q) = make_uint4(SB[4 * q], SB[4 * q + 1], SB[4 * q + 2], SB[4 * q + 3]);
                    TOT[16] = __uint_as_float(SD1);
                }
                KN_STAMP(9);                                           // accumulators -> registers
                team_barrier(team);                                    // totals visible; every row has its accumulators
                KN_STAMP(10);                                          // barrier B
                if (live) {
                    // ---- this star's partial for destination tt (fp32): v = C1 SA + C2 (Tot - SB) - self (+ leading member)
                    const float s = ref + erh;
                    const float c = ex2(-0.8f * fabsf(s));
                    const float C1 = s >= 0.f ? 1.f : c, C2 = s >= 0.f ? c : 1.f;
                    float M = lrelu(s);
                    const bool self_a = __hge(ELHall[head * KPAD + tt], th16);     // the branch the MMA put this row's own member in
                    const uint32_t xdw = *reinterpret_cast<const uint32_t *>(xrow + 512);
                    const float2 aa = __half22float2(*reinterpret_cast<const __half2 *>(&xdw));   // (A, A') as the MMA saw them
                    float den = fmaf(C1, __uint_as_float(SD0), C2 * (TOT[16] - __uint_as_float(SD1))) - (self_a ? C1 * aa.x : C2 * aa.y);
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
                    float *rv = a.recV + ((size_t)b * N + my_local) * D_ + head * F_;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {                      // 4 features at a time
                        const float4 t4 = *reinterpret_cast<const float4 *>(TOT + 4 * q);
                        const uint2 xs = *reinterpret_cast<const uint2 *>(xself + (q >> 1) * 128 + (q & 1) * 8);
                        const float2 x01 = __half22float2(*reinterpret_cast<const __half2 *>(&xs.x)), x23 = __half22float2(*reinterpret_cast<const __half2 *>(&xs.y));
                        float4 v;
                        v.x = fmaf(k1, __uint_as_float(SA[4 * q]), fmaf(-k2, __uint_as_float(SB[4 * q]), fmaf(-ks, x01.x, k2 * t4.x)));
                        v.y = fmaf(k1, __uint_as_float(SA[4 * q + 1]), fmaf(-k2, __uint_as_float(SB[4 * q + 1]), fmaf(-ks, x01.y, k2 * t4.y)));
                        v.z = fmaf(k1, __uint_as_float(SA[4 * q + 2]), fmaf(-k2, __uint_as_float(SB[4 * q + 2]), fmaf(-ks, x23.x, k2 * t4.z)));
                        v.w = fmaf(k1, __uint_as_float(SA[4 * q + 3]), fmaf(-k2, __uint_as_float(SB[4 * q + 3]), fmaf(-ks, x23.y, k2 * t4.w)));
                        if (addlead) {
                            const float4 l4 = *reinterpret_cast<const float4 *>(TOTl + 4 * q);
                            v.x += l4.x; v.y += l4.y; v.z += l4.z; v.w += l4.w;
                        }
                        if (fin) *reinterpret_cast<float4 *>(srow + head * F_ + 4 * q) = v;
                        else __stcg(reinterpret_cast<float4 *>(rv) + q, v);
                    }
                    if (fin) *reinterpret_cast<float2 *>(srow + D_ + 2 * head) = make_float2(den, M);
                    else __stcg(reinterpret_cast<float2 *>(a.recDM + ((size_t)b * N + my_local) * 2 * H_ + 2 * head), make_float2(den, M));
                }
                KN_STAMP(11);                                          // partial
            }
        } else if (haveprev) {
            finalize_prev(0);
            finalize_prev(2);
        }
        // rows whose partner was late wait until this CTA's own star of the iteration is out (after the next iteration's first barrier)
        late = 0u;
        if (haveprev) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (warp + P_WARPS * q < Rp && ROWS[warp + P_WARPS * q].y < 0) late |= 1u << q;
            late_b = b2; late_i = i2; late_buf = pbuf;
        }
        land_parity ^= 1;                                              // (the loader warps arrive once per iteration)
        b2 = b1; i2 = i1;
        b1 = havecur ? b : -1; i1 = i;
        cur += G; cb = nb; ci = ni;
    }
#ifdef KN_STAMPS
    if (tt == 0 && blockIdx.x < 148)
        for (int k = 0; k < 16; ++k) g_kn_stamps[(blockIdx.x * 4 + team) * 16 + k] = stamp_acc[k];
#endif
    // ---------------------------------------------------------------- teardown
    if (late) finalize_late(a, n, late_b, late_i, late, STASH + late_buf * R * SROW);       // (all stars of this CTA are out)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tslot), "n"(TMEM_COLS) : "memory");
}

}  // namespace

namespace gnngls {
int launch_kn_tc(const KnArgs &args, int B, cudaStream_t st) {
    const PLayout L(args.n);
    GNNGLS_REQUIRE(args.n <= KPAD, GNNGLS_ERR_UNSUPPORTED, "the tcgen05 K_n kernel handles n <= %d", KPAD);
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(gat_kn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    const int64_t stars = (int64_t)B * args.n;
    GNNGLS_REQUIRE(stars < ((int64_t)1 << 30), GNNGLS_ERR_UNSUPPORTED, "B*n too large for one launch");
    const int sms = gnngls::device_sm_count();
    const unsigned grid = (unsigned)(stars < sms ? stars : sms);
    gat_kn_tc_kernel<<<grid, P_THREADS, L.total, st>>>(args, (int)stars);
    GNNGLS_LAUNCH_OK("gat_kn_tc_kernel");
    return GNNGLS_OK;
}
}  // namespace gnngls

// debug: per-(CTA, team) phase cycle totals of the last launch (zeros unless built with -DKN_STAMPS)
extern "C" int gnngls_debug_kn_stamps(unsigned long long *out, int count) {
#ifdef KN_STAMPS
    if (count > 148 * 4 * 16) count = 148 * 4 * 16;
    GNNGLS_CUDA_OK(cudaMemcpyFromSymbol(out, g_kn_stamps, sizeof(unsigned long long) * count));
    return GNNGLS_OK;
#else
    (void)out; (void)count;
    return GNNGLS_ERR_UNSUPPORTED;
#endif
}
