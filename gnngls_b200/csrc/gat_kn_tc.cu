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
// One CTA per (instance, vertex i), two phases:
//   1. two teams of 128 threads (thread = member k when building operands, = destination row j afterwards) walk 4
//      heads each; every row's partial (16 numerators, denominator, reference max per head) lands in a shared-memory
//      stash.  No global-memory wait anywhere in this phase.
//   2. destination {i,j} belongs to two stars.  Rows j > i (this star is the lower one) are copied from the stash to the
//      record buffer with coalesced stores, then ONE flag per star is released.  Rows j < i: wait for the flags of the
//      lower stars (dispatched earlier: blockIdx order; they never wait before publishing), then one warp per row
//      merges the lower star's record with the stashed partial in fixed (lower, higher) order -- deterministic and
//      batching-invariant bitwise -- applies bias + skip + BN1 and writes h1 with full-line accesses.
#include "gat_kn.cuh"

namespace {

constexpr int TC_THREADS = 256, TEAM = 128, KPAD = 128;
constexpr int XN = 48;                                   // MMA N: 16 (A-branch) + 16 (B-branch) + 2 denominators, padded to 16s
constexpr int X_KB = (XN / 8) * 128;                     // bytes of one 8-member block of the B operand (6 core matrices)
constexpr int X_BYTES = (KPAD / 8) * X_KB;               // 12 KB per team
constexpr float kLeadGap = 6.f;                          // lead (log2 units) of the largest score above which its member is handled in fp32
constexpr int SROW = 148;                                // floats per stash row: 128 numerators + 8 x (denominator, max) + pad (148 % 32 = 20: conflict-free float4 rows)
constexpr int TMEM_COLS = 256;                           // per team: 64 columns indicator (K=128 as fp16 pairs) + 48 accumulator

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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = __uint_as_float(r[u]);
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

struct TcLayout {
    unsigned el_off, er_off, x_off, elh_off, hd_off, tot_off, stash_off, node_off, bar_off, total;
    __host__ __device__ explicit TcLayout(int n) {
        el_off = 0;                                          // [n][8] fp32
        er_off = el_off + (unsigned)n * 32u;                 // [n][8] fp32
        x_off = (er_off + (unsigned)n * 32u + 127u) & ~127u; // [2 teams][X_BYTES] B operand
        elh_off = x_off + 2u * X_BYTES;                      // [8 heads][128] fp16 centred scores el - ref
        hd_off = elh_off + 8u * 256u;                        // [8][4] words: ref, m1, arg-max member, "arg-max handled in fp32"
        tot_off = hd_off + 8u * 16u;                         // [2][36] floats: TotB (16), total dB, pad, fp32 features of the arg-max member (16)
        stash_off = tot_off + 2u * 36u * 4u;                 // [n][SROW] floats
        node_off = stash_off + (unsigned)n * SROW * 4u;      // [n] ints
        bar_off = (node_off + (unsigned)n * 4u + 15u) & ~15u;   // 2 mbarriers + tmem slot
        total = bar_off + 32u;
    }
};

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

__global__ void __launch_bounds__(TC_THREADS, 2) gat_kn_tc_kernel(const KnArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = a.n;
    const TcLayout L(n);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int team = warp >> 2, wq = warp & 3, tt = tid & (TEAM - 1);
    const int b = blockIdx.x / n, i = blockIdx.x - b * n;
    const int64_t N = (int64_t)n * (n - 1) / 2, node0 = (int64_t)b * N;

    const float *ELs = reinterpret_cast<const float *>(smem + L.el_off);
    const float *ERs = reinterpret_cast<const float *>(smem + L.er_off);
    unsigned char *Xs = smem + L.x_off + team * X_BYTES;
    __half *ELHall = reinterpret_cast<__half *>(smem + L.elh_off);
    float *HD = reinterpret_cast<float *>(smem + L.hd_off);
    float *TOT = reinterpret_cast<float *>(smem + L.tot_off) + team * 36;
    float *STASH = reinterpret_cast<float *>(smem + L.stash_off);
    int *NODE = reinterpret_cast<int *>(smem + L.node_off);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bar_off);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 2);

    // ---------------------------------------------------------------- setup + staging of the scores
    for (int k = tid; k < n; k += TC_THREADS) NODE[k] = (k != i) ? kn_node(i, k, n) : -1;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    for (int idx = tid; idx < n * 4; idx += TC_THREADS) {
        const int k = idx >> 2, p = idx & 3;
        const int node = NODE[k];
        unsigned char *dst = smem + (p < 2 ? L.el_off : L.er_off) + (size_t)k * 32 + (p & 1) * 16;
        if (node >= 0) cp_async16(dst, reinterpret_cast<const unsigned char *>(p < 2 ? a.el : a.er) + (size_t)(node0 + node) * 32 + (p & 1) * 16);
        else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    const bool live = tt < n && tt != i;                               // thread = member tt = destination row tt
    const size_t my_node = live ? (size_t)(node0 + NODE[tt]) : 0;
    const uint4 *ftrow = reinterpret_cast<const uint4 *>(static_cast<const unsigned char *>(a.ft) + my_node * 256);
    uint4 f0 = make_uint4(0u, 0u, 0u, 0u), f1 = f0;                    // this member's 16 features of the team's first head
    if (live) { f0 = __ldg(ftrow + 2 * team); f1 = __ldg(ftrow + 2 * team + 1); }
    cp_async_wait_all();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // top-2 of el per head (warp w = head w); the fp16 copy of the CENTRED scores that the branch decision uses
    {
        Top2 t2{-INFINITY, -INFINITY, 0};
        for (int k = lane; k < n; k += 32)
            if (k != i) t2 = top2_merge(t2, Top2{ELs[k * 8 + warp], -INFINITY, k});
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Top2 o;
            o.m1 = __shfl_xor_sync(0xffffffffu, t2.m1, off);
            o.m2 = __shfl_xor_sync(0xffffffffu, t2.m2, off);
            o.a1 = __shfl_xor_sync(0xffffffffu, t2.a1, off);
            t2 = top2_merge(t2, o);
        }
        const bool fix = t2.m1 - t2.m2 > kLeadGap;                     // (n >= 3: the runner-up exists)
        const float ref = fix ? t2.m2 : t2.m1;
        for (int k = lane; k < KPAD; k += 32)
            ELHall[warp * KPAD + k] = __float2half_rn((k < n && k != i) ? ELs[k * 8 + warp] - ref : -INFINITY);
        if (lane == 0) {
            HD[warp * 4] = ref; HD[warp * 4 + 1] = t2.m1;
            HD[warp * 4 + 2] = __int_as_float(t2.a1); HD[warp * 4 + 3] = __int_as_float(fix ? 1 : 0);
        }
    }
    __syncthreads();
    const uint32_t tbase = *tslot + team * 128;                        // this team's columns
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;               // this warp's lane quarter
    const int nk = (n + 15) >> 4, nch = (n + 31) >> 5;                 // MMAs (16 members each) / indicator chunks (32 members each)
    unsigned char *xrow = Xs + (tt >> 3) * X_KB + (tt & 7) * 16;       // this member's row of the B operand (6 pieces, 128 B apart)
    float *srow = STASH + tt * SROW;
    uint32_t parity = 0;

    // ---------------------------------------------------------------- phase 1: partial sums of this star, head by head
    for (int hh = 0; hh < H_ / 2; ++hh) {
        const int head = team + 2 * hh;
        const float el = live ? ELs[tt * 8 + head] : -INFINITY, er = live ? ERs[tt * 8 + head] : 0.f;
        const __half *ELH = ELHall + head * KPAD;
        const float ref = HD[head * 4], m1 = HD[head * 4 + 1];
        const int a1 = __float_as_int(HD[head * 4 + 2]);
        const bool fix = __float_as_int(HD[head * 4 + 3]) != 0;        // team-uniform
        const bool in_mma = live && !(fix && tt == a1);
        // ---- B operand row of member tt: [A ft | A' ft | A A' 0...]
        __half A16 = __float2half_rn(0.f), A516 = A16;
        if (tt < 16 * nk) {
            uint4 xa0 = make_uint4(0u, 0u, 0u, 0u), xa1 = xa0, xb0 = xa0, xb1 = xa0, xd = xa0;
            if (in_mma) {
                const float d = el - ref;
                A16 = __float2half_rn(ex2(d));
                A516 = __float2half_rn(ex2(kSlope * d));
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
            } else if (live) {                                         // the leading member of this head: fp32 features for the epilogue
                float ff[16];
                unpack8(f0, reinterpret_cast<float(&)[8]>(ff[0]));
                unpack8(f1, reinterpret_cast<float(&)[8]>(ff[8]));
#pragma unroll
                for (int f = 0; f < 16; ++f) TOT[20 + f] = ff[f];
            }
            *reinterpret_cast<uint4 *>(xrow) = xa0;
            *reinterpret_cast<uint4 *>(xrow + 128) = xa1;
            *reinterpret_cast<uint4 *>(xrow + 256) = xb0;
            *reinterpret_cast<uint4 *>(xrow + 384) = xb1;
            *reinterpret_cast<uint4 *>(xrow + 512) = xd;
            *reinterpret_cast<uint4 *>(xrow + 640) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (live && hh + 1 < H_ / 2) { f0 = __ldg(ftrow + 2 * (head + 2)); f1 = __ldg(ftrow + 2 * (head + 2) + 1); }   // next head's features
        // ---- indicator row of destination tt: I[tt][k] = [el_k - ref >= -er_tt - ref] as fp16 1.0 / 0.0, straight into tensor
        // memory.  Row i (the vertex itself, not a destination) takes threshold -inf: its accumulator row is the column totals.
        const __half th16 = live ? __float2half_rn(-er - ref) : __float2half_rn(tt == i ? -INFINITY : INFINITY);
        const __half2 th2 = __half2half2(th16);
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
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // B operand written through the generic proxy
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        team_barrier(team);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tt == 0) {
            for (int ks = 0; ks < nk; ++ks)
                umma_f16_ts(tbase + 64, tbase + ks * 8, make_mn_desc(smem_u32(Xs) + ks * 2 * X_KB, X_KB, 128), kIdesc, ks != 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[team])) : "memory");
        }
        __syncwarp();
        mbar_wait(&bars[team], parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float SA[16], SB[16], SD[16];
        tmem_ld16(tbase + lane_sel + 64, SA);
        tmem_ld16(tbase + lane_sel + 80, SB);
        tmem_ld16(tbase + lane_sel + 96, SD);                          // SD[0] = sum A, SD[1] = sum A' over the A-branch members
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (tt == i) {
#pragma unroll
            for (int f = 0; f < 16; ++f) TOT[f] = SB[f];
            TOT[16] = SD[1];
        }
        team_barrier(team);
        // ---- this star's partial for destination tt (fp32)
        if (live) {
            const float s = ref + er;
            const float c = ex2(-0.8f * fabsf(s));
            const float C1 = s >= 0.f ? 1.f : c, C2 = s >= 0.f ? c : 1.f;
            float M = lrelu(s);
            const bool self_a = __hge(ELH[tt], th16);                  // the branch the MMA put this row's own member in
            float xa[16], xb[16], v[16];
            unpack8(*reinterpret_cast<const uint4 *>(xrow), reinterpret_cast<float(&)[8]>(xa[0]));
            unpack8(*reinterpret_cast<const uint4 *>(xrow + 128), reinterpret_cast<float(&)[8]>(xa[8]));
            unpack8(*reinterpret_cast<const uint4 *>(xrow + 256), reinterpret_cast<float(&)[8]>(xb[0]));
            unpack8(*reinterpret_cast<const uint4 *>(xrow + 384), reinterpret_cast<float(&)[8]>(xb[8]));
            const float sa = self_a ? 1.f : 0.f, sb = 1.f - sa;
#pragma unroll
            for (int f = 0; f < 16; ++f)
                v[f] = fmaf(C1, fmaf(-sa, xa[f], SA[f]), C2 * (fmaf(-sb, xb[f], TOT[f] - SB[f])));
            const float wself = self_a ? C1 * __half2float(A16) : C2 * __half2float(A516);
            float den = fmaf(C1, SD[0], C2 * (TOT[16] - SD[1])) - wself;
            if (fix && tt != a1) {                                     // the leading member joins in fp32 with weight exactly 1
                const float Mrow = lrelu(m1 + er);
                const float sc = ex2(M - Mrow);
#pragma unroll
                for (int f = 0; f < 16; ++f) v[f] = fmaf(v[f], sc, TOT[20 + f]);
                den = fmaf(den, sc, 1.f);
                M = Mrow;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4 *>(srow + head * F_ + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            *reinterpret_cast<float2 *>(srow + D_ + 2 * head) = make_float2(den, M);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                   // the stash is complete
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tslot), "n"(TMEM_COLS) : "memory");

    // ---------------------------------------------------------------- phase 2a: publish rows j > i, release this star's flag
    for (int j = i + 1 + warp; j < n; j += TC_THREADS / 32) {
        const size_t node = (size_t)(node0 + NODE[j]);
        const float *sr = STASH + j * SROW;
        __stcg(reinterpret_cast<float4 *>(a.recV + node * D_) + lane, *reinterpret_cast<const float4 *>(sr + 4 * lane));
        if (lane < 4) __stcg(reinterpret_cast<float4 *>(a.recDM + node * 2 * H_) + lane, *reinterpret_cast<const float4 *>(sr + D_ + 4 * lane));
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        st_release_gpu(a.flags + (size_t)b * n + i, 1);
    }
    if (i == 0) return;
    // ---------------------------------------------------------------- phase 2b: rows j < i -- merge with the lower stars' records, finish
    const float4 sc4 = __ldg(reinterpret_cast<const float4 *>(a.bn_scale) + lane), sh4 = __ldg(reinterpret_cast<const float4 *>(a.bn_shift) + lane);
    const float4 bb4 = a.bias ? __ldg(reinterpret_cast<const float4 *>(a.bias) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < i) wait_flag(a.flags + (size_t)b * n + tid);             // stars tid < i were dispatched earlier and never wait before publishing
    __syncthreads();
    constexpr int RU = 3;                                              // rows in flight per warp
    for (int j0 = warp * RU; j0 < i; j0 += (TC_THREADS / 32) * RU) {
        float4 pv[RU], hv[RU];
        float2 pdm[RU];
        size_t node[RU];
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            const int j = min(j0 + r, i - 1);
            node[r] = (size_t)(node0 + NODE[j]);
            pv[r] = __ldcg(reinterpret_cast<const float4 *>(a.recV + node[r] * D_) + lane);
            pdm[r] = __ldcg(reinterpret_cast<const float2 *>(a.recDM + node[r] * 2 * H_) + (lane >> 2));
            hv[r] = __ldg(reinterpret_cast<const float4 *>(a.h + node[r] * D_) + lane);
        }
#pragma unroll
        for (int r = 0; r < RU; ++r) {
            if (j0 + r >= i) break;
            const float *sr = STASH + (j0 + r) * SROW;
            const float4 v = *reinterpret_cast<const float4 *>(sr + 4 * lane);
            const float2 dm = *reinterpret_cast<const float2 *>(sr + D_ + 2 * (lane >> 2));
            // flash-style merge, always (lower star, higher star): independent of timing
            const float mx = fmaxf(pdm[r].y, dm.y);
            const float s1 = ex2(pdm[r].y - mx), s2 = ex2(dm.y - mx);
            const float inv = 1.f / fmaf(pdm[r].x, s1, dm.x * s2);
            const float a1 = s1 * inv, a2 = s2 * inv;
            float4 o;
            o.x = (hv[r].x + (fmaf(pv[r].x, a1, v.x * a2) + bb4.x)) * sc4.x + sh4.x;
            o.y = (hv[r].y + (fmaf(pv[r].y, a1, v.y * a2) + bb4.y)) * sc4.y + sh4.y;
            o.z = (hv[r].z + (fmaf(pv[r].z, a1, v.z * a2) + bb4.z)) * sc4.z + sh4.z;
            o.w = (hv[r].w + (fmaf(pv[r].w, a1, v.w * a2) + bb4.w)) * sc4.w + sh4.w;
            reinterpret_cast<float4 *>(a.h1 + node[r] * D_)[lane] = o;
            if (a.h1_tf32) reinterpret_cast<float4 *>(a.h1_tf32 + node[r] * D_)[lane] = tf32_round4(o);
            // the consumed numerator record is dead (read exactly once): drop its dirty L2 lines instead of writing them back
            if ((lane & 7) == 0) discard_l2_128(a.recV + node[r] * D_ + 4 * lane);
        }
    }
}

}  // namespace

namespace gnngls {
int launch_kn_tc(const KnArgs &args, int B, cudaStream_t st) {
    const TcLayout L(args.n);
    GNNGLS_REQUIRE(args.n <= KPAD, GNNGLS_ERR_UNSUPPORTED, "the tcgen05 K_n kernel handles n <= %d", KPAD);
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(gat_kn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(gat_kn_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t grid = (int64_t)B * args.n;
    GNNGLS_REQUIRE(grid < ((int64_t)1 << 31), GNNGLS_ERR_UNSUPPORTED, "B*n too large for one launch");
    gat_kn_tc_kernel<<<(unsigned)grid, TC_THREADS, L.total, st>>>(args);
    GNNGLS_LAUNCH_OK("gat_kn_tc_kernel");
    return GNNGLS_OK;
}
}  // namespace gnngls
