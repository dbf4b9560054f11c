// GAT edge-softmax / aggregate for a batch of line graphs of K_n (dgl.nn.GATConv reached from
// gnngls/models.py:23; semantics in SURVEY.md Appendix A), fused with the skip connection and the first
// BatchNorm of the layer (models.py:12-15,27):
//
//     h1[v] = BN1( h[v] + sum_u softmax_u( leaky_relu(el[u] + er[v], 0.2) ) * ft[u]  (+ bias) )
//
// The adjacency is computed arithmetically: node {i,j} receives from the "stars" of vertex i and of
// vertex j (all TSP edges incident to the vertex), itself excluded.  Inside one star, for one head, the
// work is NOT quadratic: leaky_relu is piecewise linear, so with s = el_k + er_j (log2 domain)
//
//     2^leaky_relu(s) = 2^el_k * 2^er_j          if el_k >= -er_j      (branch A)
//                     = 2^(.2 el_k) * 2^(.2 er_j) otherwise             (branch B)
//
// i.e. with the star's members sorted by el, destination j sees a rank-1 weight matrix on each side of
// one threshold rank r_j = #{k : el_k < -er_j}.  Its aggregate over the whole star is therefore two table
// look-ups into prefix sums over the sorted members,
//
//     num_j = C1_j * SufA[r_j] + C2_j * PreB[r_j] - w_jj * ft_j ,
//     SufA[r] = sum_{rank >= r} A_k ft_k,   A_k  = 2^(el_k - m1)        (m1 = max_k el_k)
//     PreB[r] = sum_{rank <  r} A'_k ft_k,  A'_k = 2^(.2 (el_k - m1))
//     C1_j = 2^(m1 + er_j - mx_j), C2_j = 2^(.2 (m1 + er_j) - mx_j),  mx_j = leaky_relu(m1 + er_j)
//
// all factors in [0,1], everything in fp32: O(n) work per (star, head, feature) instead of O(n^2), no
// per-edge exponential and no contraction left for a tensor core to do (a row gather by a data-dependent
// rank is what remains).  Only the arg-max member's own row can lose precision to the self-exclusion
// (its own weight may dominate the sums it is subtracted from): when it holds > 90 % of its row that one
// row per (star, head) is evaluated directly.
//
// One CTA per (instance, vertex i, group of G heads); W warps per head:
//   1. stage the star's feature rows (cp.async, 16-byte pieces, head chunks XOR-swizzled by row) and scores;
//   2. sort the members by el: packed (25-bit key | slot) warp bitonic sort, then exact-order verification
//      with odd-even transposition on the full fp32 values (ties of the truncated key);
//   3. per destination: threshold rank by binary search, C factors, self weight;
//   4. chunked scans over the sorted members -> tables PreB / SufA (+ denominators) in shared memory;
//   5. per destination and 4 features: combine the two table rows.  Destination {i,j} belongs to two stars:
//      the star of the LOWER vertex publishes its partial (numerator, denominator, reference max) and raises
//      one flag per star; the star of the HIGHER vertex waits for the flags of the lower stars (they were
//      dispatched earlier: blockIdx order), merges in fixed (lower, higher) order -- deterministic and
//      batching-invariant bitwise -- applies bias + skip + BN1 and writes h1.  The consumed records are
//      discarded from L2 so that they never travel to HBM.
#include "gat_kn.cuh"

namespace {

// Shared-memory layout.  Per head: the table region T_h (rows 0..m = prefix tables, row m+1 = exact arg-max row)
// doubles as the home of the arrays that are dead before the scan starts (scores, sorted scores, ranks); the
// region S_h holds the scan inputs (row offsets + weights in rank order) and, after the scan, the per-destination
// constants (s_j, w_jj, r_j).
template <typename FT, int G, int W, int EPL>
struct KnLayout {
    static constexpr int NP = 32 * EPL;                       // sort capacity (slots >= n)
    static constexpr int ROWB = G * F_ * (int)sizeof(FT);     // bytes of one staged feature row
    int rows;                                                 // m + 2
    unsigned t_head, t_off, dn_head, dn_off, ft_off, s_head, s_off, wt_off, misc_off, node_off, total;
    __host__ __device__ explicit KnLayout(int n) {
        rows = n + 1;
        const unsigned a = (unsigned)rows * 128u, b = 14u * NP;
        t_head = ((a > b ? a : b) + 15u) & ~15u;
        t_off = 0;
        dn_head = (unsigned)rows * 8u;
        dn_off = t_off + G * t_head;
        ft_off = (dn_off + G * dn_head + 15u) & ~15u;
        s_head = 10u * NP;                                    // [0,8NP): float2 per rank / destination; [8NP,10NP): u16
        s_off = (ft_off + (unsigned)n * ROWB + 15u) & ~15u;
        wt_off = s_off + G * s_head;                          // [G][W][8][5] floats: per-warp scan totals
        misc_off = wt_off + G * W * 40u * 4u;                 // [G][4] words: m1, m2, jstar, mxfix
        node_off = misc_off + G * 16u;
        total = node_off + (unsigned)n * 4u;
    }
};

template <typename FT, int G, int W, int EPL, int MINB>
__global__ void __launch_bounds__(32 * G * W, MINB) gat_kn_scan_kernel(const KnArgs a) {
    constexpr int HG = H_ / G, THREADS = 32 * G * W, NP = 32 * EPL, GL = 32 * W;
    constexpr int ROWB = G * F_ * (int)sizeof(FT), HB = F_ * (int)sizeof(FT);
    constexpr int LOG_ROWB = ROWB == 32 ? 5 : ROWB == 64 ? 6 : ROWB == 128 ? 7 : 8;
    constexpr int RP = (NP + GL - 1) / GL;                    // destinations per thread
    constexpr uint32_t MASK = NP - 1;
    static_assert(EPL >= 2 && (EPL & (EPL - 1)) == 0, "EPL must be a power of two >= 2");
    static_assert((1 << LOG_ROWB) == ROWB, "row bytes");
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, m = n - 1;
    const KnLayout<FT, G, W, EPL> L(n);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hgl = warp / W, hw = warp % W, hl = hw * 32 + lane;      // head within the group, warp / lane within the head
    const int hg = blockIdx.x % HG, star = blockIdx.x / HG;
    const int b = star / n, i = star - b * n;
    const int head = hg * G + hgl;
    const int64_t N = (int64_t)n * (n - 1) / 2, node0 = (int64_t)b * N;
    const int bar_id = 1 + hgl;

    unsigned char *Tb = smem + L.t_off + hgl * L.t_head;
    float *Th = reinterpret_cast<float *>(Tb);                         // [rows][32]: PreB (16) | SufA (16)
    float *EL = reinterpret_cast<float *>(Tb);                         // phase-A aliases of the table region
    float *ER = EL + NP;
    float *SEL = ER + NP;
    uint16_t *RANK = reinterpret_cast<uint16_t *>(SEL + NP);
    float2 *DNh = reinterpret_cast<float2 *>(smem + L.dn_off + hgl * L.dn_head);   // [rows] (PreB den, SufA den)
    unsigned char *FTs = smem + L.ft_off;
    unsigned char *Sb = smem + L.s_off + hgl * L.s_head;
    float2 *AA = reinterpret_cast<float2 *>(Sb);                       // scan: (A, A') by rank
    uint16_t *PERMOFF = reinterpret_cast<uint16_t *>(Sb + 8 * NP);     // scan: feature-row byte offset by rank
    float2 *SW = reinterpret_cast<float2 *>(Sb);                       // consumer: (s_j, w_jj) by destination
    uint16_t *RJ = reinterpret_cast<uint16_t *>(Sb + 8 * NP);          // consumer: table row of destination j
    float *WT = reinterpret_cast<float *>(smem + L.wt_off) + hgl * (W * 40);
    float *MISC = reinterpret_cast<float *>(smem + L.misc_off) + hgl * 4;
    int *NODE = reinterpret_cast<int *>(smem + L.node_off);

    // ---------------------------------------------------------------- 1. stage the star of vertex i
    for (int k = tid; k < n; k += THREADS) NODE[k] = (k != i) ? kn_node(i, k, n) : -1;
    __syncthreads();
    {
        constexpr int PPR = ROWB / 16, PPH = HB / 16;                  // 16-byte pieces per row / per head chunk
        const unsigned char *ftg = static_cast<const unsigned char *>(a.ft);
        for (int idx = tid; idx < n * PPR; idx += THREADS) {
            const int k = idx / PPR, p = idx - k * PPR;
            const int node = NODE[k];
            const int cs = (p / PPH) ^ (k & (G - 1));                  // swizzled head chunk: rows k..k+G-1 hit distinct banks
            unsigned char *dst = FTs + (size_t)k * ROWB + (cs * PPH + (p % PPH)) * 16;
            if (node >= 0) cp_async16(dst, ftg + ((size_t)(node0 + node) * D_ + hg * G * F_) * sizeof(FT) + p * 16);
            else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
        float *EL0 = reinterpret_cast<float *>(smem + L.t_off);
        for (int k = tid; k < NP; k += THREADS) {
            const int node = k < n ? NODE[k] : -1;
            float e[G], r[G];
#pragma unroll
            for (int g = 0; g < G; ++g) { e[g] = 0.f; r[g] = 0.f; }
            if (node >= 0) {
                const float *pe = a.el + (size_t)(node0 + node) * H_ + hg * G, *pr = a.er + (size_t)(node0 + node) * H_ + hg * G;
                if (G == 4) {
                    const float4 x = *reinterpret_cast<const float4 *>(pe), y = *reinterpret_cast<const float4 *>(pr);
                    e[0] = x.x; e[1 % G] = x.y; e[2 % G] = x.z; e[3 % G] = x.w;
                    r[0] = y.x; r[1 % G] = y.y; r[2 % G] = y.z; r[3 % G] = y.w;
                } else if (G == 2) {
                    const float2 x = *reinterpret_cast<const float2 *>(pe), y = *reinterpret_cast<const float2 *>(pr);
                    e[0] = x.x; e[1 % G] = x.y; r[0] = y.x; r[1 % G] = y.y;
                } else {
#pragma unroll
                    for (int g = 0; g < G; ++g) { e[g] = pe[g]; r[g] = pr[g]; }
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float *base = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(EL0) + g * L.t_head);
                base[k] = e[g];
                base[NP + k] = r[g];
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---------------------------------------------------------------- 2. sort the members by el (one warp per head)
    if (hw == 0) {
        uint32_t v[EPL];
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            const int slot = lane * EPL + q;
            const bool live = slot < n && slot != i;
            const uint32_t u = __float_as_uint(EL[slot]);
            const uint32_t ord = u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);   // monotone in the float order
            v[q] = (live ? (ord & ~MASK) : ~MASK) | (uint32_t)slot;           // dead / padding slots sort last
        }
        // bitonic network over e = lane*EPL + q
#pragma unroll
        for (int lk = 1; (1 << lk) <= NP; ++lk) {
            const int k = 1 << lk;
#pragma unroll
            for (int lj = lk - 1; lj >= 0; --lj) {
                const int j = 1 << lj;
                if (j >= EPL) {
                    const int lanej = j / EPL;
                    const bool up = ((lane * EPL) & k) == 0;
                    const bool lower = (lane & lanej) == 0;
                    const bool take_min = up == lower;
#pragma unroll
                    for (int q = 0; q < EPL; ++q) {
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, v[q], lanej);
                        v[q] = take_min ? min(v[q], o) : max(v[q], o);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < EPL; ++q) {
                        if ((q & j) == 0) {
                            const bool up = ((lane * EPL + q) & k) == 0;
                            const uint32_t lo = min(v[q], v[q | j]), hi = max(v[q], v[q | j]);
                            v[q] = up ? lo : hi;
                            v[q | j] = up ? hi : lo;
                        }
                    }
                }
            }
        }
        // exact order: the key dropped the low mantissa bits, so members whose el agree in the kept bits came out in
        // slot order.  Odd-even transposition on the full values until nothing moves (usually zero or one pass).
        float ev[EPL];
        int sv[EPL];
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            sv[q] = (int)(v[q] & MASK);
            ev[q] = (sv[q] < n && sv[q] != i) ? EL[sv[q]] : INFINITY;
        }
        bool any;
        do {
            bool swapped = false;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
                for (int q = ph; q + 1 < EPL; q += 2) {
                    if (ev[q] > ev[q + 1]) {
                        const float te = ev[q]; ev[q] = ev[q + 1]; ev[q + 1] = te;
                        const int ts = sv[q]; sv[q] = sv[q + 1]; sv[q + 1] = ts;
                        swapped = true;
                    }
                }
            }
            const float nx = __shfl_down_sync(0xffffffffu, ev[0], 1), pv = __shfl_up_sync(0xffffffffu, ev[EPL - 1], 1);
            const int nxs = __shfl_down_sync(0xffffffffu, sv[0], 1), pvs = __shfl_up_sync(0xffffffffu, sv[EPL - 1], 1);
            const bool sw_hi = lane < 31 && ev[EPL - 1] > nx, sw_lo = lane > 0 && pv > ev[0];
            if (sw_hi) { ev[EPL - 1] = nx; sv[EPL - 1] = nxs; swapped = true; }
            if (sw_lo) { ev[0] = pv; sv[0] = pvs; swapped = true; }
            any = __any_sync(0xffffffffu, swapped);
        } while (any);
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            const int rank = lane * EPL + q;
            if (rank < m) {
                SEL[rank] = ev[q];
                RANK[sv[q]] = (uint16_t)rank;
                PERMOFF[rank] = (uint16_t)(sv[q] * ROWB + ((hgl ^ (sv[q] & (G - 1))) * HB));
            }
        }
        __syncwarp();
        const float m1 = SEL[m - 1], m2 = SEL[m - 2];                  // n >= 3: at least two members
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
            const int rank = lane * EPL + q;
            if (rank < m) {
                const float d = ev[q] - m1;
                AA[rank] = make_float2(ex2(d), ex2(kSlope * d));
            }
        }
        if (lane == 0) {
            MISC[0] = m1;
            MISC[1] = m2;
            MISC[2] = __int_as_float((int)(PERMOFF[m - 1] >> LOG_ROWB));   // slot of the arg-max member
            MISC[3] = 0.f;
        }
    }
    group_barrier<GL>(bar_id);

    // ---------------------------------------------------------------- 3. per destination: threshold rank, C factors, self weight
    float ps[RP], pw[RP];
    int pr[RP];
    {
        const float m1 = MISC[0];
#pragma unroll
        for (int t = 0; t < RP; ++t) {
            const int j = hl + t * GL;
            ps[t] = 0.f; pw[t] = 0.f; pr[t] = 0;
            if (j < n && j != i) {
                const float er = ER[j], th = -er;
                int lo = 0, len = m;                                   // lower bound: #{rank : SEL[rank] < th}
                while (len > 0) {
                    const int half = len >> 1;
                    const bool lt = SEL[lo + half] < th;
                    lo = lt ? lo + half + 1 : lo;
                    len = lt ? len - half - 1 : half;
                }
                const float s = m1 + er;
                const float c = ex2(-0.8f * fabsf(s));
                const float C1 = s >= 0.f ? 1.f : c, C2 = s >= 0.f ? c : 1.f;
                const int rk = RANK[j];
                const float2 aa = AA[rk];
                ps[t] = s;
                pw[t] = rk >= lo ? aa.x * C1 : aa.y * C2;
                pr[t] = lo;
            }
        }
    }
    group_barrier<GL>(bar_id);                                         // the table region is free now

    // ---------------------------------------------------------------- 4. chunked scans over the sorted members
    {
        constexpr int NCH = 4 * W;                                     // rank chunks; 8 lanes (branch x feature quad) per chunk
        const int c8 = hl & 7, br = c8 >> 2, q = c8 & 3, chunk = hl >> 3, cl = lane >> 3;
        const int Lc = (m + NCH - 1) / NCH;
        const int r0 = min(m, chunk * Lc), r1 = min(m, r0 + Lc);
        const int col = br * 16 + q * 4;
        const unsigned char *ftq = FTs + q * 4 * sizeof(FT);
        const float *wsel = reinterpret_cast<const float *>(AA) + (br ? 0 : 1);   // branch A uses A, branch B uses A'
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, dacc = 0.f;
        if (r0 < r1) {
            if (br == 0) {                                             // PreB[r0] (local part): nothing below yet
                *reinterpret_cast<float4 *>(Th + r0 * 32 + col) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q == 0) DNh[r0].x = 0.f;
            } else if (r1 == m) {                                      // SufA[m] = 0
                *reinterpret_cast<float4 *>(Th + m * 32 + col) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q == 0) DNh[m].y = 0.f;
            }
        }
        for (int t = 0; t < r1 - r0; ++t) {
            const int r = br ? (r1 - 1 - t) : (r0 + t);
            const float wgt = wsel[2 * r];
            const float4 f = lds_ft4(ftq + PERMOFF[r], FT());
            acc[0] = fmaf(wgt, f.x, acc[0]); acc[1] = fmaf(wgt, f.y, acc[1]);
            acc[2] = fmaf(wgt, f.z, acc[2]); acc[3] = fmaf(wgt, f.w, acc[3]);
            dacc += wgt;
            const int row = br ? r : r + 1;                            // SufA is inclusive, PreB exclusive
            if (br || r + 1 < r1 || r1 == m) {                         // (row r1 of a B chunk belongs to the next chunk)
                *reinterpret_cast<float4 *>(Th + row * 32 + col) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                if (q == 0) reinterpret_cast<float *>(DNh + row)[br] = dacc;
            }
        }
        // offsets: totals of the chunks below (B) / above (A); fixed summation order
        float off[4] = {0.f, 0.f, 0.f, 0.f}, doff = 0.f;
#pragma unroll
        for (int d = 1; d <= 3; ++d) {
            const int src = br ? lane + 8 * d : lane - 8 * d;
            const bool valid = br ? (cl + d <= 3) : (cl - d >= 0);
            const float t0 = __shfl_sync(0xffffffffu, acc[0], src & 31), t1 = __shfl_sync(0xffffffffu, acc[1], src & 31);
            const float t2 = __shfl_sync(0xffffffffu, acc[2], src & 31), t3 = __shfl_sync(0xffffffffu, acc[3], src & 31);
            const float td = __shfl_sync(0xffffffffu, dacc, src & 31);
            if (valid) { off[0] += t0; off[1] += t1; off[2] += t2; off[3] += t3; doff += td; }
        }
        if (W > 1) {
            if (cl == (br ? 0 : 3)) {                                  // this lane's offset + own chunk = the warp's total
                float *wt = WT + hw * 40 + c8 * 5;
                wt[0] = off[0] + acc[0]; wt[1] = off[1] + acc[1]; wt[2] = off[2] + acc[2]; wt[3] = off[3] + acc[3];
                wt[4] = doff + dacc;
            }
            group_barrier<GL>(bar_id);
            for (int ow = 0; ow < W; ++ow) {
                if (br ? (ow > hw) : (ow < hw)) {
                    const float *wt = WT + ow * 40 + c8 * 5;
                    off[0] += wt[0]; off[1] += wt[1]; off[2] += wt[2]; off[3] += wt[3]; doff += wt[4];
                }
            }
        }
        const bool first_chunk = br ? (r1 == m) : (chunk == 0);        // nothing above / below: offset is zero
        if (r0 < r1 && !first_chunk) {
            const int ra = r0, rb = br ? r1 - 1 : (r1 == m ? m : r1 - 1);
            for (int row = ra; row <= rb; ++row) {
                float4 x = *reinterpret_cast<float4 *>(Th + row * 32 + col);
                x.x += off[0]; x.y += off[1]; x.z += off[2]; x.w += off[3];
                *reinterpret_cast<float4 *>(Th + row * 32 + col) = x;
                if (q == 0) reinterpret_cast<float *>(DNh + row)[br] += doff;
            }
        }
    }
    group_barrier<GL>(bar_id);                                         // tables complete; AA / PERMOFF dead
#pragma unroll
    for (int t = 0; t < RP; ++t) {
        const int j = hl + t * GL;
        if (j < n && j != i) {
            SW[j] = make_float2(ps[t], pw[t]);
            RJ[j] = (uint16_t)pr[t];
        }
    }
    group_barrier<GL>(bar_id);

    // ---------------------------------------------------------------- 4b. the arg-max member's own row, when its self weight dominates
    if (hw == 0) {
        const int js = __float_as_int(MISC[2]);
        const float2 sw = SW[js];
        const int r = RJ[js];
        const float c = ex2(-0.8f * fabsf(sw.x));
        const float C1 = sw.x >= 0.f ? 1.f : c, C2 = sw.x >= 0.f ? c : 1.f;
        const float2 dn = DNh[r];
        const float all = fmaf(C1, dn.y, C2 * dn.x);
        if (sw.y > kFixFrac * all) {                                   // warp-uniform
            const float erj = __ldg(a.er + (size_t)(node0 + NODE[js]) * H_ + head);
            const float mxf = lrelu(MISC[1] + erj);                    // reference: the runner-up bounds every remaining member
            const int mg = lane >> 2, fq = lane & 3;
            float num[4] = {0.f, 0.f, 0.f, 0.f}, den = 0.f;
            for (int k = mg; k < n; k += 8) {
                if (k == i || k == js) continue;
                const float elk = __ldg(a.el + (size_t)(node0 + NODE[k]) * H_ + head);
                const float wk = ex2(lrelu(elk + erj) - mxf);
                const float4 f = lds_ft4(FTs + (size_t)k * ROWB + ((hgl ^ (k & (G - 1))) * HB) + fq * 4 * sizeof(FT), FT());
                num[0] = fmaf(wk, f.x, num[0]); num[1] = fmaf(wk, f.y, num[1]);
                num[2] = fmaf(wk, f.z, num[2]); num[3] = fmaf(wk, f.w, num[3]);
                den += wk;
            }
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
                num[0] += __shfl_xor_sync(0xffffffffu, num[0], o); num[1] += __shfl_xor_sync(0xffffffffu, num[1], o);
                num[2] += __shfl_xor_sync(0xffffffffu, num[2], o); num[3] += __shfl_xor_sync(0xffffffffu, num[3], o);
                den += __shfl_xor_sync(0xffffffffu, den, o);
            }
            if (mg == 0) *reinterpret_cast<float4 *>(Th + (m + 1) * 32 + fq * 4) = make_float4(num[0], num[1], num[2], num[3]);
            if (mg == 1) *reinterpret_cast<float4 *>(Th + (m + 1) * 32 + 16 + fq * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane == 0) {
                DNh[m + 1] = make_float2(den, 0.f);
                MISC[3] = mxf;
                SW[js].y = 0.f;
                RJ[js] = (uint16_t)(m + 1);
            }
        }
    }
    group_barrier<GL>(bar_id);

    // ---------------------------------------------------------------- 5. destinations
    const int q = hl & 3;                                              // feature quad of this lane (GL is a multiple of 4)
    const int fcol = head * F_ + q * 4;
    const float mxfix = MISC[3];
    // this star's partial for destination j: numerator of 4 features, denominator, reference max
    auto partial = [&](int j, float (&v)[4], float &den, float &M) {
        const float2 sw = SW[j];
        const int r = RJ[j];
        const float c = ex2(-0.8f * fabsf(sw.x));
        float C1 = sw.x >= 0.f ? 1.f : c, C2 = sw.x >= 0.f ? c : 1.f;
        M = lrelu(sw.x);
        if (r == m + 1) { C1 = 0.f; C2 = 1.f; M = mxfix; }
        const float *row = Th + r * 32 + q * 4;
        const int par = j & 1;                                         // neighbouring destinations read opposite halves first: no bank conflicts
        const float4 x = *reinterpret_cast<const float4 *>(row + (par ? 16 : 0));
        const float4 y = *reinterpret_cast<const float4 *>(row + (par ? 0 : 16));
        const float cx = par ? C1 : C2, cy = par ? C2 : C1;
        const float4 f = lds_ft4(FTs + (size_t)j * ROWB + ((hgl ^ (j & (G - 1))) * HB) + q * 4 * sizeof(FT), FT());
        const float w = sw.y;
        v[0] = fmaf(cx, x.x, fmaf(cy, y.x, -w * f.x));
        v[1] = fmaf(cx, x.y, fmaf(cy, y.y, -w * f.y));
        v[2] = fmaf(cx, x.z, fmaf(cy, y.z, -w * f.z));
        v[3] = fmaf(cx, x.w, fmaf(cy, y.w, -w * f.w));
        const float2 dn = DNh[r];
        den = fmaf(C1, dn.y, C2 * dn.x) - w;
    };

    // 5a. higher partners j > i: this star is the lower one -> publish
    for (int e = hl; e < 4 * (m - i); e += GL) {
        const int j = i + 1 + (e >> 2);
        float v[4], den, M;
        partial(j, v, den, M);
        const size_t node = (size_t)(node0 + NODE[j]);
        __stcg(reinterpret_cast<float4 *>(a.recV + node * D_ + fcol), make_float4(v[0], v[1], v[2], v[3]));
        if (q == 0) __stcg(reinterpret_cast<float2 *>(a.recDM + (node * H_ + head) * 2), make_float2(den, M));
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        st_release_gpu(a.flags + (size_t)star * HG + hg, 1);
    }
    // 5b. lower partners j < i: their stars were dispatched before this one (blockIdx order) and never wait
    // before publishing, so this cannot deadlock; fail loudly instead of hanging if it ever takes seconds
    if (i > 0) {
        for (int t = tid; t < i; t += THREADS) {
            const int *f = a.flags + ((size_t)b * n + t) * HG + hg;
            unsigned long long t0 = 0;
            while (ld_acquire_gpu(f) == 0) {
                __nanosleep(64);
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t0 == 0) t0 = t1;
                else if (t1 - t0 > 4000000000ull) asm volatile("trap;");
            }
        }
        __syncthreads();
        const float4 sc = *reinterpret_cast<const float4 *>(a.bn_scale + fcol);
        const float4 sh = *reinterpret_cast<const float4 *>(a.bn_shift + fcol);
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias) bb = *reinterpret_cast<const float4 *>(a.bias + fcol);
        for (int e = hl; e < 4 * i; e += GL) {
            const int j = e >> 2;
            const size_t node = (size_t)(node0 + NODE[j]);
            const float4 pv = __ldcg(reinterpret_cast<const float4 *>(a.recV + node * D_ + fcol));
            const float2 pdm = __ldcg(reinterpret_cast<const float2 *>(a.recDM + (node * H_ + head) * 2));
            const float4 hv = __ldg(reinterpret_cast<const float4 *>(a.h + node * D_ + fcol));
            float v[4], den, M;
            partial(j, v, den, M);
            // flash-style merge, always (lower star, higher star): independent of timing
            const float mx = fmaxf(pdm.y, M);
            const float s1 = ex2(pdm.y - mx), s2 = ex2(M - mx);
            const float inv = 1.f / fmaf(pdm.x, s1, den * s2);
            const float a1 = s1 * inv, a2 = s2 * inv;
            float4 o;
            o.x = (hv.x + (fmaf(pv.x, a1, v[0] * a2) + bb.x)) * sc.x + sh.x;
            o.y = (hv.y + (fmaf(pv.y, a1, v[1] * a2) + bb.y)) * sc.y + sh.y;
            o.z = (hv.z + (fmaf(pv.z, a1, v[2] * a2) + bb.z)) * sc.z + sh.z;
            o.w = (hv.w + (fmaf(pv.w, a1, v[3] * a2) + bb.w)) * sc.w + sh.w;
            *reinterpret_cast<float4 *>(a.h1 + node * D_ + fcol) = o;
            if (a.h1_tf32) *reinterpret_cast<float4 *>(a.h1_tf32 + node * D_ + fcol) = tf32_round4(o);
        }
        if constexpr (G >= 2) {
            // the consumed numerator records are dead (each is read exactly once): drop their dirty L2 lines
            // instead of writing them back to HBM.  One CTA owns whole 128-byte lines only when G >= 2.
            __syncthreads();
            constexpr int LINES = G * F_ * 4 / 128;
            for (int idx = tid; idx < i * LINES; idx += THREADS) {
                const int j = idx / LINES, ln = idx - j * LINES;
                discard_l2_128(a.recV + (size_t)(node0 + NODE[j]) * D_ + hg * G * F_ + ln * 32);
            }
        }
    }
}

template <typename FT, int G, int W, int EPL, int MINB>
int launch_cfg(const KnArgs &args, int B, cudaStream_t st) {
    const KnLayout<FT, G, W, EPL> L(args.n);
    GNNGLS_REQUIRE(L.total <= (unsigned)gnngls::device_max_optin_smem(), GNNGLS_ERR_UNSUPPORTED,
                   "n=%d: a vertex star (%u B) does not fit shared memory; use the CSR path", args.n, L.total);
    auto kernel = gat_kn_scan_kernel<FT, G, W, EPL, MINB>;
    if (L.total > 48 * 1024)
        GNNGLS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    // several CTAs per SM only fit under the largest shared-memory carve-out
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t grid = (int64_t)B * args.n * (H_ / G);
    GNNGLS_REQUIRE(grid < ((int64_t)1 << 31), GNNGLS_ERR_UNSUPPORTED, "B*n too large for one launch");
    kernel<<<(unsigned)grid, 32 * G * W, L.total, st>>>(args);
    GNNGLS_LAUNCH_OK("gat_kn_scan_kernel");
    return GNNGLS_OK;
}

// GNNGLS_KN_WARPS_PER_HEAD=4 selects the 512-thread variant for 64 < n <= 128 (A/B knob; default 2)
int kn_warps_per_head() {
    static const int w = [] {
        const char *e = getenv("GNNGLS_KN_WARPS_PER_HEAD");
        return (e && atoi(e) == 4) ? 4 : 2;
    }();
    return w;
}

template <typename FT>
int launch_for_n(const KnArgs &args, int B, cudaStream_t st) {
    const int n = args.n;
    if (n <= 64) return launch_cfg<FT, 4, 2, 2, 3>(args, B, st);
    if (n <= 128) {
        if (kn_warps_per_head() == 4) return launch_cfg<FT, 4, 4, 4, 2>(args, B, st);
        return launch_cfg<FT, 4, 2, 4, 3>(args, B, st);
    }
    if (n <= 256) return launch_cfg<FT, 2, 4, 8, 2>(args, B, st);
    if (n <= 512) return launch_cfg<FT, 1, 8, 16, 2>(args, B, st);
    return launch_cfg<FT, 1, 8, 32, 1>(args, B, st);
}

}  // namespace

extern "C" size_t gnngls_gat_kn_workspace_bytes(int B, int n) {
    if (B <= 0 || n <= 0) return 0;
    const size_t M = (size_t)B * ((size_t)n * (n - 1) / 2);
    // numerator records + (denominator, max) records + one flag per (star, head group); the tcgen05 kernel keeps two records per node
    const size_t scan = sizeof(float) * M * (D_ + 2 * H_) + sizeof(int) * (size_t)B * n * H_;
    const size_t tc = n <= 128 ? gnngls::kn_tc_workspace_bytes(B, n) : 0;
    return scan > tc ? scan : tc;
}

extern "C" int gnngls_gat_aggregate_kn(int B, int n, const void *ft, int ft_dtype, const float *el, const float *er,
                                       const float *h, const float *gat_bias, const float *bn_scale,
                                       const float *bn_shift, float *h1, float *h1_tf32, void *workspace,
                                       size_t workspace_bytes, void *stream) {
    GNNGLS_REQUIRE(ft && el && er && h && bn_scale && bn_shift && h1, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(ft_dtype == GNNGLS_FT_F32 || ft_dtype == GNNGLS_FT_TF32 || ft_dtype == GNNGLS_FT_F16, GNNGLS_ERR_BAD_ARG,
                   "unknown ft_dtype %d", ft_dtype);
    GNNGLS_REQUIRE(n >= 3, GNNGLS_ERR_UNSUPPORTED, "line graph of K_n needs n >= 3 (got %d)", n);
    GNNGLS_REQUIRE(n <= 1024, GNNGLS_ERR_UNSUPPORTED, "n=%d: the K_n path handles n <= 1024; use the CSR path", n);
    if (B <= 0) return GNNGLS_OK;
    GNNGLS_REQUIRE(workspace && workspace_bytes >= gnngls_gat_kn_workspace_bytes(B, n), GNNGLS_ERR_WORKSPACE,
                   "gat_kn workspace too small: need %zu bytes", gnngls_gat_kn_workspace_bytes(B, n));
    GNNGLS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0, GNNGLS_ERR_BAD_ARG,
                   "gat_kn workspace must be 128-byte aligned");
    GNNGLS_REQUIRE(((reinterpret_cast<uintptr_t>(ft) | reinterpret_cast<uintptr_t>(el) | reinterpret_cast<uintptr_t>(er) |
                     reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(h1) | reinterpret_cast<uintptr_t>(h1_tf32) |
                     reinterpret_cast<uintptr_t>(bn_scale) | reinterpret_cast<uintptr_t>(bn_shift) |
                     reinterpret_cast<uintptr_t>(gat_bias)) & 15) == 0,
                   GNNGLS_ERR_BAD_ARG, "gat_kn tensors must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t M = (size_t)B * ((size_t)n * (n - 1) / 2);
    KnArgs args;
    args.n = n;
    args.ft = ft; args.el = el; args.er = er;
    args.recV = static_cast<float *>(workspace);
    args.recDM = args.recV + M * D_;
    args.flags = reinterpret_cast<int *>(args.recDM + M * 2 * H_);
    args.h = h; args.bias = gat_bias; args.bn_scale = bn_scale; args.bn_shift = bn_shift;
    args.h1 = h1; args.h1_tf32 = h1_tf32;
    // fp16 features, n <= 128: tcgen05 indicator-matrix kernel (gat_kn_tc.cu); otherwise, or with GNNGLS_KN_IMPL=scan,
    // the exact fp32 sorted-prefix kernel of this file
    static const bool force_scan = [] {
        const char *e = getenv("GNNGLS_KN_IMPL");
        return e && (e[0] == 's' || e[0] == 'S');
    }();
    if (ft_dtype == GNNGLS_FT_F16 && n <= 128 && !force_scan) return gnngls::launch_kn_tc(args, B, workspace, st);
    GNNGLS_CUDA_OK(cudaMemsetAsync(args.flags, 0, sizeof(int) * (size_t)B * n * H_, st));
    if (ft_dtype == GNNGLS_FT_F16) return launch_for_n<__half>(args, B, st);
    return launch_for_n<float>(args, B, st);
}
