// Search half of the gnngls hot path on sm_100a: all-pairs / one-to-all move evaluation,
// local_search and the persistent guided_local_search loop, one CTA per instance.
//
// Exactness contract: every fp64 expression is evaluated with explicit
// round-to-nearest intrinsics in the reference's left-to-right association (this file is also
// compiled with -fmad=false); the parallel arg-min over (delta, scan rank) equals the
// reference's sequential first-strict-minimum scan.
//
// Reference lines followed (relative to the reference repository root):
//   gnngls/operators.py:6-147, gnngls/algorithms.py:9-18,111-195, gnngls/__init__.py:17-21.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include <cooperative_groups.h>
#include "common.h"

namespace cg = cooperative_groups;

namespace {

using gnngls::set_error;

// Phase timing of the cluster tier (debug builds only: GNNGLS_GLS_STAMPS=1 python -m gnngls_b200.build --force): thread 0 of every
// CTA accumulates clock64() differences per phase of a sweep; read back with gnngls_debug_gls_stamps() (tools/gls_stamps.py).
#ifdef GLS_STAMPS
__device__ unsigned long long g_gls_stamps[16 * 8];
__device__ __forceinline__ long long *gls_stamp_buf() {
    static __shared__ long long buf[9];      // [8]: last clock
    return buf;
}
#define GLS_STAMP(slot)                                              \
    do {                                                             \
        if (threadIdx.x == 0) {                                      \
            long long *sb__ = gls_stamp_buf();                       \
            const long long now__ = clock64();                       \
            sb__[slot] += now__ - sb__[8];                           \
            sb__[8] = now__;                                         \
        }                                                            \
    } while (0)
#else
#define GLS_STAMP(slot) do { } while (0)
#endif

// ----------------------------------------------------------------------------------------------
// candidate bookkeeping
// ----------------------------------------------------------------------------------------------
struct Best {
    double delta;
    int key;   // scan rank: (i << 16) | j for a2a, j for o2a; < 0 == none
    int pad;
};

// numpy.isclose(0, d): |0 - d| <= atol + rtol * |d|  (operators.py:42)
__device__ __forceinline__ bool close_to_zero(double d) {
    const double ad = fabs(d);
    if (!(ad < INFINITY)) return false;                 // np.isclose(0, +-inf) and np.isclose(0, nan) are False
    return ad <= __dadd_rn(1e-8, __dmul_rn(1e-5, ad));
}

// operators.py:41-46: accept iff delta < best_delta (best starts at 0) and not isclose(0, delta).
// Every thread meets its candidates in increasing scan rank (rows ascending, j ascending), so inside a thread an equal delta
// never replaces the incumbent and, with first_improvement, only the first acceptable candidate is ever taken: one fp64
// compare per candidate, the isclose test only for candidates that would otherwise win.
__device__ __forceinline__ void consider(Best &b, double delta, int key, bool fi) {
    if (delta < b.delta) {                                  // b.delta is 0 until a candidate is accepted, < 0 afterwards
        if ((!fi || b.key < 0) && !close_to_zero(delta)) { b.delta = delta; b.key = key; }
    }
}

// general order on (value, rank): used for the utility arg-max, whose values have any sign
__device__ __forceinline__ Best pick_any(Best a, Best b) {
    if (b.key < 0) return a;
    if (a.key < 0) return b;
    if (a.delta < b.delta) return a;
    if (b.delta < a.delta) return b;
    return (a.key < b.key) ? a : b;
}

// The winner of two move candidates.  An accepted candidate has delta < 0 and key >= 0; "none" is (+0.0, -1).  For non-positive
// doubles "more negative" is "larger bit pattern", and -1 is the largest unsigned key, so the reference's order -- smaller delta,
// then smaller scan rank; with first_improvement the smaller scan rank alone -- is two integer compares, no branches.
__device__ __forceinline__ Best pick(Best a, Best b, bool fi) {
    const unsigned long long ua = (unsigned long long)__double_as_longlong(a.delta), ub = (unsigned long long)__double_as_longlong(b.delta);
    const bool kb = (unsigned)b.key < (unsigned)a.key;
    const bool tb = fi ? kb : (ub > ua || (ub == ua && kb));
    Best r;
    r.delta = tb ? b.delta : a.delta; r.key = tb ? b.key : a.key; r.pad = 0;
    return r;
}

template <bool ANY>
__device__ __forceinline__ Best warp_reduce_best(Best v, bool fi) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Best o;
        o.delta = __shfl_xor_sync(0xffffffffu, v.delta, off);
        o.key = __shfl_xor_sync(0xffffffffu, v.key, off);
        v = ANY ? pick_any(v, o) : pick(v, o, fi);
    }
    return v;
}

// result is returned to every thread; `red` is shared scratch of 32 entries.  Every warp combines the warps' winners itself,
// so there are two block barriers: partial winners visible / scratch free again.
template <bool ANY = false>
__device__ Best block_reduce_best(Best v, bool fi, Best *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_reduce_best<ANY>(v, fi);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    Best w;
    w.delta = 0.0; w.key = -1; w.pad = 0;
    if (lane < nw) w = red[lane];
    w = warp_reduce_best<ANY>(w, fi);
    __syncthreads();   // red may be reused immediately by the caller
    return w;
}

// the same with one barrier: the scratch (2 x 32 entries) is double buffered by the caller's reduction count -- a warp can be at
// most one barrier ahead of the slowest, so the half it overwrites was read two barriers ago
__device__ Best block_reduce_best_db(Best v, bool fi, Best *red2, int &parity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_reduce_best<false>(v, fi);
    Best *r = red2 + parity * 32;
    if (lane == 0) r[warp] = v;
    __syncthreads();
    Best w;
    w.delta = 0.0; w.key = -1; w.pad = 0;
    if (lane < nw) w = r[lane];
    w = warp_reduce_best<false>(w, fi);
    parity ^= 1;
    return w;
}

// ----------------------------------------------------------------------------------------------
// who shares one all-pairs sweep: the warps of one CTA (Solo), or the warps of every CTA of a thread-block
// cluster (Cluster; SURVEY.md section 8(f) rank 3, north star (4) "one CTA or cluster per instance").  Rows i of the
// scan are dealt round-robin to the team's warps; the winners are combined with pick(), a total order on
// (delta, scan rank), so the result does not depend on how the rows were dealt -- it is the reference's
// sequential first-strict-minimum whatever the team.
// ----------------------------------------------------------------------------------------------
constexpr int kMaxCluster = 16;

struct Solo {
    __device__ __forceinline__ int first_row(int warp) const { return warp; }
    __device__ __forceinline__ int row_stride(int nw) const { return nw; }
    int rp = 0;                                               // parity of the double-buffered block reduction
    __device__ __forceinline__ Best reduce(Best v, bool fi, Best *red) { return block_reduce_best_db(v, fi, red + 32, rp); }
    __device__ __forceinline__ bool lead() const { return true; }
    static constexpr bool kDeep = false;
};

// Every CTA of the cluster holds a replica of the tour; each sweeps its share of the rows, writes its winner into
// slot [rank] of EVERY member's exchange buffer through distributed shared memory, and after one cluster barrier all
// members hold the same kMaxCluster candidates and apply the same move to their replica.  The buffer is double
// buffered by the parity of the sweep count: a member can only be one barrier ahead, so the slots it overwrites were
// read two barriers ago.
struct Cluster {
    int rank, size, parity, rp;
    Best *xch;      // [2][kMaxCluster] in this CTA's shared memory
    // Row cache (large n): a lane's gather D[a][t[j]] touches a different 128-byte line for almost every lane, and the L1 tag
    // stage serves about one line per cycle -- at n = 500 that, not latency or bandwidth, is what a sweep costs (ncu: 26 sectors
    // per request).  So a warp first copies the one or two matrix rows its scan row needs into shared memory with coalesced
    // loads (2 lines per instruction) and gathers from there.  `symmetric` (D[a][b] == D[b][a] bitwise, checked once per
    // instance) lets relocate take its column term D[d][b] from row b as well.
    double *rows;   // [warps][2][n] or null
    int symmetric;
    __device__ __forceinline__ int first_row(int warp) const { return warp * size + rank; }
    __device__ __forceinline__ int row_stride(int nw) const { return nw * size; }
    __device__ __forceinline__ bool lead() const { return rank == 0; }   // the member that owns the global results
    // D comes from L2 (cluster barriers invalidate L1) and a warp owns only one or two rows: keep four candidates'
    // worth of gathers in flight per lane
    static constexpr bool kDeep = true;
    __device__ Best reduce(Best v, bool fi, Best *red) {
        GLS_STAMP(1);                                          // rows
        v = block_reduce_best_db(v, fi, red + 32, rp);
        GLS_STAMP(2);                                          // reduction inside the CTA (waits for its slowest warp)
        cg::cluster_group cl = cg::this_cluster();
        Best *mine = xch + parity * kMaxCluster;
        if ((int)threadIdx.x < size) *cl.map_shared_rank(mine + rank, threadIdx.x) = v;
        cl.sync();
        GLS_STAMP(3);                                          // exchange + cluster barrier (waits for the slowest member)
        Best w;                                                // every warp combines the (at most 16) winners itself
        w.delta = 0.0; w.key = -1; w.pad = 0;
        const int lane = threadIdx.x & 31;
        if (lane < size) w = mine[lane];
#pragma unroll
        for (int off = kMaxCluster / 2; off > 0; off >>= 1) {
            Best o;
            o.delta = __shfl_xor_sync(0xffffffffu, w.delta, off);
            o.key = __shfl_xor_sync(0xffffffffu, w.key, off);
            w = pick(w, o, fi);
        }
        w.delta = __shfl_sync(0xffffffffu, w.delta, 0);
        w.key = __shfl_sync(0xffffffffu, w.key, 0);
        parity ^= 1;
        GLS_STAMP(4);                                          // combine
        return w;
    }
};

// ----------------------------------------------------------------------------------------------
// distance-matrix accessors
// ----------------------------------------------------------------------------------------------
struct MatPlain {
    const double *p;
    int ld;
    __device__ __forceinline__ double operator()(int a, int b) const { return p[a * ld + b]; }
};

// algorithms.py:163-164: edge_weight + k * edge_penalties, evaluated per element on read
template <typename PenT>
struct MatPen {
    const double *p;
    const PenT *pen;
    int ld, ldp;
    double k;
    __device__ __forceinline__ double operator()(int a, int b) const {
        return __dadd_rn(p[a * ld + b], __dmul_rn(k, (double)pen[a * ldp + b]));
    }
};

// row cache plumbing: a warp copies one matrix row into its shared-memory buffer with asynchronous 8-byte copies (rows of an
// odd-n matrix are only 8-byte aligned), all in flight at once: one L2 round trip per row instead of one per candidate
__device__ __forceinline__ void stage_row_async(double *dst, const double *src, int n, int lane) {
    if ((n & 1) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {   // (warp-uniform)
        for (int c = 2 * lane; c < n; c += 64) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + c);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + c) : "memory");
        }
        return;
    }
    for (int c = lane; c < n; c += 32) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + c);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + c) : "memory");
    }
}
__device__ __forceinline__ void stage_elem_async(double *dst, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// ----------------------------------------------------------------------------------------------
// all-pairs sweeps (operators.py:32-50, :129-147): one warp per row i, lanes over j
// ----------------------------------------------------------------------------------------------
// two_opt_cost (operators.py:14-29), i<j: ((D[a,c] + D[b,d]) - D[a,b]) - D[c,d]
//   a=t[i] b=t[i-1] c=t[j] d=t[j-1];  E[p] := D[t[p], t[p-1]] so D[a,b]=E[i], D[c,d]=E[j]
template <class M, class T>
__device__ Best sweep_two_opt_a2a(const int *t, int n, const M &D, double *E, bool fi, Best *red, T &team) {
    Best best;
    best.delta = 0.0; best.key = -1; best.pad = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if constexpr (T::kDeep) {
        if (team.rows) {                                       // row cache (see struct Cluster); D is the plain global matrix
            double *ra = team.rows + (size_t)warp * 2 * n, *rb = ra + n;
            int i = 1 + team.first_row(warp);
            // rows with few candidates (the last eighth) do not pay for copying two matrix rows
            bool ahead = i <= n - 3 && (n - 2 - i) * 8 >= n;
            if (ahead) {                                       // the first scan row's matrix rows and E travel together
                stage_row_async(ra, D.p + (size_t)t[i] * D.ld, n, lane);
                stage_row_async(rb, D.p + (size_t)t[i - 1] * D.ld, n, lane);
            }
            for (int p = 1 + threadIdx.x; p <= n - 1; p += blockDim.x) stage_elem_async(E + p, D.p + (size_t)t[p] * D.ld + t[p - 1]);
            stage_commit();
            stage_wait<0>();
            __syncthreads();
            GLS_STAMP(0);                                      // E (+ first rows) landed
            for (; i <= n - 3; i += team.row_stride(nw)) {
                const int a = t[i], b = t[i - 1];
                const double Ei = E[i];
                if ((n - 2 - i) * 8 >= n) {
                    if (!ahead) {
                        __syncwarp();                          // every lane is done with the previous rows
                        stage_row_async(ra, D.p + (size_t)a * D.ld, n, lane);
                        stage_row_async(rb, D.p + (size_t)b * D.ld, n, lane);
                        stage_commit();
                        stage_wait<0>();
                        __syncwarp();
                    }
                    ahead = false;
                    int j = i + 2 + lane;
                    for (; j + 96 <= n - 1; j += 128) {        // four candidates per lane in flight (few warps: latency, not issue, bound)
                        double x[4], y[4], e[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) { x[u] = ra[t[j + 32 * u]]; y[u] = rb[t[j + 32 * u - 1]]; e[u] = E[j + 32 * u]; }
#pragma unroll
                        for (int u = 0; u < 4; ++u) x[u] = __dsub_rn(__dsub_rn(__dadd_rn(x[u], y[u]), Ei), e[u]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) consider(best, x[u], (i << 16) | (j + 32 * u), fi);
                    }
                    for (; j <= n - 1; j += 32) {
                        double x = __dadd_rn(ra[t[j]], rb[t[j - 1]]);
                        x = __dsub_rn(x, Ei);
                        x = __dsub_rn(x, E[j]);
                        consider(best, x, (i << 16) | j, fi);
                    }
                } else {
                    for (int j = i + 2 + lane; j <= n - 1; j += 32) {
                        double x = __dadd_rn(D(a, t[j]), D(b, t[j - 1]));
                        x = __dsub_rn(x, Ei);
                        x = __dsub_rn(x, E[j]);
                        consider(best, x, (i << 16) | j, fi);
                    }
                }
            }
            return team.reduce(best, fi, red);
        }
    }
    for (int p = 1 + threadIdx.x; p <= n - 1; p += blockDim.x) E[p] = D(t[p], t[p - 1]);
    __syncthreads();
    for (int i = 1 + team.first_row(warp); i <= n - 3; i += team.row_stride(nw)) {
        const int a = t[i], b = t[i - 1];
        const double Ei = E[i];
        int j = i + 2 + lane;
        if constexpr (T::kDeep) {
            for (; j + 96 <= n - 1; j += 128) {
                double x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = D(a, t[j + 32 * u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = __dadd_rn(x[u], D(b, t[j + 32 * u - 1]));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    x[u] = __dsub_rn(__dsub_rn(x[u], Ei), E[j + 32 * u]);
                    consider(best, x[u], (i << 16) | (j + 32 * u), fi);
                }
            }
        }
        for (; j <= n - 1; j += 32) {
            const int c = t[j], d = t[j - 1];
            double x = __dadd_rn(D(a, c), D(b, d));
            x = __dsub_rn(x, Ei);
            x = __dsub_rn(x, E[j]);
            consider(best, x, (i << 16) | j, fi);
        }
    }
    return team.reduce(best, fi, red);
}

// relocate_cost (operators.py:83-103): (((((-D[a,b]) - D[b,c]) + D[a,c]) - D[d,e]) + D[d,b]) + D[b,e]
//   a=t[i-1] b=t[i] c=t[i+1]; (d,e) = (t[q],t[q+1]) with q = j if i<j else j-1
//   E[p] := D[t[p], t[p+1]] so D[a,b]=E[i-1], D[b,c]=E[i], D[d,e]=E[q]
template <class M, class T>
__device__ Best sweep_relocate_a2a(const int *t, int n, const M &D, double *E, bool fi, Best *red, T &team) {
    Best best;
    best.delta = 0.0; best.key = -1; best.pad = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if constexpr (T::kDeep) {
        if (team.rows) {       // row cache: row b serves D[b][e] and, for a symmetric matrix, D[d][b]; double buffered over scan rows
            double *rbuf = team.rows + (size_t)warp * 2 * n;
            const bool sym = team.symmetric != 0;
            const int stride = team.row_stride(nw);
            int i = 1 + team.first_row(warp), cur = 0;
            if (i <= n - 1) stage_row_async(rbuf, D.p + (size_t)t[i] * D.ld, n, lane);
            for (int p = threadIdx.x; p <= n - 1; p += blockDim.x) stage_elem_async(E + p, D.p + (size_t)t[p] * D.ld + t[p + 1]);
            stage_commit();
            stage_wait<0>();
            __syncthreads();
            GLS_STAMP(0);                                      // E (+ first rows) landed
            for (; i <= n - 1; i += stride) {
                const int a = t[i - 1], b = t[i], c = t[i + 1];
                if (i + stride <= n - 1) stage_row_async(rbuf + (cur ^ 1) * n, D.p + (size_t)t[i + stride] * D.ld, n, lane);
                stage_commit();
                double base = __dsub_rn(-E[i - 1], E[i]);
                base = __dadd_rn(base, D(a, c));
                stage_wait<1>();                               // this scan row's matrix row has landed (the next may be in flight)
                __syncwarp();
                const double *rb = rbuf + cur * n;
                int j = 1 + lane;
                if (sym) {
                    for (; j + 96 <= n - 1; j += 128) {        // four candidates per lane in flight
                        double x[4], y[4], e[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int ju = j + 32 * u, q = (i < ju) ? ju : ju - 1;   // ju in {i, i-1}: valid indices, dropped below
                            x[u] = rb[t[q]]; y[u] = rb[t[q + 1]]; e[u] = E[q];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) x[u] = __dadd_rn(__dadd_rn(__dsub_rn(base, e[u]), x[u]), y[u]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int ju = j + 32 * u;
                            if (ju != i && ju != i - 1) consider(best, x[u], (i << 16) | ju, fi);
                        }
                    }
                }
                for (; j <= n - 1; j += 32) {
                    if (j == i || j == i - 1) continue;
                    const int q = (i < j) ? j : j - 1;
                    const int d = t[q], e = t[q + 1];
                    double x = __dsub_rn(base, E[q]);
                    x = __dadd_rn(x, sym ? rb[d] : D(d, b));
                    x = __dadd_rn(x, rb[e]);
                    consider(best, x, (i << 16) | j, fi);
                }
                __syncwarp();                                  // before this buffer is refilled two scan rows from now
                cur ^= 1;
            }
            stage_wait<0>();
            return team.reduce(best, fi, red);
        }
    }
    for (int p = threadIdx.x; p <= n - 1; p += blockDim.x) E[p] = D(t[p], t[p + 1]);
    __syncthreads();
    for (int i = 1 + team.first_row(warp); i <= n - 1; i += team.row_stride(nw)) {
        const int a = t[i - 1], b = t[i], c = t[i + 1];
        double base = __dsub_rn(-E[i - 1], E[i]);
        base = __dadd_rn(base, D(a, c));
        int j = 1 + lane;
        if constexpr (T::kDeep) {
            for (; j + 96 <= n - 1; j += 128) {
                double x[4], y[4];
                int q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ju = j + 32 * u;
                    q[u] = (i < ju) ? ju : ju - 1;           // ju in {i, i-1} is evaluated (valid indices) and dropped below
                    x[u] = D(t[q[u]], b);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) y[u] = D(b, t[q[u] + 1]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ju = j + 32 * u;
                    double z = __dsub_rn(base, E[q[u]]);
                    z = __dadd_rn(z, x[u]);
                    z = __dadd_rn(z, y[u]);
                    if (ju != i && ju != i - 1) consider(best, z, (i << 16) | ju, fi);
                }
            }
        }
        for (; j <= n - 1; j += 32) {
            if (j == i || j == i - 1) continue;   // permutations(.,2) has no i==j; operators.py:135 skips i-j==1
            const int q = (i < j) ? j : j - 1;
            const int d = t[q], e = t[q + 1];
            double x = __dsub_rn(base, E[q]);
            x = __dadd_rn(x, D(d, b));
            x = __dadd_rn(x, D(b, e));
            consider(best, x, (i << 16) | j, fi);
        }
    }
    return team.reduce(best, fi, red);
}

// ----------------------------------------------------------------------------------------------
// one-to-all scans (operators.py:53-73, :106-126): fixed i, threads over j
// ----------------------------------------------------------------------------------------------
template <class M>
__device__ Best scan_two_opt_o2a(const int *t, int n, const M &D, int i, bool fi, Best *red) {
    Best best;
    best.delta = 0.0; best.key = -1; best.pad = 0;
    for (int j = 1 + threadIdx.x; j <= n - 1; j += blockDim.x) {
        const int df = i - j;
        if (df < 2 && df > -2) continue;
        const int lo = (i < j) ? i : j, hi = (i < j) ? j : i;   // two_opt_cost swaps when j < i
        const int a = t[lo], b = t[lo - 1], c = t[hi], d = t[hi - 1];
        double x = __dadd_rn(D(a, c), D(b, d));
        x = __dsub_rn(x, D(a, b));
        x = __dsub_rn(x, D(c, d));
        consider(best, x, j, fi);
    }
    return block_reduce_best(best, fi, red);
}

template <class M>
__device__ Best scan_relocate_o2a(const int *t, int n, const M &D, int i, bool fi, Best *red) {
    Best best;
    best.delta = 0.0; best.key = -1; best.pad = 0;
    const int a = t[i - 1], b = t[i], c = t[i + 1];
    double base = __dsub_rn(-D(a, b), D(b, c));
    base = __dadd_rn(base, D(a, c));
    for (int j = 1 + threadIdx.x; j <= n - 1; j += blockDim.x) {
        if (j == i) continue;
        const int q = (i < j) ? j : j - 1;
        const int d = t[q], e = t[q + 1];
        double x = __dsub_rn(base, D(d, e));
        x = __dadd_rn(x, D(d, b));
        x = __dadd_rn(x, D(b, e));
        consider(best, x, j, fi);
    }
    return block_reduce_best(best, fi, red);
}

// ----------------------------------------------------------------------------------------------
// move application (operators.py:6-11, :76-80); block-wide, ends with the tour consistent
// ----------------------------------------------------------------------------------------------
__device__ void apply_two_opt(int *t, int i, int j) {
    if (j < i) { const int s = i; i = j; j = s; }
    const int half = (j - i) >> 1;                         // reverse positions i .. j-1
    for (int p = threadIdx.x; p < half; p += blockDim.x) {
        const int x = t[i + p], y = t[j - 1 - p];
        t[i + p] = y; t[j - 1 - p] = x;
    }
    __syncthreads();
}

__device__ void apply_relocate(int *t, int *tmp, int i, int j) {   // pop(i); insert(j, node)
    const int node = t[i];
    const int lo = (i < j) ? i : j, hi = (i < j) ? j : i;
    for (int p = lo + threadIdx.x; p <= hi; p += blockDim.x) {
        int v;
        if (p == j) v = node;
        else v = (i < j) ? t[p + 1] : t[p - 1];
        tmp[p] = v;
    }
    __syncthreads();
    for (int p = lo + threadIdx.x; p <= hi; p += blockDim.x) t[p] = tmp[p];
    __syncthreads();
}

__device__ __forceinline__ void apply_move(int op, int *t, int *tmp, int i, int j) {
    if (op == GNNGLS_OP_TWO_OPT) apply_two_opt(t, i, j);
    else apply_relocate(t, tmp, i, j);
}

// strictly sequential fp64 sum of E[0..n) (gnngls/__init__.py:17-21), one thread.  Loads are batched four at a time (they do not depend
// on the running sum); the remainder loop is kept rolled so that ptxas does not predicate the uniform-register moves it likes to use for a
// single-thread reduction (tests/test_abi_and_host.py::test_no_divergent_uniform_register_moves_in_sass).
__device__ __forceinline__ double sequential_sum(const double *E, int n) {
    double c = 0.0;
    int p = 0;
#pragma unroll 1
    for (; p + 4 <= n; p += 4) {
        const double e0 = E[p], e1 = E[p + 1], e2 = E[p + 2], e3 = E[p + 3];
        c = __dadd_rn(c, e0); c = __dadd_rn(c, e1); c = __dadd_rn(c, e2); c = __dadd_rn(c, e3);
    }
#pragma unroll 1
    for (; p < n; ++p) c = __dadd_rn(c, E[p]);
    return c;
}

// gnngls/__init__.py:17-21: c = 0; c += w for consecutive edges (strictly sequential fp64 sum).
// Edge weights are gathered in parallel into E, then thread 0 adds them in order.
template <class M>
__device__ double tour_cost_seq(const int *t, int n, const M &D, double *E, double *slot) {
    for (int p = threadIdx.x; p < n; p += blockDim.x) E[p] = D(t[p], t[p + 1]);
    __syncthreads();
    if (threadIdx.x == 0) *slot = sequential_sum(E, n);
    __syncthreads();
    return *slot;
}

// ----------------------------------------------------------------------------------------------
// shared-memory carve-up
// ----------------------------------------------------------------------------------------------
struct Smem {
    double *D;       // n*ld (only when staged)
    double *E;       // n+1 per-position edge terms
    double *slot;    // 4 doubles of block-shared scalars
    Best *red;       // 32 (two-barrier reductions) + 2 x 32 (one-barrier, double buffered)
    Best *xch;       // 2 x kMaxCluster winners exchanged between the CTAs of a cluster
    int *tour;       // n+1
    int *tmp;        // n+1
    int *best_tour;  // n+1 (GLS only)
    int *ivars;      // 8 block-shared ints
    uint16_t *pen;   // n*n (GLS, staged only)
};

__host__ __device__ inline int ld_for(int n) { return n | 1; }   // odd row stride: column reads hit all banks

__host__ __device__ inline size_t smem_layout(int n, bool stage_d, bool gls, bool stage_pen, Smem *s, unsigned char *base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    size_t oD = stage_d ? take(sizeof(double) * (size_t)n * ld_for(n)) : 0;
    size_t oE = take(sizeof(double) * (n + 1));
    size_t oS = take(sizeof(double) * 4);
    size_t oR = take(sizeof(Best) * 96);
    size_t oX = take(sizeof(Best) * 2 * kMaxCluster);
    size_t oT = take(sizeof(int) * (n + 1));
    size_t oM = take(sizeof(int) * (n + 1));
    size_t oB = gls ? take(sizeof(int) * (n + 1)) : 0;
    size_t oI = take(sizeof(int) * 8);
    size_t oP = (gls && stage_pen) ? take(sizeof(uint16_t) * (size_t)n * n) : 0;
    if (s) {
        s->D = stage_d ? reinterpret_cast<double *>(base + oD) : nullptr;
        s->E = reinterpret_cast<double *>(base + oE);
        s->slot = reinterpret_cast<double *>(base + oS);
        s->red = reinterpret_cast<Best *>(base + oR);
        s->xch = reinterpret_cast<Best *>(base + oX);
        s->tour = reinterpret_cast<int *>(base + oT);
        s->tmp = reinterpret_cast<int *>(base + oM);
        s->best_tour = gls ? reinterpret_cast<int *>(base + oB) : nullptr;
        s->ivars = reinterpret_cast<int *>(base + oI);
        s->pen = (gls && stage_pen) ? reinterpret_cast<uint16_t *>(base + oP) : nullptr;
    }
    return off;
}

__device__ void stage_matrix(double *dst, const double *src, int n) {
    const int ld = ld_for(n);
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int r = idx / n, c = idx - r * n;
        dst[r * ld + c] = src[idx];
    }
}

struct EventLog {
    double *ev;
    int max_ev;
    int count;   // maintained by thread 0 only
    __device__ __forceinline__ void push(double c) {
        if (ev && count < max_ev) ev[count] = c;
        ++count;
    }
};

// algorithms.py:111-132.  All threads call; tour in shared memory is updated in place; *cost is
// a block-shared slot updated by thread 0.  Returns nothing; counters are thread-0 registers.
template <class M, class T>
__device__ void local_search_dev(const Smem &s, int n, const M &D, bool fi, double *cost_slot,
                                 EventLog &log, long long *cnt /* thread-0 local [4] */, T &team) {
    bool improved = true;
    while (improved) {
        improved = false;
#pragma unroll 1
        for (int op = 0; op < 2; ++op) {
            Best b = (op == 0) ? sweep_two_opt_a2a(s.tour, n, D, s.E, fi, s.red, team)
                               : sweep_relocate_a2a(s.tour, n, D, s.E, fi, s.red, team);
            if (threadIdx.x == 0) cnt[op] += 1;
            if (b.key >= 0) {                                  // delta < 0 by construction
                improved = true;
                apply_move(op, s.tour, s.tmp, b.key >> 16, b.key & 0xffff);
                if (threadIdx.x == 0) {
                    *cost_slot = __dadd_rn(*cost_slot, b.delta);   // cur_cost += delta (:124)
                    log.push(*cost_slot);
                    cnt[3] += 1;
                }
            }
            GLS_STAMP(5);                                      // move applied, cost and log updated
        }
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------------------------
// kernels
// ----------------------------------------------------------------------------------------------
// Team of the launch: CL == false -> the CTA; CL == true -> the cluster the CTA belongs to (instances are then dealt
// to clusters, every member keeps a tour replica and member 0 writes the results).
template <bool CL> struct TeamOf { using type = Solo; };
template <> struct TeamOf<true> { using type = Cluster; };

template <bool CL>
__device__ __forceinline__ typename TeamOf<CL>::type make_team(const Smem &s, int *inst0, int *inst_stride, double *rows = nullptr) {
    typename TeamOf<CL>::type team;
    if constexpr (CL) {
        cg::cluster_group cl = cg::this_cluster();
        team.rank = (int)cl.block_rank();
        team.size = (int)cl.num_blocks();
        team.parity = 0;
        team.rp = 0;
        team.xch = s.xch;
        team.rows = rows;
        team.symmetric = 0;
        *inst0 = (int)blockIdx.x / team.size;
        *inst_stride = (int)gridDim.x / team.size;
        cl.sync();             // distributed shared memory of a member may only be touched once that member has started
                               // (compute-sanitizer: "located in a block that might not have entered yet")
    } else {
        *inst0 = (int)blockIdx.x;
        *inst_stride = (int)gridDim.x;
    }
    return team;
}

// D[a][b] == D[b][a] bit for bit?  Checked once per instance by the whole cluster (each member takes a share of the
// pairs; the verdict travels through the winners' exchange: a candidate with key >= 0 means "found an asymmetric pair").
__device__ void cluster_check_symmetric(const double *Dg, int n, Cluster &team, Best *red) {
    if (!team.rows) return;
    const long long *Db = reinterpret_cast<const long long *>(Dg);
    int ok = 1;
    for (int r = team.rank; r < n; r += team.size)
        for (int c = r + 1 + threadIdx.x; c < n; c += blockDim.x) ok &= Db[(size_t)r * n + c] == Db[(size_t)c * n + r];
    Best v;
    v.delta = 0.0; v.key = ok ? -1 : 0; v.pad = 0;
    team.symmetric = team.reduce(v, false, red).key < 0;
}

template <bool STAGE_D, bool CL>
__device__ __forceinline__ void moves_body(int op, bool o2a, const double *Dg, int64_t d_stride, const int *tours,
                             const int *pos, int B, int n, int fi, double *out_delta, int *out_move,
                             int *out_tours, int row_cache = 0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s;
    const size_t used = smem_layout(n, STAGE_D, false, false, &s, smem_raw);
    int inst0, inst_stride;
    // (a single sweep does not pay for a symmetry check: relocate's column term stays a global gather)
    auto team = make_team<CL>(s, &inst0, &inst_stride, row_cache ? reinterpret_cast<double *>(smem_raw + used) : nullptr);
    const bool writer = team.lead();
    for (int b = inst0; b < B; b += inst_stride) {
        const double *Db = Dg + (size_t)b * d_stride;
        for (int p = threadIdx.x; p <= n; p += blockDim.x) s.tour[p] = tours[(size_t)b * (n + 1) + p];
        MatPlain D;
        if (STAGE_D) {
            if (b == inst0 || d_stride != 0) stage_matrix(s.D, Db, n);
            D.p = s.D; D.ld = ld_for(n);
        } else {
            D.p = Db; D.ld = n;
        }
        __syncthreads();
        Best best;
        int i_fixed = 0;
        if (o2a) {
            i_fixed = pos[b];
            // the reference asserts 0 < i < len(tour) - 1 (operators.py:107,129); a position outside that range reads
            // t[i-1] / t[i+1] out of bounds, so it is reported as "no move" with position -2 instead
            if (i_fixed < 1 || i_fixed > n - 1) {
                best.delta = 0.0; best.key = -1; best.pad = 0;
                if (threadIdx.x == 0) { out_delta[b] = 0.0; out_move[2 * b] = -2; out_move[2 * b + 1] = -2; }
                if (out_tours)
                    for (int p = threadIdx.x; p <= n; p += blockDim.x) out_tours[(size_t)b * (n + 1) + p] = s.tour[p];
                __syncthreads();
                continue;
            }
            best = (op == GNNGLS_OP_TWO_OPT) ? scan_two_opt_o2a(s.tour, n, D, i_fixed, fi != 0, s.red)
                                             : scan_relocate_o2a(s.tour, n, D, i_fixed, fi != 0, s.red);
        } else {
            best = (op == GNNGLS_OP_TWO_OPT) ? sweep_two_opt_a2a(s.tour, n, D, s.E, fi != 0, s.red, team)
                                             : sweep_relocate_a2a(s.tour, n, D, s.E, fi != 0, s.red, team);
        }
        const bool found = best.key >= 0;
        const int mi = !found ? -1 : (o2a ? i_fixed : (best.key >> 16));
        const int mj = !found ? -1 : (o2a ? best.key : (best.key & 0xffff));
        if (threadIdx.x == 0 && writer) {
            out_delta[b] = found ? best.delta : 0.0;
            out_move[2 * b] = mi; out_move[2 * b + 1] = mj;
        }
        if (out_tours && writer) {
            if (found) apply_move(op, s.tour, s.tmp, mi, mj);
            for (int p = threadIdx.x; p <= n; p += blockDim.x) out_tours[(size_t)b * (n + 1) + p] = s.tour[p];
        }
        __syncthreads();
    }
}

template <bool STAGE_D, bool CL>
__device__ __forceinline__ void local_search_body(const double *Dg, int *tours, double *costs, int B, int n, int fi,
                                    double *events, int *n_events, int max_events, int *status,
                                    long long *counters, int row_cache = 0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s;
    const size_t used = smem_layout(n, STAGE_D, false, false, &s, smem_raw);
    int inst0, inst_stride;
    auto team = make_team<CL>(s, &inst0, &inst_stride, row_cache ? reinterpret_cast<double *>(smem_raw + used) : nullptr);
    const bool writer = team.lead();
    for (int b = inst0; b < B; b += inst_stride) {
        const double *Db = Dg + (size_t)b * n * n;
        for (int p = threadIdx.x; p <= n; p += blockDim.x) s.tour[p] = tours[(size_t)b * (n + 1) + p];
        MatPlain D;
        if (STAGE_D) { stage_matrix(s.D, Db, n); D.p = s.D; D.ld = ld_for(n); }
        else { D.p = Db; D.ld = n; }
        if (threadIdx.x == 0) s.slot[0] = costs[b];
        __syncthreads();
        if constexpr (CL) cluster_check_symmetric(Db, n, team, s.red);
#ifdef GLS_STAMPS
        if (threadIdx.x == 0) { long long *sb = gls_stamp_buf(); for (int q = 0; q < 8; ++q) sb[q] = 0; sb[8] = clock64(); }
#endif
        EventLog log{(events && writer) ? events + (size_t)b * max_events : nullptr, max_events, 0};
        long long cnt[4] = {0, 0, 0, 0};
        local_search_dev(s, n, D, fi != 0, &s.slot[0], log, cnt, team);
#ifdef GLS_STAMPS
        if constexpr (CL) if (threadIdx.x == 0 && b == inst0) for (int q = 0; q < 8; ++q) g_gls_stamps[(blockIdx.x & 15) * 8 + q] = gls_stamp_buf()[q];
#endif
        // every member has read the instance's tour before the first barrier of the search; member 0 writes it back
        if (writer) for (int p = threadIdx.x; p <= n; p += blockDim.x) tours[(size_t)b * (n + 1) + p] = s.tour[p];
        if (threadIdx.x == 0 && writer) {
            costs[b] = s.slot[0];
            if (n_events) n_events[b] = log.count;
            if (status) status[b] = (events && log.count > max_events) ? GNNGLS_INST_EVENTS_TRUNCATED : 0;
            if (counters) for (int q = 0; q < 4; ++q) counters[4 * (size_t)b + q] = cnt[q];
        }
        __syncthreads();
    }
}

// guide value of the undirected edge (u,v)
struct GuideRef {
    const double *mat;   // [n,n] or null
    const float *vec;    // [N]  or null
    int n;
    __device__ __forceinline__ double operator()(int u, int v) const {
        if (mat) return mat[u * n + v];
        const int i = u < v ? u : v, j = u < v ? v : u;
        return (double)vec[i * (2 * n - i - 1) / 2 + (j - i - 1)];
    }
};

struct GlsDev {
    gnngls_gls_args a;
    int row_cache;   // cluster tier: per-warp row cache behind the regular shared-memory layout
};

constexpr int kStallCap = 1 << 16;   // safety cap on perturbation-loop trips per outer iteration

// algorithms.py:135-195.  STAGED: D (fp64) and the penalties (u16) live in shared memory;
// otherwise both are read from global memory (L2-resident).
// CL (never with STAGED): one thread-block CLUSTER per instance.  The all-pairs sweeps of every local_search -- all of the
// O(n^2) work -- are shared by the cluster's CTAs (struct Cluster); the perturbation loop, whose scans are O(n) and which
// owns the penalties, runs on member 0 alone while the others wait at the cluster barrier and then copy member 0's tour
// and cost out of its shared memory.  Every member applies the same moves to its replica, so all of them track the
// same current and best tour; member 0 writes the state back.
template <bool STAGED, bool CL>
__device__ __forceinline__ void gls_body(const GlsDev &P) {
    static_assert(!(STAGED && CL), "the cluster tier reads D and the penalties from global memory");
    const gnngls_gls_args &a = P.a;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n;
    Smem s;
    const size_t used = smem_layout(n, STAGED, true, STAGED, &s, smem_raw);
    const bool fi = a.first_improvement != 0;
    const size_t nn = (size_t)n * n;
    if (threadIdx.x == 0) s.ivars[0] = 0;
    int inst0, inst_stride;
    auto team = make_team<CL>(s, &inst0, &inst_stride, P.row_cache ? reinterpret_cast<double *>(smem_raw + used) : nullptr);
    const bool lead = team.lead();

    for (int b = inst0; b < a.B; b += inst_stride) {
        const double *Db = a.D + (size_t)b * nn;
        int *pen_g = a.penalties ? a.penalties + (size_t)b * nn : nullptr;
        int status = 0;
        // ---- load state
        for (int p = threadIdx.x; p <= n; p += blockDim.x) s.tour[p] = a.cur_tours[(size_t)b * (n + 1) + p];
        MatPlain D;
        if (STAGED) {
            stage_matrix(s.D, Db, n);
            D.p = s.D; D.ld = ld_for(n);
            for (int idx = threadIdx.x; idx < (int)nn; idx += blockDim.x) {
                int v = (a.resume && pen_g) ? pen_g[idx] : 0;
                if (v > 65535) { v = 65535; status |= GNNGLS_INST_PENALTY_OVERFLOW; }
                s.pen[idx] = (uint16_t)v;
            }
        } else {
            D.p = Db; D.ld = n;
            // only member 0 ever touches the penalties, so it alone clears them
            if (!a.resume && lead) for (int idx = threadIdx.x; idx < (int)nn; idx += blockDim.x) pen_g[idx] = 0;
        }
        // slot[0] = cur_cost, slot[1] = best_cost, slot[2] = scratch for tour_cost
        if (threadIdx.x == 0) {
            s.slot[0] = a.cur_costs[b];
            if (a.resume) s.slot[1] = a.best_costs[b];
        }
        if (a.resume) for (int p = threadIdx.x; p <= n; p += blockDim.x) s.best_tour[p] = a.best_tours[(size_t)b * (n + 1) + p];
        __syncthreads();
        if constexpr (CL) cluster_check_symmetric(Db, n, team, s.red);
#ifdef GLS_STAMPS
        if (threadIdx.x == 0 && b == inst0) { long long *sb = gls_stamp_buf(); for (int q = 0; q < 8; ++q) sb[q] = 0; sb[8] = clock64(); }
#endif
        double k;
        if (a.resume) k = a.k[b];
        else k = __ddiv_rn(__dmul_rn(0.1, s.slot[0]), (double)n);                // :137
        EventLog log{(a.events && lead) ? a.events + (size_t)b * a.max_events : nullptr, a.max_events, 0};
        long long cnt[4] = {0, 0, 0, 0};

        if (!a.resume) {
            local_search_dev(s, n, D, fi, &s.slot[0], log, cnt, team);            // :142
            if (b == inst0) GLS_STAMP(7);                                         // (one-CTA tier: everything of the local search)
            for (int p = threadIdx.x; p <= n; p += blockDim.x) s.best_tour[p] = s.tour[p];   // :143
            if (threadIdx.x == 0) s.slot[1] = s.slot[0];
            __syncthreads();
        }

        for (int it = a.iter_begin; it < a.iter_begin + a.n_iters; ++it) {        // :146 (explicit range)
            const int gsel = it % a.n_guides;                                     // :147
            GuideRef guide;
            guide.n = n;
            if (a.guide_kind == GNNGLS_GUIDE_MATRIX_F64) {
                guide.mat = static_cast<const double *>(a.guides) + ((size_t)b * a.n_guides + gsel) * nn;
                guide.vec = nullptr;
            } else {
                guide.mat = nullptr;
                guide.vec = static_cast<const float *>(a.guides) + ((size_t)b * a.n_guides + gsel) * (nn - n) / 2;
            }
            int moves = 0, trips = 0;
            while (lead && moves < a.perturbation_moves) {                        // :151
                if (++trips > kStallCap) { status |= GNNGLS_INST_STALLED; break; }
                // ---- :153-159 arg-max utility over tour edges, first edge wins ties
                Best u;
                u.delta = 0.0; u.key = -1; u.pad = 0;
                for (int p = threadIdx.x; p < n; p += blockDim.x) {
                    const int x = s.tour[p], y = s.tour[p + 1];
                    const double pen = STAGED ? (double)s.pen[x * n + y] : (double)pen_g[x * n + y];
                    const double util = __ddiv_rn(guide(x, y), __dadd_rn(1.0, pen));
                    // reuse the min-reduction on the negated utility: min(-util), ties -> smaller p
                    const double neg = (util != util) ? INFINITY : -util;       // NaN guide: never the arg-max
                    if (u.key < 0 || neg < u.delta) { u.delta = neg; u.key = p; }
                }
                u = block_reduce_best<true>(u, false, s.red);
                const int pe = u.key;
                const int eu = s.tour[pe], ev = s.tour[pe + 1];
                __syncthreads();
                if (threadIdx.x == 0) {                                           // :161
                    if (STAGED) {
                        unsigned v = (unsigned)s.pen[eu * n + ev] + 1u;
                        if (v > 65535u) { v = 65535u; s.ivars[0] = 1; }
                        s.pen[eu * n + ev] = (uint16_t)v; s.pen[ev * n + eu] = (uint16_t)v;
                    } else {
                        const int v = pen_g[eu * n + ev] + 1;
                        pen_g[eu * n + ev] = v; pen_g[ev * n + eu] = v;
                    }
                }
                __syncthreads();
                // ---- :163-164 penalised matrix, evaluated on read
#define GLS_PERTURB(MATPEN)                                                                       \
    for (int e2 = 0; e2 < 2; ++e2) {                                          /* :167 */          \
        const int node = e2 == 0 ? eu : ev;                                                       \
        if (node == 0) continue;                                              /* :168 */          \
        for (int p = 1 + threadIdx.x; p <= n - 1; p += blockDim.x)            /* :169 */          \
            if (s.tour[p] == node) s.ivars[1] = p;                                                \
        __syncthreads();                                                                          \
        const int i = s.ivars[1];                                                                 \
        for (int op = 0; op < 2; ++op) {                                      /* :171 */          \
            Best m = (op == 0) ? scan_two_opt_o2a(s.tour, n, MATPEN, i, fi, s.red)                \
                               : scan_relocate_o2a(s.tour, n, MATPEN, i, fi, s.red);              \
            if (threadIdx.x == 0) cnt[2] += 1;                                                    \
            if (m.key >= 0) {                                                 /* :175-185 */      \
                apply_move(op, s.tour, s.tmp, i, m.key);                                          \
                const double c = tour_cost_seq(s.tour, n, D, s.E, &s.slot[2]);                    \
                if (threadIdx.x == 0) { s.slot[0] = c; log.push(c); cnt[3] += 1; }                \
                moves += 1;                                                                       \
            }                                                                                     \
        }                                                                                         \
    }
                if (STAGED) {
                    MatPen<uint16_t> Dp{s.D, s.pen, ld_for(n), n, k};
                    GLS_PERTURB(Dp)
                } else {
                    MatPen<int> Dp{Db, pen_g, n, n, k};
                    GLS_PERTURB(Dp)
                }
#undef GLS_PERTURB
            }
            if (b == inst0) GLS_STAMP(6);                                         // perturbation
            if constexpr (CL) {
                // hand member 0's perturbed tour and its cost to the other members.  Member 0 changes neither before
                // the first barrier of the local search below, which no member reaches before it has finished copying.
                cg::cluster_group cl = cg::this_cluster();
                cl.sync();
                if (!lead) {
                    const int *t0 = cl.map_shared_rank(s.tour, 0);
                    for (int p = threadIdx.x; p <= n; p += blockDim.x) s.tour[p] = t0[p];
                    if (threadIdx.x == 0) s.slot[0] = *cl.map_shared_rank(&s.slot[0], 0);
                }
                __syncthreads();
            }
            local_search_dev(s, n, D, fi, &s.slot[0], log, cnt, team);            // :188
            if (b == inst0) GLS_STAMP(7);
            if (s.slot[0] < s.slot[1]) {                                          // :190-191 (block-uniform)
                for (int p = threadIdx.x; p <= n; p += blockDim.x) s.best_tour[p] = s.tour[p];
                __syncthreads();
                if (threadIdx.x == 0) s.slot[1] = s.slot[0];
            }
            __syncthreads();
        }

#ifdef GLS_STAMPS
        if (threadIdx.x == 0 && b == inst0 && blockIdx.x < 16) for (int q = 0; q < 8; ++q) g_gls_stamps[blockIdx.x * 8 + q] = gls_stamp_buf()[q];
#endif
        // ---- store state (member 0 of a cluster; every member holds the same tours and costs)
        if (lead) {
            for (int p = threadIdx.x; p <= n; p += blockDim.x) {
                a.cur_tours[(size_t)b * (n + 1) + p] = s.tour[p];
                a.best_tours[(size_t)b * (n + 1) + p] = s.best_tour[p];
            }
            if (STAGED && pen_g) for (int idx = threadIdx.x; idx < (int)nn; idx += blockDim.x) pen_g[idx] = s.pen[idx];
        }
        if (threadIdx.x == 0 && lead) {
            a.cur_costs[b] = s.slot[0];
            a.best_costs[b] = s.slot[1];
            if (!a.resume) a.k[b] = k;
            if (a.n_events) a.n_events[b] = log.count;
            if (STAGED && s.ivars[0]) status |= GNNGLS_INST_PENALTY_OVERFLOW;
            if (a.events && log.count > a.max_events) status |= GNNGLS_INST_EVENTS_TRUNCATED;
            if (a.status) a.status[b] = status;
            if (a.counters) for (int q = 0; q < 4; ++q) a.counters[4 * (size_t)b + q] = cnt[q];
            s.ivars[0] = 0;
        }
        __syncthreads();
    }
    // no member may exit while another can still read its shared memory
    if constexpr (CL) cg::this_cluster().sync();
}

// Kernel entry points.  The one-CTA tiers keep the compiler's own register choice (48 registers for the L2-resident GLS
// tier: five CTAs of 256 threads per SM); the cluster tier is bounded for CTAs of up to 1024 threads.
template <bool STAGE_D>
__global__ void moves_kernel(int op, bool o2a, const double *Dg, int64_t d_stride, const int *tours, const int *pos, int B,
                             int n, int fi, double *out_delta, int *out_move, int *out_tours) {
    moves_body<STAGE_D, false>(op, o2a, Dg, d_stride, tours, pos, B, n, fi, out_delta, out_move, out_tours);
}
__global__ void __launch_bounds__(1024) moves_cluster_kernel(int op, bool o2a, const double *Dg, int64_t d_stride,
                                                             const int *tours, const int *pos, int B, int n, int fi,
                                                             double *out_delta, int *out_move, int *out_tours,
                                                             int row_cache) {
    moves_body<false, true>(op, o2a, Dg, d_stride, tours, pos, B, n, fi, out_delta, out_move, out_tours, row_cache);
}
template <bool STAGE_D>
__global__ void local_search_kernel(const double *Dg, int *tours, double *costs, int B, int n, int fi, double *events,
                                    int *n_events, int max_events, int *status, long long *counters) {
    local_search_body<STAGE_D, false>(Dg, tours, costs, B, n, fi, events, n_events, max_events, status, counters);
}
__global__ void __launch_bounds__(1024) local_search_cluster_kernel(const double *Dg, int *tours, double *costs, int B,
                                                                    int n, int fi, double *events, int *n_events,
                                                                    int max_events, int *status, long long *counters,
                                                                    int row_cache) {
    local_search_body<false, true>(Dg, tours, costs, B, n, fi, events, n_events, max_events, status, counters, row_cache);
}
template <bool STAGED>
__global__ void gls_kernel(const GlsDev P) { gls_body<STAGED, false>(P); }
__global__ void __launch_bounds__(1024) gls_cluster_kernel(const GlsDev P) { gls_body<false, true>(P); }

// algorithms.py:9-18 + __init__.py:17-21.  One warp per instance.
__global__ void nn_init_kernel(int guide_kind, const void *guides, const double *Dg, int B, int n, int depot,
                               int *out_tours, double *out_costs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double *E = reinterpret_cast<double *>(smem_raw) + (size_t)warp * (n + 1);
    int *tour = reinterpret_cast<int *>(smem_raw + sizeof(double) * (size_t)wpb * (n + 1)) + (size_t)warp * (n + 1);
    const size_t nn = (size_t)n * n;
    for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
        GuideRef g;
        g.n = n;
        g.mat = guide_kind == GNNGLS_GUIDE_MATRIX_F64 ? static_cast<const double *>(guides) + (size_t)b * nn : nullptr;
        g.vec = guide_kind == GNNGLS_GUIDE_MATRIX_F64 ? nullptr : static_cast<const float *>(guides) + (size_t)b * (nn - n) / 2;
        unsigned visited = 0;   // bit q <-> node lane + 32*q  (n <= 1024)
        if ((depot & 31) == lane) visited |= 1u << (depot >> 5);
        int cur = depot;
        if (lane == 0) tour[0] = depot;
        for (int len = 1; len < n; ++len) {
            double bw = 0.0;
            int bj = -1;
            for (int q = 0, j = lane; j < n; ++q, j += 32) {
                if (visited & (1u << q)) continue;
                double w = g(cur, j);
                if (w != w) w = INFINITY;                      // NaN guide values: keep the comparison a total order
                if (bj < 0 || w < bw) { bw = w; bj = j; }      // ascending j per lane: first minimum kept
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ow = __shfl_xor_sync(0xffffffffu, bw, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (oj >= 0 && (bj < 0 || ow < bw || (ow == bw && oj < bj))) { bw = ow; bj = oj; }
            }
            if ((bj & 31) == lane) visited |= 1u << (bj >> 5);
            if (lane == 0) tour[len] = bj;
            cur = bj;
        }
        if (lane == 0) tour[n] = depot;
        __syncwarp();
        for (int p = lane; p <= n; p += 32) out_tours[(size_t)b * (n + 1) + p] = tour[p];
        if (Dg && out_costs) {
            const double *Db = Dg + (size_t)b * nn;
            for (int p = lane; p < n; p += 32) E[p] = Db[(size_t)tour[p] * n + tour[p + 1]];
            __syncwarp();
            if (lane == 0) {
                double c = 0.0;
                for (int p = 0; p < n; ++p) c = __dadd_rn(c, E[p]);
                out_costs[b] = c;
            }
        }
        __syncwarp();
    }
}

// The same constructor with one CTA per instance, one candidate node per thread (n <= 1024): for a handful of large instances
// the n-1 dependent steps are the whole cost, and a step is then one guide load + one (value, node) reduction with a single
// block barrier (the warps' partial minima are double buffered by step parity and combined redundantly by every warp).
__global__ void __launch_bounds__(1024) nn_init_block_kernel(int guide_kind, const void *guides, const double *Dg, int B,
                                                             int n, int depot, int *out_tours, double *out_costs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *E = reinterpret_cast<double *>(smem_raw);                        // n + 1
    double *pw = E + (n + 1);                                                // 2 x 32 partial minima
    int *pj = reinterpret_cast<int *>(pw + 64);                              // 2 x 32 their nodes
    int *tour = pj + 64;                                                     // n + 1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int j = threadIdx.x;
    const size_t nn = (size_t)n * n;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        GuideRef g;
        g.n = n;
        g.mat = guide_kind == GNNGLS_GUIDE_MATRIX_F64 ? static_cast<const double *>(guides) + (size_t)b * nn : nullptr;
        g.vec = guide_kind == GNNGLS_GUIDE_MATRIX_F64 ? nullptr : static_cast<const float *>(guides) + (size_t)b * (nn - n) / 2;
        bool visited = (j >= n) || (j == depot);
        int cur = depot;
        if (threadIdx.x == 0) { tour[0] = depot; tour[n] = depot; }
        for (int len = 1; len < n; ++len) {
            double bw = 0.0;
            int bj = -1;
            if (!visited) {
                bw = g(cur, j);
                if (bw != bw) bw = INFINITY;                                  // NaN guide values: keep the comparison a total order
                bj = j;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ow = __shfl_xor_sync(0xffffffffu, bw, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (oj >= 0 && (bj < 0 || ow < bw || (ow == bw && oj < bj))) { bw = ow; bj = oj; }
            }
            const int par = (len & 1) * 32;
            if (lane == 0) { pw[par + warp] = bw; pj[par + warp] = bj; }
            __syncthreads();
            bw = 0.0; bj = -1;
            if (lane < nw) { bw = pw[par + lane]; bj = pj[par + lane]; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ow = __shfl_xor_sync(0xffffffffu, bw, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (oj >= 0 && (bj < 0 || ow < bw || (ow == bw && oj < bj))) { bw = ow; bj = oj; }
            }
            if (bj == j) visited = true;
            if (threadIdx.x == 0) tour[len] = bj;
            cur = bj;
        }
        __syncthreads();
        for (int p = threadIdx.x; p <= n; p += blockDim.x) out_tours[(size_t)b * (n + 1) + p] = tour[p];
        if (Dg && out_costs) {
            const double *Db = Dg + (size_t)b * nn;
            for (int p = threadIdx.x; p < n; p += blockDim.x) E[p] = Db[(size_t)tour[p] * n + tour[p + 1]];
            __syncthreads();
            if (threadIdx.x == 0) out_costs[b] = sequential_sum(E, n);
        }
        __syncthreads();
    }
}

__global__ void tour_cost_kernel(const double *Dg, const int *tours, int B, int n, double *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double *E = reinterpret_cast<double *>(smem_raw) + (size_t)warp * (n + 1);
    for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
        const double *Db = Dg + (size_t)b * n * n;
        const int *t = tours + (size_t)b * (n + 1);
        for (int p = lane; p < n; p += 32) E[p] = Db[(size_t)t[p] * n + t[p + 1]];
        __syncwarp();
        if (lane == 0) {
            double c = 0.0;
            for (int p = 0; p < n; ++p) c = __dadd_rn(c, E[p]);
            out[b] = c;
        }
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------
// launch helpers
// ----------------------------------------------------------------------------------------------
int pick_threads(int n) {
    if (const char *e = getenv("GNNGLS_SEARCH_THREADS")) { const int t = atoi(e); if (t >= 32 && t <= 1024 && t % 32 == 0) return t; }   // A/B knob
    if (n <= 40) return 64;
    if (n <= 72) return 128;
    if (n <= 160) return 192;      // TSP100: 7 CTAs of 6 warps per SM; same-box A/B of the GLS stage: 128 / 192 / 256 / 384 threads = 63.8 / 59.3 / 63.0 / 67.9 ms

    if (n <= 400) return 512;
    return 1024;
}

template <typename K>
int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        GNNGLS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return GNNGLS_OK;
}

int grid_for(int B) {
    const int cap = gnngls::device_sm_count() * 16;
    return B < cap ? (B > 0 ? B : 1) : cap;
}

// Cluster tier (few, large instances).  One CTA per instance leaves most of the GPU idle when B is far below the SM
// count -- TSP500 x 8 used 8 of 148 SMs -- so the instance's sweeps are shared by a thread-block cluster instead.
// cluster_limit_for() gives the largest size worth trying: the largest power of two <= min(16, 2 SMs / B) (CTAs of the
// cluster tier have at most 512 threads, two fit an SM), when that is at least 2 and n is large enough for a sweep to
// outlast a cluster barrier (~0.2 us); launch_clustered() then picks, among the sizes up to that limit, the one that
// finishes the batch in the fewest (rounds of co-resident clusters) / (cluster size).
// GNNGLS_CLUSTER=0 disables the tier, GNNGLS_CLUSTER=2|4|8|16 forces a size (tests, A/B).
constexpr int kClusterMinN = 48;
constexpr int kClusterMaxThreads = 512;

int cluster_threads(int n) { const int t = pick_threads(n); return t < kClusterMaxThreads ? t : kClusterMaxThreads; }

// > 0: upper limit chosen by the policy; < 0: -(forced size); 1: one CTA per instance
int cluster_limit_for(int B, int n) {
    int forced = -1;
    if (const char *e = getenv("GNNGLS_CLUSTER")) forced = atoi(e);
    if (forced == 0 || forced == 1) return 1;
    if (forced > 1) {
        int c = 2;
        while (c * 2 <= forced && c * 2 <= kMaxCluster) c *= 2;
        return -c;
    }
    if (n < kClusterMinN) return 1;
    const int sms = gnngls::device_sm_count();
    int c = 1;
    while (c * 2 <= kMaxCluster && (long long)B * (c * 2) <= 2LL * sms) c *= 2;
    // no more members than there are rows to deal
    while (c > 1 && c * (cluster_threads(n) / 32) > 2 * n) c /= 2;
    return c;
}

// how many clusters of `csize` CTAs of this kernel can be resident at once (0: cannot be placed); cached
int max_active_clusters(const void *kernel, int csize, int threads, size_t smem) {
    struct Key { const void *k; int c, t; size_t s; int dev; int val; };
    static std::mutex mu;
    static std::vector<Key> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> g(mu);
        for (const Key &e : cache)
            if (e.k == kernel && e.c == csize && e.t == threads && e.s == smem && e.dev == dev) return e.val;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.gridDim = dim3(csize);
    int mc = 0;
    if (cudaOccupancyMaxActiveClusters(&mc, kernel, &cfg) != cudaSuccess) { (void)cudaGetLastError(); mc = 0; }
    std::lock_guard<std::mutex> g(mu);
    cache.push_back(Key{kernel, csize, threads, smem, dev, mc});
    return mc;
}

// Shape of a cluster-tier launch.  `limit` comes from cluster_limit_for().  Among the cluster sizes up to the limit and the CTA
// sizes on offer the plan takes the combination with the smallest (rounds of co-resident clusters) / (threads per instance).
// With the row cache (n >= kRowCacheMinN, or GNNGLS_ROWCACHE=1; =0 disables) a CTA has as many warps as 16 n bytes of shared memory
// each allow; without it the CTA sizes are the one-CTA tier's and 512 threads (two CTAs per SM).
constexpr int kRowCacheMinN = 192;

struct ClusterPlan {
    int csize = 0, threads = 0, clusters = 0, row_cache = 0;
    size_t smem = 0;
};

int plan_cluster(const void *kernel, int limit, int B, int n, size_t base_smem, const char *what, ClusterPlan *plan,
                 bool rows_allowed = true) {
    int want_rows = rows_allowed && n >= kRowCacheMinN;
    if (const char *e = getenv("GNNGLS_ROWCACHE")) want_rows = rows_allowed && atoi(e) != 0;
    const size_t row_bytes = 2 * sizeof(double) * (size_t)n;          // per warp
    const size_t room = (size_t)gnngls::device_max_optin_smem() - 1024;
    struct Option { int threads, rows; } options[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    options[0].threads = pick_threads(n);
    options[1].threads = cluster_threads(n) < options[0].threads ? cluster_threads(n) : 0;
    if (want_rows && base_smem + 4 * row_bytes <= room) {
        int warps = (int)((room - base_smem) / row_bytes);
        if (warps > pick_threads(n) / 32) warps = pick_threads(n) / 32;
        options[2] = Option{32 * warps, 1};
        options[3] = Option{warps >= 8 ? 32 * (warps / 2) : 0, 1};
        if (getenv("GNNGLS_ROWCACHE")) options[0].threads = options[1].threads = 0;     // forced on: only the row-cache shapes
    }
    double best_score = 0.0, best_gain = 1.0;
    for (const Option &o : options) {
        if (!o.threads) continue;
        const size_t smem = base_smem + (o.rows ? (size_t)(o.threads / 32) * row_bytes : 0);
        for (int c = limit < 0 ? -limit : limit; c >= 2; c /= 2) {
            const int mc = max_active_clusters(kernel, c, o.threads, smem);
            if (mc < 1) continue;
            const int rounds = (B + mc - 1) / mc;
            // what a thread with the row cache gets done against one without grows with the row length (n = 500: 864 threads 7.0 ms
            // against 1024 threads 8.1 ms = 1.4 x per thread; n = 1000: 416 threads 35 ms against 1024 threads 52 ms = 3.6 x)
            const double gain = o.rows ? (n > 300 ? n / 300.0 : 1.0) : 1.0;
            const double score = (double)rounds / ((double)c * o.threads * gain);
            // ties go to the smaller CTA: two of them share an SM's L1 and registers more evenly than one large one
            if (!plan->csize || score <= best_score) {
                plan->csize = c; plan->threads = o.threads; plan->clusters = B < mc ? B : mc; plan->smem = smem;
                plan->row_cache = o.rows;
                best_score = score;
                best_gain = gain;
            }
            if (limit < 0) break;                      // forced: that size, or the next smaller one that can be placed
        }
    }
    if (plan->csize < 2) {                             // no cluster shape can be placed on this device: one CTA per instance
        GNNGLS_REQUIRE(limit > 0, GNNGLS_ERR_CUDA, "launch of %s failed: the forced cluster size cannot be placed", what);
        plan->csize = 1;
        return GNNGLS_OK;
    }
    // A cluster must bring clearly more threads to an instance than the one-CTA tier would (which runs every eligible batch in one
    // round): with the row cache's small CTAs and a batch that needs two rounds of clusters it does not (n = 1000 x 64: 2 x 416
    // against 1024 threads, measured slower) -- csize = 1 tells the caller to launch the one-CTA tier.
    if (limit > 0 && 2.0 * plan->csize * plan->threads * plan->clusters * best_gain < 3.0 * pick_threads(n) * B) plan->csize = 1;
    return GNNGLS_OK;
}

template <typename K>
int prepare_cluster_kernel(K kernel) {
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaFuncAttributes fa;
    GNNGLS_CUDA_OK(cudaFuncGetAttributes(&fa, kernel));
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        gnngls::device_max_optin_smem() - (int)fa.sharedSizeBytes));
    return GNNGLS_OK;
}

template <typename... KArgs, typename... Args>
int launch_planned(void (*kernel)(KArgs...), const ClusterPlan &plan, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(plan.threads); cfg.dynamicSmemBytes = plan.smem; cfg.stream = st;
    cfg.gridDim = dim3(plan.clusters * plan.csize);
    GNNGLS_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
    return GNNGLS_OK;
}

int check_n(int n) {
    GNNGLS_REQUIRE(n >= 3 && n <= 1024, GNNGLS_ERR_UNSUPPORTED, "n=%d outside supported range [3,1024]", n);
    return GNNGLS_OK;
}

int launch_moves(int op, bool o2a, const double *D, int64_t stride, const int *tours, const int *pos, int B, int n,
                 int fi, double *out_delta, int *out_move, int *out_tours, cudaStream_t st) {
    GNNGLS_REQUIRE(op == GNNGLS_OP_TWO_OPT || op == GNNGLS_OP_RELOCATE, GNNGLS_ERR_BAD_ARG, "bad move op %d", op);
    GNNGLS_REQUIRE(D && tours && out_delta && out_move && (!o2a || pos), GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(stride == 0 || stride == (int64_t)n * n, GNNGLS_ERR_BAD_ARG, "d_batch_stride must be 0 or n*n");
    if (int rc = check_n(n)) return rc;
    if (B <= 0) return GNNGLS_OK;
    const int threads = pick_threads(n);
    const size_t staged = smem_layout(n, true, false, false, nullptr, nullptr);
    const size_t plain = smem_layout(n, false, false, false, nullptr, nullptr);
    const size_t limit = (size_t)gnngls::device_max_optin_smem();
    // One sweep per instance does not amortise staging an n x n fp64 matrix into shared memory (measured: 6x slower at
    // n=100); stage only a matrix shared by the whole batch.  local_search / GLS, which sweep many times, always stage.
    const int csize = o2a ? 1 : cluster_limit_for(B, n);
    if (csize != 1) {
        ClusterPlan plan;
        if (int rc = prepare_cluster_kernel(moves_cluster_kernel)) return rc;
        // (one sweep per launch: no symmetry verdict, so relocate could use the row cache for one of its two gathers only -- off)
        if (int rc = plan_cluster(reinterpret_cast<const void *>(moves_cluster_kernel), csize, B, n, plain, "moves_cluster_kernel", &plan,
                                  false)) return rc;
        if (plan.csize > 1)
            return launch_planned(moves_cluster_kernel, plan, st, op, o2a, D, stride, tours, pos, B, n, fi, out_delta, out_move,
                                  out_tours, plan.row_cache);
    }
    if (stride == 0 && staged <= limit) {
        if (int rc = ensure_smem(moves_kernel<true>, staged)) return rc;
        moves_kernel<true><<<grid_for(B), threads, staged, st>>>(op, o2a, D, stride, tours, pos, B, n, fi,
                                                                       out_delta, out_move, out_tours);
    } else {
        moves_kernel<false><<<grid_for(B), threads, plain, st>>>(op, o2a, D, stride, tours, pos, B, n, fi,
                                                                        out_delta, out_move, out_tours);
    }
    GNNGLS_LAUNCH_OK("moves_kernel");
    return GNNGLS_OK;
}

}  // namespace

extern "C" int gnngls_moves_eval_a2a(int op, const double *D, int64_t d_batch_stride, const int32_t *tours, int B,
                                     int n, int first_improvement, double *out_delta, int32_t *out_move,
                                     int32_t *out_tours, void *stream) {
    return launch_moves(op, false, D, d_batch_stride, tours, nullptr, B, n, first_improvement, out_delta, out_move,
                        out_tours, static_cast<cudaStream_t>(stream));
}

extern "C" int gnngls_moves_eval_o2a(int op, const double *D, int64_t d_batch_stride, const int32_t *tours,
                                     const int32_t *pos, int B, int n, int first_improvement, double *out_delta,
                                     int32_t *out_move, int32_t *out_tours, void *stream) {
    return launch_moves(op, true, D, d_batch_stride, tours, pos, B, n, first_improvement, out_delta, out_move,
                        out_tours, static_cast<cudaStream_t>(stream));
}

extern "C" int gnngls_local_search_batch(const double *D, int32_t *tours, double *costs, int B, int n,
                                         int first_improvement, double *events, int32_t *n_events, int max_events,
                                         int32_t *status, int64_t *counters, void *stream) {
    GNNGLS_REQUIRE(D && tours && costs, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(max_events >= 0, GNNGLS_ERR_BAD_ARG, "max_events < 0");
    if (int rc = check_n(n)) return rc;
    if (B <= 0) return GNNGLS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = pick_threads(n);
    const size_t staged = smem_layout(n, true, false, false, nullptr, nullptr);
    const size_t plain = smem_layout(n, false, false, false, nullptr, nullptr);
    const int csize = cluster_limit_for(B, n);
    if (csize != 1) {
        ClusterPlan plan;
        if (int rc = prepare_cluster_kernel(local_search_cluster_kernel)) return rc;
        if (int rc = plan_cluster(reinterpret_cast<const void *>(local_search_cluster_kernel), csize, B, n, plain,
                                  "local_search_cluster_kernel", &plan)) return rc;
        if (plan.csize > 1)
            return launch_planned(local_search_cluster_kernel, plan, st, D, tours, costs, B, n, first_improvement, events, n_events,
                                  max_events, status, reinterpret_cast<long long *>(counters), plan.row_cache);
    }
    if (staged <= (size_t)gnngls::device_max_optin_smem()) {
        if (int rc = ensure_smem(local_search_kernel<true>, staged)) return rc;
        local_search_kernel<true><<<grid_for(B), threads, staged, st>>>(
            D, tours, costs, B, n, first_improvement, events, n_events, max_events, status,
            reinterpret_cast<long long *>(counters));
    } else {
        local_search_kernel<false><<<grid_for(B), threads, plain, st>>>(
            D, tours, costs, B, n, first_improvement, events, n_events, max_events, status,
            reinterpret_cast<long long *>(counters));
    }
    GNNGLS_LAUNCH_OK("local_search_kernel");
    return GNNGLS_OK;
}

extern "C" size_t gnngls_sizeof_gls_args(void) { return sizeof(gnngls_gls_args); }

extern "C" int gnngls_gls_batch(const gnngls_gls_args *args, void *stream) {
    GNNGLS_REQUIRE(args, GNNGLS_ERR_BAD_ARG, "null args");
    const gnngls_gls_args &a = *args;
    GNNGLS_REQUIRE(a.D && a.guides && a.cur_tours && a.cur_costs && a.best_tours && a.best_costs && a.k,
                   GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(a.n_guides >= 1 && a.n_iters >= 0 && a.iter_begin >= 0 && a.perturbation_moves >= 0 &&
                       a.max_events >= 0,
                   GNNGLS_ERR_BAD_ARG, "bad scalar argument");
    GNNGLS_REQUIRE(a.guide_kind == GNNGLS_GUIDE_MATRIX_F64 || a.guide_kind == GNNGLS_GUIDE_EDGEVEC_F32,
                   GNNGLS_ERR_BAD_ARG, "bad guide_kind %d", a.guide_kind);
    GNNGLS_REQUIRE(!a.resume || a.penalties, GNNGLS_ERR_BAD_ARG, "resume needs a penalties buffer");
    if (int rc = check_n(a.n)) return rc;
    GNNGLS_REQUIRE(a.n >= 4 || a.n_iters == 0 || a.perturbation_moves == 0, GNNGLS_ERR_UNSUPPORTED,
                   "guided_local_search cannot make progress for n < 4 (the reference spins forever)");
    if (a.B <= 0) return GNNGLS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GlsDev P;
    P.a = a;
    P.row_cache = 0;
    const int threads = pick_threads(a.n);
    const size_t staged = smem_layout(a.n, true, true, true, nullptr, nullptr);
    const size_t plain = smem_layout(a.n, false, true, false, nullptr, nullptr);
    // Tier choice.  With D (fp64) + penalties staged in shared memory only 2 CTAs fit per SM at n=100; reading them
    // through L1/L2 instead lets 8 CTAs share an SM, which measured ~10% faster for this latency-bound loop
    // (profiles/r1_gls_tiers.md).  The global tier needs the caller's int32 penalties buffer; GNNGLS_GLS_TIER=staged|global
    // overrides.
    static int tier = -1;                                      // 0 auto, 1 staged, 2 global
    if (tier < 0) { const char *e = getenv("GNNGLS_GLS_TIER"); tier = !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'g' ? 2 : 0)); }
    const bool fits = staged <= (size_t)gnngls::device_max_optin_smem();
    const bool use_global = a.penalties && (tier == 2 || (tier == 0 && a.n >= 64)) ;
    const int csize = a.penalties ? cluster_limit_for(a.B, a.n) : 1;   // the cluster tier keeps the penalties in global memory
    if (csize != 1) {
        ClusterPlan plan;
        if (int rc = prepare_cluster_kernel(gls_cluster_kernel)) return rc;
        if (int rc = plan_cluster(reinterpret_cast<const void *>(gls_cluster_kernel), csize, a.B, a.n, plain, "gls_cluster_kernel", &plan))
            return rc;
        P.row_cache = plan.row_cache;
        if (plan.csize > 1) return launch_planned(gls_cluster_kernel, plan, st, P);
        P.row_cache = 0;
    }
    if (fits && !use_global) {
        if (int rc = ensure_smem(gls_kernel<true>, staged)) return rc;
        gls_kernel<true><<<grid_for(a.B), threads, staged, st>>>(P);
    } else {
        GNNGLS_REQUIRE(a.penalties, GNNGLS_ERR_WORKSPACE,
                       "n=%d does not fit shared memory: a [B,n,n] int32 penalties buffer is required", a.n);
        gls_kernel<false><<<grid_for(a.B), threads, plain, st>>>(P);
    }
    GNNGLS_LAUNCH_OK("gls_kernel");
    return GNNGLS_OK;
}

extern "C" int gnngls_nn_init_batch(int guide_kind, const void *guide, const double *D, int B, int n, int depot,
                                    int32_t *out_tours, double *out_costs, void *stream) {
    GNNGLS_REQUIRE(guide && out_tours, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(guide_kind == GNNGLS_GUIDE_MATRIX_F64 || guide_kind == GNNGLS_GUIDE_EDGEVEC_F32, GNNGLS_ERR_BAD_ARG,
                   "bad guide_kind %d", guide_kind);
    GNNGLS_REQUIRE(n >= 2 && n <= 1024, GNNGLS_ERR_UNSUPPORTED, "n=%d outside supported range [2,1024]", n);
    GNNGLS_REQUIRE(depot >= 0 && depot < n, GNNGLS_ERR_BAD_ARG, "depot out of range");
    if (B <= 0) return GNNGLS_OK;
    // few large instances: one CTA per instance (same condition as the cluster tier of the search kernels)
    if (n >= kClusterMinN && cluster_limit_for(B, n) != 1) {
        const int threads = (n + 31) / 32 * 32;
        const size_t smem_b = sizeof(double) * (n + 1 + 64) + sizeof(int) * (64 + n + 1) + 16;
        nn_init_block_kernel<<<B, threads, smem_b, static_cast<cudaStream_t>(stream)>>>(guide_kind, guide, D, B, n, depot,
                                                                                      out_tours, out_costs);
        GNNGLS_LAUNCH_OK("nn_init_block_kernel");
        return GNNGLS_OK;
    }
    const int wpb = 4;
    const size_t smem = (sizeof(double) + sizeof(int)) * (size_t)wpb * (n + 1) + 16;
    if (int rc = ensure_smem(nn_init_kernel, smem)) return rc;
    const int blocks = (B + wpb - 1) / wpb;
    nn_init_kernel<<<blocks, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(guide_kind, guide, D, B, n, depot,
                                                                              out_tours, out_costs);
    GNNGLS_LAUNCH_OK("nn_init_kernel");
    return GNNGLS_OK;
}

extern "C" int gnngls_tour_cost_batch(const double *D, const int32_t *tours, int B, int n, double *out_costs,
                                      void *stream) {
    GNNGLS_REQUIRE(D && tours && out_costs, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(n >= 1 && n <= 65535, GNNGLS_ERR_UNSUPPORTED, "n=%d unsupported", n);
    if (B <= 0) return GNNGLS_OK;
    const int wpb = 4;
    const size_t smem = sizeof(double) * (size_t)wpb * (n + 1);
    if (int rc = ensure_smem(tour_cost_kernel, smem)) return rc;
    tour_cost_kernel<<<(B + wpb - 1) / wpb, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(D, tours, B, n,
                                                                                              out_costs);
    GNNGLS_LAUNCH_OK("tour_cost_kernel");
    return GNNGLS_OK;
}

// debug: per-(cluster member, phase) cycle totals of the last cluster local_search launch (error unless built with -DGLS_STAMPS)
extern "C" int gnngls_debug_gls_stamps(unsigned long long *out, int count) {
#ifdef GLS_STAMPS
    if (count > 16 * 8) count = 16 * 8;
    GNNGLS_CUDA_OK(cudaMemcpyFromSymbol(out, g_gls_stamps, sizeof(unsigned long long) * count));
    return GNNGLS_OK;
#else
    (void)out; (void)count;
    return GNNGLS_ERR_UNSUPPORTED;
#endif
}
