// Dense contractions of the edge-regret model on sm_100a.
//
//   fc   : ft = h Wfc^T (+ attention scores el/er in the epilogue)      GATConv.fc, models.py:23
//   ff1  : hid = relu(h1 W1^T + b1)                                      models.py:30-31
//   ff2  : h   = BN2(h1 + hid W2^T + b2)                                 models.py:32-35
//
// Main path: one persistent, warp-specialised kernel template — TMA (cp.async.bulk.tensor, 128B
// swizzle) feeds a 6-stage shared-memory ring, a single thread issues tcgen05.mma kind::tf32 with
// fp32 accumulators in TMEM (double-buffered, 2 x 128 columns), eight epilogue warps drain TMEM with
// tcgen05.ld (thread == output row) and apply the fused epilogue straight from registers with
// 128-bit global accesses.  Operands are fp32 words in memory that already hold TF32-rounded values
// (weights are rounded once on the host, activations by the producing kernel into a second
// "_tf32" copy), so the tensor core's truncation of the low 13 mantissa bits is exact and the
// residual stream itself stays in full fp32.
//
// Debug path (GNNGLS_DENSE_SIMT): plain fp32 CUDA-core GEMM with the same epilogues, used by the
// tests to cross-check the tensor-core path.  Also here: embed_layer and decision_layer, which
// are K=in_dim / N=out_dim degenerate and run as fused elementwise / dot-product kernels.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdlib>
#include "common.h"

namespace {

constexpr int D_ = GNNGLS_EMBED_DIM;     // 128
constexpr int H_ = GNNGLS_HEADS;         // 8
constexpr int HID_ = GNNGLS_HIDDEN_DIM;  // 512
constexpr float kLog2e = 1.4426950408889634f;

enum { EPI_FC = 0, EPI_FF1 = 1, EPI_FF2 = 2 };

struct EpiParams {
    int64_t M;
    float *out;          // FC: ft [M,128]; FF1: hid [M,512]; FF2: h_out [M,128]
    __half *out_f16;     // FC only (nullable): store ft as fp16 [M,128] instead of fp32 `out`
    float *out_tf32;     // FF2 only (nullable): operand copy of h_out for the next tensor-core GEMM: fp32 rounded to TF32,
    int out_op_f16;      //   or (out_op_f16 != 0) the same pointer is a __half array and the copy is fp16
    float *el, *er;      // FC only, [M,8]
    const float *v0;     // FC: attn_l[128]; FF1: b1[512]; FF2: b2[128]
    const float *v1;     // FC: attn_r[128]; FF2: bn_scale[128]
    const float *v2;     // FF2: bn_shift[128]
    const float *skip;   // FF2: h1 [M,128] (fp32, unrounded)
    int round_tf32;      // FC: store ft rounded to TF32; FF1 (debug path): same for the hidden activations
    int skip_is_a;       // fused FF: the staged A operand is the skip tensor itself (same pointer)
};

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32_rna4(float4 v) {
    return make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
}
// two fp32 -> packed fp16x2 (lo in the low half), round-to-nearest-even, saturating to +-65504
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// operand copy of 4 consecutive activations at element offset `idx`: TF32-rounded fp32 or fp16
__device__ __forceinline__ void store_op_copy(float *base, int f16, int64_t idx, float4 v) {
    if (f16) *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(base) + idx) = make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
    else *reinterpret_cast<float4 *>(base + idx) = tf32_rna4(v);
}

// Fused epilogue for 4 consecutive columns [col, col+4) of one row (SIMT debug GEMM).  Called with a
// full warp in which lanes 4q..4q+3 hold consecutive column quads of the same row (FC reduction).
template <int EPI>
__device__ __forceinline__ void epilogue_quad(const EpiParams &p, int64_t row, int col, float4 acc, bool valid) {
    if (EPI == EPI_FC) {
        const float4 al = *reinterpret_cast<const float4 *>(p.v0 + col);
        const float4 ar = *reinterpret_cast<const float4 *>(p.v1 + col);
        float sl = acc.x * al.x + acc.y * al.y + acc.z * al.z + acc.w * al.w;
        float sr = acc.x * ar.x + acc.y * ar.y + acc.z * ar.z + acc.w * ar.w;
        sl += __shfl_xor_sync(0xffffffffu, sl, 1);
        sr += __shfl_xor_sync(0xffffffffu, sr, 1);
        sl += __shfl_xor_sync(0xffffffffu, sl, 2);
        sr += __shfl_xor_sync(0xffffffffu, sr, 2);
        if (valid) {
            if (p.out_f16) *reinterpret_cast<uint2 *>(p.out_f16 + row * D_ + col) = make_uint2(pack_f16x2(acc.x, acc.y), pack_f16x2(acc.z, acc.w));
            else *reinterpret_cast<float4 *>(p.out + row * D_ + col) = p.round_tf32 ? tf32_rna4(acc) : acc;
            if ((col & 15) == 0) {   // first quad of a head: 16 columns per head
                p.el[row * H_ + (col >> 4)] = sl * kLog2e;
                p.er[row * H_ + (col >> 4)] = sr * kLog2e;
            }
        }
    } else if (EPI == EPI_FF1) {
        if (!valid) return;
        const float4 b = *reinterpret_cast<const float4 *>(p.v0 + col);
        float4 o;
        o.x = fmaxf(acc.x + b.x, 0.f); o.y = fmaxf(acc.y + b.y, 0.f);
        o.z = fmaxf(acc.z + b.z, 0.f); o.w = fmaxf(acc.w + b.w, 0.f);
        if (p.round_tf32) o = tf32_rna4(o);
        *reinterpret_cast<float4 *>(p.out + row * HID_ + col) = o;
    } else {
        if (!valid) return;
        const float4 b = *reinterpret_cast<const float4 *>(p.v0 + col);
        const float4 sc = *reinterpret_cast<const float4 *>(p.v1 + col);
        const float4 sh = *reinterpret_cast<const float4 *>(p.v2 + col);
        const float4 s = *reinterpret_cast<const float4 *>(p.skip + row * D_ + col);
        float4 o;
        o.x = (s.x + (acc.x + b.x)) * sc.x + sh.x; o.y = (s.y + (acc.y + b.y)) * sc.y + sh.y;
        o.z = (s.z + (acc.z + b.z)) * sc.z + sh.z; o.w = (s.w + (acc.w + b.w)) * sc.w + sh.w;
        *reinterpret_cast<float4 *>(p.out + row * D_ + col) = o;
        if (p.out_tf32) store_op_copy(p.out_tf32, p.out_op_f16, row * D_ + col, o);
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-private 32x32 fp32 staging tile (4 KB, 128-byte rows, 16-byte pieces XOR-swizzled by row&7):
// thread-per-row accesses (what tcgen05.ld produces) and 4-rows-per-instruction coalesced accesses
// are both bank-conflict free, so global memory is only ever touched with full 128-byte lines.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 *stg_slot(float *stg, int row, int piece) {
    return reinterpret_cast<float4 *>(stg + row * 32 + ((piece ^ (row & 7)) << 2));
}
// coalesced global -> staging: 32 rows x 32 columns starting at (row0, col0) of a row-major [M, ld] matrix
__device__ __forceinline__ void stg_load_tile(float *stg, const float *src, int64_t row0, int col0, int ld, int64_t M, int lane) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), pc = lane & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) v = *reinterpret_cast<const float4 *>(src + (row0 + r) * ld + col0 + 4 * pc);
        *stg_slot(stg, r, pc) = v;
    }
}
// staging -> coalesced global (optionally a second, TF32-rounded copy)
__device__ __forceinline__ void stg_store_tile(const float *stg, float *dst, float *dst_tf32, int64_t row0, int col0, int ld,
                                               int64_t M, int lane, int op_f16 = 0) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), pc = lane & 7;
        if (row0 + r < M) {
            const float4 v = *stg_slot(const_cast<float *>(stg), r, pc);
            *reinterpret_cast<float4 *>(dst + (row0 + r) * ld + col0 + 4 * pc) = v;
            if (dst_tf32) store_op_copy(dst_tf32, op_f16, (row0 + r) * ld + col0 + 4 * pc, v);
        }
    }
}

// Tensor-core epilogue for 32 accumulator columns [col, col+32) of the 32 rows starting at row0 that
// this warp owns (thread == row out of tcgen05.ld), through the warp's staging tile.
template <int EPI>
__device__ __forceinline__ void epilogue_tile32(const EpiParams &p, const float *sv, float *stg, int64_t row0, int col,
                                                float (&v)[32], int lane) {
    const int64_t row = row0 + lane;
    if (EPI == EPI_FC) {
        float sl[2] = {0.f, 0.f}, sr[2] = {0.f, 0.f};       // 32 columns = two heads of 16
        const bool f16 = p.out_f16 != nullptr;
        // fp16 output: the tile holds this warp's 64 columns as 32 rows x 128 bytes; this call fills half of it
        const int piece0 = (col & 32) >> 3;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 al = *reinterpret_cast<const float4 *>(sv + col + 4 * j);
            const float4 ar = *reinterpret_cast<const float4 *>(sv + D_ + col + 4 * j);
            sl[j >> 2] += v[4 * j] * al.x + v[4 * j + 1] * al.y + v[4 * j + 2] * al.z + v[4 * j + 3] * al.w;
            sr[j >> 2] += v[4 * j] * ar.x + v[4 * j + 1] * ar.y + v[4 * j + 2] * ar.z + v[4 * j + 3] * ar.w;
            if (f16) {
                if (j & 1) continue;
                *reinterpret_cast<uint4 *>(stg_slot(stg, lane, piece0 + (j >> 1))) =
                    make_uint4(pack_f16x2(v[4 * j], v[4 * j + 1]), pack_f16x2(v[4 * j + 2], v[4 * j + 3]),
                               pack_f16x2(v[4 * j + 4], v[4 * j + 5]), pack_f16x2(v[4 * j + 6], v[4 * j + 7]));
            } else {
                float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (p.round_tf32) o = tf32_rna4(o);         // ft only feeds the aggregate's tensor-core B operand
                *stg_slot(stg, lane, j) = o;
            }
        }
        if (row < p.M) {
            *reinterpret_cast<float2 *>(p.el + row * H_ + (col >> 4)) = make_float2(sl[0] * kLog2e, sl[1] * kLog2e);
            *reinterpret_cast<float2 *>(p.er + row * H_ + (col >> 4)) = make_float2(sr[0] * kLog2e, sr[1] * kLog2e);
        }
        if (f16) {
            if (col & 32) {                                 // second half staged: store 32 rows x 128 B, 4 rows per instruction
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + (lane >> 3), pc = lane & 7;
                    if (row0 + r < p.M)
                        *reinterpret_cast<uint4 *>(p.out_f16 + (row0 + r) * D_ + (col & ~63) + 8 * pc) =
                            *reinterpret_cast<const uint4 *>(stg_slot(stg, r, pc));
                }
                __syncwarp();
            }
        } else {
            __syncwarp();
            stg_store_tile(stg, p.out, nullptr, row0, col, D_, p.M, lane);
            __syncwarp();
        }
    } else if (EPI == EPI_FF1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = *reinterpret_cast<const float4 *>(sv + col + 4 * j);
            float4 r;
            r.x = fmaxf(v[4 * j] + b.x, 0.f); r.y = fmaxf(v[4 * j + 1] + b.y, 0.f);
            r.z = fmaxf(v[4 * j + 2] + b.z, 0.f); r.w = fmaxf(v[4 * j + 3] + b.w, 0.f);
            if (p.round_tf32) r = tf32_rna4(r);
            *stg_slot(stg, lane, j) = r;
        }
        __syncwarp();
        stg_store_tile(stg, p.out, nullptr, row0, col, HID_, p.M, lane);
        __syncwarp();
    } else {
        stg_load_tile(stg, p.skip, row0, col, D_, p.M, lane);          // skip connection (fp32 h1), coalesced
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 s = *stg_slot(stg, lane, j);
            const float4 b = *reinterpret_cast<const float4 *>(sv + col + 4 * j);
            const float4 sc = *reinterpret_cast<const float4 *>(sv + D_ + col + 4 * j);
            const float4 sh = *reinterpret_cast<const float4 *>(sv + 2 * D_ + col + 4 * j);
            float4 r;
            r.x = (s.x + (v[4 * j] + b.x)) * sc.x + sh.x; r.y = (s.y + (v[4 * j + 1] + b.y)) * sc.y + sh.y;
            r.z = (s.z + (v[4 * j + 2] + b.z)) * sc.z + sh.z; r.w = (s.w + (v[4 * j + 3] + b.w)) * sc.w + sh.w;
            *stg_slot(stg, lane, j) = r;
        }
        __syncwarp();
        stg_store_tile(stg, p.out, p.out_tf32, row0, col, D_, p.M, lane, p.out_op_f16);
        __syncwarp();
    }
}

// ================================================================================================
// PTX wrappers (sm_100a)
// ================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one lane of a converged warp (lets the compiler keep MMA operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by one thread
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 operands, K = 16 per instruction)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// same, arriving on the barrier at this offset in every CTA of `mask` (operands were multicast to all of them)
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (ignored for swizzled K-major, 1) | [32,46) SBO>>4 = 1024B>>4
//   [46,48) version = 1 | [61,64) layout = SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// cute::UMMA::InstrDescriptor: c_format=F32 (1<<4), a/b_format=TF32 (2<<7, 2<<10), K-major A and B,
// n_dim = N>>3 at bit 17, m_dim = M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 instruction descriptor: F32 accumulate, A/B = F16 (format 0), K-major A and B
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ================================================================================================
// tcgen05 GEMM:  C[M, N_TOTAL] = A[M, K_TOTAL] * W[N_TOTAL, K_TOTAL]^T  with fused epilogue
// ================================================================================================
constexpr int BM = 128, BN = 128, BK = 32;           // BK fp32 = 128 bytes = one swizzle row
constexpr int STAGES = 5;
constexpr int STAGE_BYTES = (BM + BN) * BK * 4;      // 32 KB
constexpr int EPI_WARPS = 8;                         // two per TMEM lane quarter, 64 columns each
constexpr int GEMM_THREADS = (2 + EPI_WARPS) * 32;   // warp0 TMA, warp1 MMA/TMEM, warps 2..9 epilogue
constexpr int TMEM_COLS = 256;                       // two 128-column fp32 accumulators
constexpr int SVEC_FLOATS = 512;                     // per-column epilogue vectors staged in smem
constexpr int STG_TILE_BYTES = 32 * 32 * 4;              // per-warp epilogue staging tile
constexpr size_t GEMM_SMEM = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + EPI_WARPS * STG_TILE_BYTES + SVEC_FLOATS * 4 + 256;

// F16: both operands are fp16 in global memory (a 128-byte swizzle row then holds 64 k), MMAs are kind::f16
template <int N_TOTAL, int K_TOTAL, int EPI, bool F16 = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiParams p) {
    constexpr int BKE = F16 ? 2 * BK : BK;               // k elements per 128-byte row
    constexpr int KB = K_TOTAL / BKE;
    constexpr int NB = N_TOTAL / BN;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024B-aligned, still .shared
    unsigned char *stage_base = smem;
    float *staging = reinterpret_cast<float *>(smem + (size_t)STAGES * STAGE_BYTES);
    float *svec = staging + EPI_WARPS * (STG_TILE_BYTES / 4);
    uint64_t *bars = reinterpret_cast<uint64_t *>(svec + SVEC_FLOATS);
    uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m_tiles = (p.M + BM - 1) / BM;
    const int64_t n_work = m_tiles * NB;

    // per-column epilogue vectors -> shared memory
    if (EPI == EPI_FC) {
        for (int c = threadIdx.x; c < D_; c += GEMM_THREADS) { svec[c] = p.v0[c]; svec[D_ + c] = p.v1[c]; }
    } else if (EPI == EPI_FF1) {
        for (int c = threadIdx.x; c < HID_; c += GEMM_THREADS) svec[c] = p.v0[c];
    } else {
        for (int c = threadIdx.x; c < D_; c += GEMM_THREADS) { svec[c] = p.v0[c]; svec[D_ + c] = p.v1[c]; svec[2 * D_ + c] = p.v2[c]; }
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int m_blk = (int)(w / NB), n_blk = (int)(w % NB);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char *sa = stage_base + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(&tmA, &full[stage], sa, kb * BKE, m_blk * BM);
                    tma_load_2d(&tmB, &full[stage], sa + BM * BK * 4, kb * BKE, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (warp-uniform control flow)
        constexpr uint32_t idesc = F16 ? make_idesc_f16(BM, BN) : make_idesc_tf32(BM, BN);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(stage_base + (size_t)stage * STAGE_BYTES);
                const uint64_t da = make_sw128_kmajor_desc(sa);
                const uint64_t db = make_sw128_kmajor_desc(sa + BM * BK * 4);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {       // one MMA per 32 bytes along K: 8 tf32 or 16 fp16 values
                        if (F16) umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                        else umma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);              // frees the smem slot when these MMAs finish
                    if (kb == KB - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int q = warp & 3;                               // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                     // which 64 accumulator columns
        float *stg = staging + (warp - 2) * (STG_TILE_BYTES / 4);
        uint32_t acc = 0, acc_phase = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int m_blk = (int)(w / NB), n_blk = (int)(w % NB);
            const int64_t row0 = (int64_t)m_blk * BM + q * 32;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int col_in_tile = half * 64 + c * 32;
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + col_in_tile, v);
                epilogue_tile32<EPI>(p, svec, stg, row0, n_blk * BN + col_in_tile, v, lane);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// row-major fp32 [rows, cols] -> box [box_rows rows, 32 cols], 128B swizzle, zero fill out of bounds
int make_map(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint32_t box_rows = BM) {
    EncodeTiledFn fn = get_encode_fn();
    GNNGLS_REQUIRE(fn, GNNGLS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    GNNGLS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, GNNGLS_ERR_BAD_ARG, "TMA operand not 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GNNGLS_REQUIRE(r == CUDA_SUCCESS, GNNGLS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return GNNGLS_OK;
}

// row-major fp16 [rows, cols] -> box [box_rows rows, 64 cols] (128 bytes), 128B swizzle
int make_map_f16(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    GNNGLS_REQUIRE(fn, GNNGLS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    GNNGLS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, GNNGLS_ERR_BAD_ARG, "TMA operand not 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GNNGLS_REQUIRE(r == CUDA_SUCCESS, GNNGLS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return GNNGLS_OK;
}

template <int N_TOTAL, int K_TOTAL, int EPI, bool F16 = false>
int launch_tc_gemm(const void *A, const void *W, const EpiParams &p, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    if (F16) {
        if (int rc = make_map_f16(&tmA, A, (uint64_t)p.M, K_TOTAL, BM)) return rc;
        if (int rc = make_map_f16(&tmB, W, N_TOTAL, K_TOTAL, BN)) return rc;
    } else {
        if (int rc = make_map(&tmA, static_cast<const float *>(A), (uint64_t)p.M, K_TOTAL)) return rc;
        if (int rc = make_map(&tmB, static_cast<const float *>(W), N_TOTAL, K_TOTAL)) return rc;
    }
    auto kern = gemm_tf32_kernel<N_TOTAL, K_TOTAL, EPI, F16>;
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    const int64_t work = ((p.M + BM - 1) / BM) * (N_TOTAL / BN);
    const int sms = gnngls::device_sm_count();
    const int grid = (int)(work < sms ? work : sms);
    kern<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(tmA, tmB, p);
    GNNGLS_LAUNCH_OK("gemm_tf32_kernel");
    return GNNGLS_OK;
}

// ================================================================================================
// Fused feed-forward block on tcgen05:  h_out = BN2(h1 + relu(h1 W1^T + b1) W2^T + b2)
//
// One persistent CTA per 128-row tile; the 128x512 hidden activation never leaves the SM and BOTH
// left-hand operands live in tensor memory, so shared memory only carries the streamed weights:
//   * A tile (TF32 copy of h1): TMA -> smem (64 KB) -> four warps copy it into TMEM (tcgen05.st,
//     lane = row, column = k); the smem buffer is released at once, so the next tile's TMA overlaps
//     the whole tile.  (Re-reading A from smem for every hidden chunk made the first version
//     shared-memory-port bound: tools/umma_bench.cu measures 40 cycles per N=32 smem-A MMA.)
//   * W1 / W2 stream through a 7-stage TMA ring of 16 KB half-chunks (64 hidden units per chunk);
//   * GEMM1(c): DH[c&1] (TMEM, 64 cols) = A[tmem] . W1c^T         (16 x tcgen05.mma 128x64x8)
//   * epilogue-1 (8 warps, 32 rows x 32 units each): tcgen05.ld -> +b1, ReLU, cvt.rna.tf32 -> tcgen05.st back
//     IN PLACE: the hidden chunk becomes the TMEM A operand of the second contraction;
//   * GEMM2(c): D2[tile&1] (TMEM, 128 cols) += H[tmem] . W2c^T    (8 x tcgen05.mma 128x128x8)
//   * final epilogue (the four A-staging warps, one tile behind, D2 double-buffered): tcgen05.ld ->
//     +b2 + skip(h1 fp32) -> BN2 -> h_out (+ TF32 copy) through a swizzled staging tile so that
//     global memory only sees full 128-byte lines.
// MMAs of one thread execute in order, so DH[g] needs no "empty" barriers: GEMM1(c+2) is issued after
// GEMM2(c).  The MMA warp issues GEMM1(c+1) before GEMM2(c) so the tensor core works during epilogue-1.
// TMEM columns: D2[0] [0,128) | D2[1] [128,256) | DH[g] [256+64g,+64) | A [384,512).
// ================================================================================================
constexpr int FF_HC = 64;                              // hidden units per chunk
constexpr int FF_CHUNKS = HID_ / FF_HC;                // 8
constexpr int FF_WSTAGES = 7;
constexpr int FF_WSTAGES_F16 = 4;                      // fp16 weights: one stage = a whole chunk of W1 or W2
constexpr int FF_WSTAGE_BYTES = 16384;                 // half of W1c: 2 boxes [64 x 32]; half of W2c: 1 box [128 x 32]
constexpr int FF_A_BYTES = BM * D_ * 4;                // 64 KB: 4 boxes [128 x 32]
constexpr int FF_SVEC = 4 * D_ + HID_;                 // b2 | bn_scale | bn_shift | b1 | b2*bn_scale + bn_shift
constexpr int FF_THREADS = 15 * 32;                    // warp0 W-TMA, warp1 MMA, warps 2..9 epilogue-1, warp10 A-TMA, warps 11..14 A staging + final epilogue
constexpr int FF_NBARS = 4 + 2 * FF_WSTAGES + 8 + 4 + 16;
constexpr int FF_THREADS_F16 = 19 * 32;                // fp16 variant: + warps 15..18, final epilogue only
constexpr int FF_TMEM_COLS = 512;
constexpr size_t ff_smem_bytes(bool f16) {
    return 1024 + (f16 ? 2 : 1) * (size_t)FF_A_BYTES + (size_t)(f16 ? FF_WSTAGES_F16 : FF_WSTAGES) * FF_WSTAGE_BYTES +
           4 * STG_TILE_BYTES + FF_SVEC * 4 + FF_NBARS * 8 + 16;
}

// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16: K = 16 per instruction, A: lane = row, one 32-bit column per PAIR of k
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32 (A: lane = row, one fp32 column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp writes TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// F16: weights arrive as fp16, both TMEM operands are packed fp16 pairs and the contractions run as kind::f16
// (K = 16 per instruction: half the MMAs and half the streamed weight bytes; fp16 carries the same 10-bit
// mantissa as the TF32 operands of the other variant, values saturate at +-65504).
template <int CL, bool F16>      // CL: CTAs per cluster sharing (multicasting) the streamed weights: 1, 2 or 4
__global__ void __launch_bounds__(F16 ? FF_THREADS_F16 : FF_THREADS, 1)
ff_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const EpiParams p, const float *__restrict__ b1,
                     const int round_a) {
    static_assert(!F16 || CL == 1, "the fp16 variant does not multicast");
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024B-aligned, still .shared
    // fp16 variant: the weights need half the ring, which pays for a second A buffer: A(t) stays in shared memory
    // until the final epilogue of tile t has taken its skip connection from it (no second trip to HBM for h1)
    constexpr int WST = F16 ? FF_WSTAGES_F16 : FF_WSTAGES;
    constexpr int NA = F16 ? 2 : 1;
    unsigned char *sA = smem;
    unsigned char *sW = sA + NA * FF_A_BYTES;
    float *staging = reinterpret_cast<float *>(sW + (size_t)WST * FF_WSTAGE_BYTES);
    float *svec = staging + 4 * (STG_TILE_BYTES / 4);
    uint64_t *bars = reinterpret_cast<uint64_t *>(svec + FF_SVEC);
    uint64_t *a_full = bars, *a_empty = bars + 1, *at_full = bars + 2, *at_empty = bars + 3;
    uint64_t *w_full = bars + 4, *w_empty = w_full + FF_WSTAGES;
    uint64_t *d1_full = w_empty + FF_WSTAGES, *h_full = d1_full + 2, *d2_full = h_full + 2, *d2_empty = d2_full + 2;
    uint64_t *a_full2 = d2_empty + 2, *a_empty2 = d2_empty + 3;      // second A buffer in shared memory (fp16 variant)
    uint64_t *at_full2 = d2_empty + 4, *at_empty2 = d2_empty + 5;    // second A buffer in tensor memory (fp16 variant)
    // fp16 variant: the A buffers are handed over per 32-column k-block [buffer][kb], so the next tile's TMA boxes start
    // landing as soon as the final epilogue has finished with the corresponding output chunk
    uint64_t *ak_full = d2_empty + 6, *ak_empty = d2_empty + 14;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d2_empty + 22);
    constexpr int NTHREADS = F16 ? FF_THREADS_F16 : FF_THREADS;
    const bool skip_smem = F16 && p.skip_is_a;                       // the staged operand IS the fp32 skip tensor

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Every CTA of a cluster walks the same number of tile groups (the weight ring is shared); a CTA whose
    // tile index falls past the end runs a dummy tile (TMA zero-fills, stores are masked by row < M).
    const int64_t m_tiles_real = (p.M + BM - 1) / BM;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    const int64_t tile0 = (int64_t)(blockIdx.x / CL) * CL + crank, tile_step = gridDim.x;
    const int64_t m_tiles = (m_tiles_real + CL - 1) / CL * CL;       // loops run `w < m_tiles` with w = tile0 + k*tile_step
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);

    for (int c = threadIdx.x; c < D_; c += NTHREADS) {
        svec[c] = p.v0[c]; svec[D_ + c] = p.v1[c]; svec[2 * D_ + c] = p.v2[c];
        svec[3 * D_ + HID_ + c] = fmaf(p.v0[c], p.v1[c], p.v2[c]);    // b2*scale + shift: (skip + acc + b2)*scale + shift in two ops
    }
    for (int c = threadIdx.x; c < HID_; c += NTHREADS) svec[3 * D_ + c] = b1[c];
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
        constexpr int FE_WARPS = F16 ? 8 : 4;                 // warps that share a tile's final epilogue
        mbar_init(a_full, 1); mbar_init(a_empty, F16 ? FE_WARPS : 4); mbar_init(at_full, 4); mbar_init(at_empty, 1);
        mbar_init(a_full2, 1); mbar_init(a_empty2, FE_WARPS); mbar_init(at_full2, 4); mbar_init(at_empty2, 1);
        for (int s2 = 0; s2 < 8; ++s2) { mbar_init(&ak_full[s2], 1); mbar_init(&ak_empty[s2], 4); }
        for (int s = 0; s < WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], CL); }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&d1_full[g], 1); mbar_init(&h_full[g], EPI_WARPS);
            mbar_init(&d2_full[g], 1); mbar_init(&d2_empty[g], F16 ? 8 : 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, FF_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                           // peers' barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_d2 = tmem_base, tm_dh = tmem_base + 256, tm_a = tmem_base + 384;   // fp16: A[0] [384,448) | A[1] [448,512)

    if (warp == 10) {
        // ------------------------------------------------------------------ A-tile TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t w = tile0; w < m_tiles; w += tile_step, ++it) {
                const uint32_t ab = F16 ? (it & 1) : 0, an = F16 ? (it >> 1) : it;     // buffer, use count of that buffer
                if (F16) {
                    for (int o = 0; o < 4; ++o) {
                        const int kb = ((o & 1) << 1) | (o >> 1);    // 0, 2, 1, 3: the order the final epilogue frees them
                        mbar_wait(&ak_empty[ab * 4 + kb], (an & 1) ^ 1);
                        mbar_expect_tx(&ak_full[ab * 4 + kb], BM * BK * 4);
                        tma_load_2d(&tmA, &ak_full[ab * 4 + kb], sA + ab * FF_A_BYTES + kb * (BM * BK * 4), kb * BK, (int)w * BM);
                    }
                    continue;
                }
                uint64_t *af = ab ? a_full2 : a_full, *ae = ab ? a_empty2 : a_empty;
                mbar_wait(ae, (an & 1) ^ 1);                  // the previous user of this buffer is done with it
                mbar_expect_tx(af, FF_A_BYTES);
                for (int kb = 0; kb < D_ / BK; ++kb)
                    tma_load_2d(&tmA, af, sA + ab * FF_A_BYTES + kb * (BM * BK * 4), kb * BK, (int)w * BM);
            }
        }
    } else if (warp >= 11) {
        // ------------------------------------------------------------------ A staging (smem -> TMEM) + final epilogue, one tile behind
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        float *stg = staging + ((warp - 11) & 3) * (STG_TILE_BYTES / 4);
        // tile index w, local tile counter t, 32-column chunks [c_begin, c_end)
        auto final_epilogue = [&](int64_t w, uint32_t t, int c_begin, int c_end) {
            const uint32_t d = t & 1;
            mbar_wait(&d2_full[d], (t >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c2 = c_begin; c2 < c_end; ++c2) {
                float v[32];
                tmem_ld_32x32(tm_d2 + lane_sel + d * BN + c2 * 32, v);
                if (skip_smem) {
                    // skip connection straight from the A tile still resident in shared memory (TMA 128B-swizzled
                    // layout: k-block c2, row r, 16-byte pieces XOR-ed with r & 7), then the usual coalesced store
                    // The 32-row x 128-byte block this warp reads has exactly the geometry of a staging tile, so the
                    // results are written back in place and leave through the coalesced tile store.
                    float *blk = reinterpret_cast<float *>(sA + (t & 1) * FF_A_BYTES + c2 * (BM * BK * 4) + q * 32 * 128);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 sk = *stg_slot(blk, lane, j);
                        const float4 sc = *reinterpret_cast<const float4 *>(svec + D_ + c2 * 32 + 4 * j);
                        const float4 sh = *reinterpret_cast<const float4 *>(svec + 3 * D_ + HID_ + c2 * 32 + 4 * j);
                        float4 o;
                        o.x = fmaf(sk.x + v[4 * j], sc.x, sh.x); o.y = fmaf(sk.y + v[4 * j + 1], sc.y, sh.y);
                        o.z = fmaf(sk.z + v[4 * j + 2], sc.z, sh.z); o.w = fmaf(sk.w + v[4 * j + 3], sc.w, sh.w);
                        *stg_slot(blk, lane, j) = o;
                    }
                    __syncwarp();
                    stg_store_tile(blk, p.out, p.out_tf32, w * BM + q * 32, c2 * 32, D_, p.M, lane, p.out_op_f16);
                    __syncwarp();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our smem accesses before the next TMA box lands here
                    if (lane == 0) mbar_arrive(&ak_empty[(t & 1) * 4 + c2]);         // k-block c2 of this tile's A buffer is free
                } else {
                    epilogue_tile32<EPI_FF2>(p, svec, stg, w * BM + q * 32, c2 * 32, v, lane);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[d]);
        };
        if (F16 && warp >= 15) {
            // ---- fp16 variant: warps 15..18 take output columns [0,64) of every tile's final epilogue; the staging
            // warps 11..14 take [64,128) of the previous tile after they have put the next A into tensor memory
            uint32_t it = 0;
            for (int64_t w = tile0; w < m_tiles; w += tile_step, ++it) final_epilogue(w, it, 0, 2);
        } else {
        uint32_t it = 0;
        int64_t w_prev = -1;
        for (int64_t w = tile0; w < m_tiles; w += tile_step, ++it) {
            const uint32_t ab = F16 ? (it & 1) : 0, an = F16 ? (it >> 1) : it;
            if (!F16) mbar_wait(a_full, an & 1);
            mbar_wait(ab ? at_empty2 : at_empty, (an & 1) ^ 1);   // GEMM1 of the previous user has finished reading this A[tmem]
            tc_fence_after();
#pragma unroll 1
            for (int o = 0; o < D_ / BK; ++o) {
                const int kb = F16 ? (((o & 1) << 1) | (o >> 1)) : o;
                if (F16) mbar_wait(&ak_full[ab * 4 + kb], an & 1);
                float v[32];
                const unsigned char *row = sA + ab * FF_A_BYTES + kb * (BM * BK * 4) + r * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 x = *reinterpret_cast<const float4 *>(row + ((j ^ (r & 7)) << 4));
                    if (round_a) x = tf32_rna4(x);            // operand is the raw fp32 activation: round here, not truncate in the MMA
                    v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
                }
                if (F16) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_f16x2(v[2 * j], v[2 * j + 1]);
                    tmem_st_32x16(tm_a + ab * 64 + lane_sel + kb * (BK / 2), pk);
                } else {
                    tmem_st_32x32(tm_a + lane_sel + kb * BK, v);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (!F16) mbar_arrive(a_empty);               // (fp16 variant: released by the tile's final epilogue)
                mbar_arrive(ab ? at_full2 : at_full);
            }
            if (w_prev >= 0) final_epilogue(w_prev, it - 1, F16 ? 2 : 0, 4);
            w_prev = w;
        }
        if (w_prev >= 0) final_epilogue(w_prev, it - 1, F16 ? 2 : 0, 4);
        }
    } else if (warp == 0) {
        // ------------------------------------------------------------------ weight producer (ring order == MMA order)
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            // Each stage is 16 KB = CL sub-boxes of [BR rows x 32 k]; this CTA fetches sub-box `crank` and
            // multicasts it to the whole cluster, so every weight byte crosses L2->SM once per cluster.
            constexpr int BR = 128 / CL;                      // rows per sub-box (W tensor maps are built with this box height)
            auto load_w1 = [&](int c, int half) {             // hidden units [64c, 64c+64), K range [64*half, +64): 2 x [64 rows x 32 k]
                mbar_wait(&w_empty[stage], phase ^ 1);
                unsigned char *dst = sW + (size_t)stage * FF_WSTAGE_BYTES;
                mbar_expect_tx(&w_full[stage], FF_WSTAGE_BYTES);
                if (CL == 1) {
                    for (int kk = 0; kk < 2; ++kk) tma_load_2d(&tmW1, &w_full[stage], dst + kk * (FF_HC * BK * 4), (2 * half + kk) * BK, c * FF_HC);
                } else {
                    constexpr int per = BR >= 64 ? 1 : 64 / BR;  // sub-boxes per [64 x 32] W1 block
                    const int kk = (int)crank / per, r0 = ((int)crank % per) * BR;
                    tma_load_2d_mc(&tmW1, &w_full[stage], dst + kk * (FF_HC * BK * 4) + r0 * 128, (2 * half + kk) * BK, c * FF_HC + r0, kMask);
                }
                if (++stage == WST) { stage = 0; phase ^= 1; }
            };
            auto load_w2 = [&](int c, int half) {             // all 128 outputs, hidden units [64c + 32*half, +32): [128 rows x 32 k]
                mbar_wait(&w_empty[stage], phase ^ 1);
                unsigned char *dst = sW + (size_t)stage * FF_WSTAGE_BYTES;
                mbar_expect_tx(&w_full[stage], FF_WSTAGE_BYTES);
                if (CL == 1) tma_load_2d(&tmW2, &w_full[stage], dst, c * FF_HC + half * BK, 0);
                else tma_load_2d_mc(&tmW2, &w_full[stage], dst + (int)crank * BR * 128, c * FF_HC + half * BK, (int)crank * BR, kMask);
                if (++stage == WST) { stage = 0; phase ^= 1; }
            };
            // fp16: one stage holds a whole chunk of W1 (2 boxes [64 units x 64 k]) or of W2 (1 box [128 outputs x 64 units])
            auto load_w1_f16 = [&](int c) {
                mbar_wait(&w_empty[stage], phase ^ 1);
                unsigned char *dst = sW + (size_t)stage * FF_WSTAGE_BYTES;
                mbar_expect_tx(&w_full[stage], FF_WSTAGE_BYTES);
                for (int kk = 0; kk < 2; ++kk) tma_load_2d(&tmW1, &w_full[stage], dst + kk * (FF_HC * 128), kk * 64, c * FF_HC);
                if (++stage == WST) { stage = 0; phase ^= 1; }
            };
            auto load_w2_f16 = [&](int c) {
                mbar_wait(&w_empty[stage], phase ^ 1);
                mbar_expect_tx(&w_full[stage], FF_WSTAGE_BYTES);
                tma_load_2d(&tmW2, &w_full[stage], sW + (size_t)stage * FF_WSTAGE_BYTES, c * FF_HC, 0);
                if (++stage == WST) { stage = 0; phase ^= 1; }
            };
            for (int64_t w = tile0; w < m_tiles; w += tile_step) {
                if (F16) {
                    for (int c = 0; c < FF_CHUNKS; ++c) {
                        load_w1_f16(c);
                        if (c >= 1) load_w2_f16(c - 1);
                    }
                    load_w2_f16(FF_CHUNKS - 1);
                } else {
                    for (int c = 0; c < FF_CHUNKS; ++c) {
                        load_w1(c, 0); load_w1(c, 1);
                        if (c >= 1) { load_w2(c - 1, 0); load_w2(c - 1, 1); }
                    }
                    load_w2(FF_CHUNKS - 1, 0); load_w2(FF_CHUNKS - 1, 1);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (warp-uniform control flow)
        constexpr uint32_t idesc1 = F16 ? make_idesc_f16(BM, FF_HC) : make_idesc_tf32(BM, FF_HC);
        constexpr uint32_t idesc2 = F16 ? make_idesc_f16(BM, BN) : make_idesc_tf32(BM, BN);
        uint32_t stage = 0, phase = 0, it = 0;
        uint32_t n_h[2] = {0, 0};
        for (int64_t w = tile0; w < m_tiles; w += tile_step, ++it) {
            const uint32_t d = it & 1;
            const uint32_t tm_acc = tm_d2 + d * BN;
            auto gemm2 = [&](int c) {
                const int g = c & 1;
                mbar_wait(&h_full[g], n_h[g] & 1); ++n_h[g];  // epilogue-1 has rewritten DH[g] with the hidden chunk
                if (F16) {
                    mbar_wait(&w_full[stage], phase);
                    tc_fence_after();
                    const uint64_t db = make_sw128_kmajor_desc(smem_u32(sW + (size_t)stage * FF_WSTAGE_BYTES));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < FF_HC / 16; ++k)     // hidden units [16k, 16k+16): packed by the epilogue warp hf = k>>1 at columns 32*hf + 8*(k&1)
                            umma_f16_ts(tm_acc, tm_dh + g * FF_HC + (k >> 1) * 32 + (k & 1) * 8, db + (uint64_t)(2 * k), idesc2, (c | k) != 0);
                        umma_commit(&w_empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == WST) { stage = 0; phase ^= 1; }
                } else {
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        mbar_wait(&w_full[stage], phase);
                        tc_fence_after();
                        const uint64_t db = make_sw128_kmajor_desc(smem_u32(sW + (size_t)stage * FF_WSTAGE_BYTES));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 8; ++k)
                                umma_tf32_ts(tm_acc, tm_dh + g * FF_HC + half * BK + 8 * k, db + (uint64_t)(2 * k), idesc2, (c | half | k) != 0);
                            if (CL == 1) umma_commit(&w_empty[stage]); else umma_commit_mc(&w_empty[stage], kMask);
                        }
                        __syncwarp();
                        if (++stage == WST) { stage = 0; phase ^= 1; }
                    }
                }
            };
            const uint32_t ab = F16 ? (it & 1) : 0, an = F16 ? (it >> 1) : it;
            mbar_wait(ab ? at_full2 : at_full, an & 1);       // A tile is in TMEM
            for (int c = 0; c < FF_CHUNKS; ++c) {
                const int g = c & 1;
                if (F16) {
                    mbar_wait(&w_full[stage], phase);
                    tc_fence_after();
                    const uint32_t sw = smem_u32(sW + (size_t)stage * FF_WSTAGE_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint64_t db = make_sw128_kmajor_desc(sw + kk * (FF_HC * 128));
#pragma unroll
                            for (int k = 0; k < 4; ++k)       // k range [64kk + 16k, +16) = packed columns 32kk + 8k
                                umma_f16_ts(tm_dh + g * FF_HC, tm_a + ab * 64 + kk * 32 + 8 * k, db + (uint64_t)(2 * k), idesc1, (kk | k) != 0);
                        }
                        umma_commit(&w_empty[stage]);
                        umma_commit(&d1_full[g]);
                        if (c == FF_CHUNKS - 1) umma_commit(ab ? at_empty2 : at_empty);   // last reader of this A[tmem]
                    }
                    __syncwarp();
                    if (++stage == WST) { stage = 0; phase ^= 1; }
                } else {
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    mbar_wait(&w_full[stage], phase);
                    tc_fence_after();
                    const uint32_t sw = smem_u32(sW + (size_t)stage * FF_WSTAGE_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint64_t db = make_sw128_kmajor_desc(sw + kk * (FF_HC * BK * 4));
#pragma unroll
                            for (int k = 0; k < BK / 8; ++k)
                                umma_tf32_ts(tm_dh + g * FF_HC, tm_a + (2 * half + kk) * BK + 8 * k, db + (uint64_t)(2 * k), idesc1,
                                             (half | kk | k) != 0);
                        }
                        if (CL == 1) umma_commit(&w_empty[stage]); else umma_commit_mc(&w_empty[stage], kMask);
                        if (half == 1) {
                            umma_commit(&d1_full[g]);
                            if (c == FF_CHUNKS - 1) umma_commit(at_empty);   // last reader of A[tmem]
                        }
                    }
                    __syncwarp();
                    if (++stage == WST) { stage = 0; phase ^= 1; }
                }
                }
                if (c == 1) { mbar_wait(&d2_empty[d], ((it >> 1) & 1) ^ 1); tc_fence_after(); }   // final epilogue two tiles back drained D2[d]
                if (c >= 1) gemm2(c - 1);
            }
            gemm2(FF_CHUNKS - 1);
            if (elect_one()) umma_commit(&d2_full[d]);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue-1 warps: all eight work on every
        // chunk (32 rows x 32 hidden units each), which halves the GEMM1 -> GEMM2 latency of a chunk
        const int q = warp & 3;                               // TMEM lane quarter
        const int hf = (warp - 2) >> 2;                       // which 32 of the chunk's 64 hidden units
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        uint32_t n_e1[2] = {0, 0};
        for (int64_t w = tile0; w < m_tiles; w += tile_step) {
            for (int c = 0; c < FF_CHUNKS; ++c) {
                const int g = c & 1;
                mbar_wait(&d1_full[g], n_e1[g] & 1); ++n_e1[g];
                tc_fence_after();
                float v[32];
                const uint32_t ta = tm_dh + lane_sel + g * FF_HC + hf * 32;
                tmem_ld_32x32(ta, v);
                const float *bb = svec + 3 * D_ + c * FF_HC + hf * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = *reinterpret_cast<const float4 *>(bb + 4 * j);
                    v[4 * j] = fmaxf(v[4 * j] + b.x, 0.f); v[4 * j + 1] = fmaxf(v[4 * j + 1] + b.y, 0.f);
                    v[4 * j + 2] = fmaxf(v[4 * j + 2] + b.z, 0.f); v[4 * j + 3] = fmaxf(v[4 * j + 3] + b.w, 0.f);
                    if (!F16) {                               // (the fp16 variant rounds when it packs)
                        v[4 * j] = tf32_rna(v[4 * j]); v[4 * j + 1] = tf32_rna(v[4 * j + 1]);
                        v[4 * j + 2] = tf32_rna(v[4 * j + 2]); v[4 * j + 3] = tf32_rna(v[4 * j + 3]);
                    }
                }
                if (F16) {                                    // relu output, packed: 16 columns at 32*hf (never the columns the
                    uint32_t pk[16];                          // other half's warp is still reading)
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_f16x2(v[2 * j], v[2 * j + 1]);
                    tmem_st_32x16(ta, pk);
                } else {
                    tmem_st_32x32(ta, v);                     // in place: D1 -> H
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_full[g]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                           // no CTA exits while a peer can still multicast into it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FF_TMEM_COLS);
    }
}

template <int CL, bool F16>
int launch_ff_fused_cl(const float *a_op, int round_a, const void *W1, const float *b1, const void *W2, const EpiParams &p,
                       cudaStream_t st) {
    CUtensorMap tmA, tmW1, tmW2;
    if (int rc = make_map(&tmA, a_op, (uint64_t)p.M, D_, BM)) return rc;
    if (F16) {
        if (int rc = make_map_f16(&tmW1, W1, HID_, D_, FF_HC)) return rc;                   // box [64 units x 64 k]
        if (int rc = make_map_f16(&tmW2, W2, D_, HID_, BM)) return rc;                      // box [128 outputs x 64 units]
    } else {
        if (int rc = make_map(&tmW1, static_cast<const float *>(W1), HID_, D_, CL == 1 ? FF_HC : 128 / CL)) return rc;     // box [rows x 32 k]
        if (int rc = make_map(&tmW2, static_cast<const float *>(W2), D_, HID_, CL == 1 ? BM : 128 / CL)) return rc;
    }
    auto kern = ff_fused_kernel<CL, F16>;
    GNNGLS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff_smem_bytes(F16)));
    const int64_t tiles = (p.M + BM - 1) / BM;
    const int sms = gnngls::device_sm_count();
    int64_t grid = tiles < sms ? tiles : sms;
    grid = (grid + CL - 1) / CL * CL;
    if (grid > sms) grid = sms / CL * CL;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(F16 ? FF_THREADS_F16 : FF_THREADS);
    cfg.dynamicSmemBytes = ff_smem_bytes(F16);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GNNGLS_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmW1, tmW2, p, b1, round_a));
    GNNGLS_LAUNCH_OK("ff_fused_kernel");
    return GNNGLS_OK;
}

int launch_ff_fused(const float *a_op, int round_a, const void *W1, const float *b1, const void *W2, const EpiParams &p,
                    cudaStream_t st, bool f16) {
    if (f16) return launch_ff_fused_cl<1, true>(a_op, round_a, W1, b1, W2, p, st);
    static int cl = -1;                                       // GNNGLS_FF_CLUSTER = 1 | 2 | 4 (default 1)
    if (cl < 0) {
        const char *e = getenv("GNNGLS_FF_CLUSTER");
        cl = e ? atoi(e) : 1;      // multicast measured neutral at 2, slower at 4 (profiles/r1_ff_cluster.md): default off
        if (cl != 1 && cl != 2 && cl != 4) cl = 1;
    }
    if (cl == 1) return launch_ff_fused_cl<1, false>(a_op, round_a, W1, b1, W2, p, st);
    if (cl == 4) return launch_ff_fused_cl<4, false>(a_op, round_a, W1, b1, W2, p, st);
    return launch_ff_fused_cl<2, false>(a_op, round_a, W1, b1, W2, p, st);
}

// ================================================================================================
// SIMT fp32 cross-check GEMM (debug): 64x64 tile, 256 threads, 4x4 micro-tile with the 4 columns
// contiguous so the shared epilogue_quad() applies unchanged.
// ================================================================================================
template <int N_TOTAL, int K_TOTAL, int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float *__restrict__ A, const float *__restrict__ W, const EpiParams p) {
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx: column quad, ty: row quad
    const int64_t row0 = (int64_t)blockIdx.x * 64;
    const int col0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K_TOTAL; k0 += 16) {
        for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
            const int r = idx >> 4, k = idx & 15;
            const int64_t gr = row0 + r;
            As[k][r] = gr < p.M ? A[gr * K_TOTAL + k0 + k] : 0.f;
            Ws[k][r] = W[(int64_t)(col0 + r) * K_TOTAL + k0 + k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    // lanes of a warp: tx = lane & 15 -> 16 consecutive quads of one row => groups of 4 lanes share a head
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = row0 + ty * 4 + i;
        epilogue_quad<EPI>(p, row, col0 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]), row < p.M);
    }
}

template <int N_TOTAL, int K_TOTAL, int EPI>
int launch_simt_gemm(const float *A, const float *W, const EpiParams &p, cudaStream_t st) {
    dim3 grid((unsigned)((p.M + 63) / 64), N_TOTAL / 64);
    gemm_simt_kernel<N_TOTAL, K_TOTAL, EPI><<<grid, 256, 0, st>>>(A, W, p);
    GNNGLS_LAUNCH_OK("gemm_simt_kernel");
    return GNNGLS_OK;
}

// ================================================================================================
// embed_layer / decision_layer
// ================================================================================================
__global__ void embed_kernel(const float *__restrict__ x, int64_t M, int in_dim, const float *__restrict__ W,
                             const float *__restrict__ b, float *__restrict__ h, float *__restrict__ h_tf32, int op_f16) {
    // one warp per node, lane owns 4 output channels
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp0; m < M; m += nwarps) {
        float4 o = *reinterpret_cast<const float4 *>(b + 4 * lane);
        for (int k = 0; k < in_dim; ++k) {
            const float xv = x[m * in_dim + k];
            o.x = fmaf(xv, W[(4 * lane + 0) * in_dim + k], o.x);
            o.y = fmaf(xv, W[(4 * lane + 1) * in_dim + k], o.y);
            o.z = fmaf(xv, W[(4 * lane + 2) * in_dim + k], o.z);
            o.w = fmaf(xv, W[(4 * lane + 3) * in_dim + k], o.w);
        }
        *reinterpret_cast<float4 *>(h + m * D_ + 4 * lane) = o;
        if (h_tf32) store_op_copy(h_tf32, op_f16, m * D_ + 4 * lane, o);
    }
}

__global__ void decision_kernel(const float *__restrict__ h, int64_t M, int out_dim, const float *__restrict__ Wd,
                                const float *__restrict__ bd, float *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp0; m < M; m += nwarps) {
        const float4 v = *reinterpret_cast<const float4 *>(h + m * D_ + 4 * lane);
        for (int o = 0; o < out_dim; ++o) {
            const float4 w = *reinterpret_cast<const float4 *>(Wd + o * D_ + 4 * lane);
            float s = v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) y[m * out_dim + o] = s + bd[o];
        }
    }
}

int elementwise_grid(int64_t warps_needed, int threads) {
    const int64_t blocks = (warps_needed * 32 + threads - 1) / threads;
    const int64_t cap = (int64_t)gnngls::device_sm_count() * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

extern "C" int gnngls_embed_forward(const float *x, int64_t M, int in_dim, const float *W, const float *b, float *h,
                                    void *h_op, int op_dtype, void *stream) {
    GNNGLS_REQUIRE(x && W && b && h, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(!h_op || op_dtype == GNNGLS_FT_TF32 || op_dtype == GNNGLS_FT_F16, GNNGLS_ERR_BAD_ARG, "operand copy must be TF32 or fp16");
    float *h_tf32 = static_cast<float *>(h_op);
    const int op_f16 = op_dtype == GNNGLS_FT_F16;
    GNNGLS_REQUIRE(in_dim >= 1, GNNGLS_ERR_BAD_ARG, "in_dim must be >= 1");
    if (M <= 0) return GNNGLS_OK;
    embed_kernel<<<elementwise_grid(M, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, M, in_dim, W, b, h, h_tf32, op_f16);
    GNNGLS_LAUNCH_OK("embed_kernel");
    return GNNGLS_OK;
}

extern "C" int gnngls_decision_forward(const float *h, int64_t M, int out_dim, const float *Wd, const float *bd,
                                       float *y, void *stream) {
    GNNGLS_REQUIRE(h && Wd && bd && y, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(out_dim >= 1, GNNGLS_ERR_BAD_ARG, "out_dim must be >= 1");
    if (M <= 0) return GNNGLS_OK;
    decision_kernel<<<elementwise_grid(M, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(h, M, out_dim, Wd, bd, y);
    GNNGLS_LAUNCH_OK("decision_kernel");
    return GNNGLS_OK;
}

extern "C" int gnngls_fc_forward(int impl, const void *h, int64_t M, const void *Wfc, const float *attn_l,
                                 const float *attn_r, void *ft, int ft_dtype, float *el, float *er, void *stream) {
    GNNGLS_REQUIRE(h && Wfc && attn_l && attn_r && ft && el && er, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(ft_dtype == GNNGLS_FT_F32 || ft_dtype == GNNGLS_FT_TF32 || ft_dtype == GNNGLS_FT_F16, GNNGLS_ERR_BAD_ARG,
                   "unknown ft_dtype %d", ft_dtype);
    if (M <= 0) return GNNGLS_OK;
    EpiParams p{};
    p.M = M; p.el = el; p.er = er; p.v0 = attn_l; p.v1 = attn_r;
    if (ft_dtype == GNNGLS_FT_F16) p.out_f16 = static_cast<__half *>(ft);
    else p.out = static_cast<float *>(ft);
    p.round_tf32 = ft_dtype == GNNGLS_FT_TF32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (impl == GNNGLS_DENSE_TCGEN05) return launch_tc_gemm<D_, D_, EPI_FC>(h, Wfc, p, st);
    if (impl == GNNGLS_DENSE_TCGEN05_F16) return launch_tc_gemm<D_, D_, EPI_FC, true>(h, Wfc, p, st);      // h and Wfc are fp16
    if (impl == GNNGLS_DENSE_SIMT) return launch_simt_gemm<D_, D_, EPI_FC>(static_cast<const float *>(h), static_cast<const float *>(Wfc), p, st);
    GNNGLS_REQUIRE(false, GNNGLS_ERR_BAD_ARG, "unknown dense impl %d", impl);
}

extern "C" size_t gnngls_ff_workspace_bytes(int impl, int64_t M) {
    if (impl == GNNGLS_DENSE_TCGEN05 || impl == GNNGLS_DENSE_TCGEN05_F16) return 0;   // fused: the hidden activations stay on chip
    return M > 0 ? (size_t)M * HID_ * sizeof(float) : 0;     // debug path materialises hidden [M,512]
}

extern "C" int gnngls_ff_forward(int impl, const float *h1, const float *h1_tf32, int64_t M, const void *W1,
                                 const float *b1, const void *W2, const float *b2, const float *bn_scale,
                                 const float *bn_shift, float *h_out, void *h_out_op, int op_dtype, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    GNNGLS_REQUIRE(h1 && W1 && b1 && W2 && b2 && bn_scale && bn_shift && h_out, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(!h_out_op || op_dtype == GNNGLS_FT_TF32 || op_dtype == GNNGLS_FT_F16, GNNGLS_ERR_BAD_ARG, "operand copy must be TF32 or fp16");
    float *h_out_tf32 = static_cast<float *>(h_out_op);
    if (M <= 0) return GNNGLS_OK;
    const size_t need = gnngls_ff_workspace_bytes(impl, M);
    GNNGLS_REQUIRE(need == 0 || (workspace && workspace_bytes >= need), GNNGLS_ERR_WORKSPACE,
                   "ff workspace too small: need %zu bytes", need);
    float *hid = static_cast<float *>(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EpiParams p1{};
    p1.M = M; p1.out = hid; p1.v0 = b1;
    EpiParams p2{};
    p2.M = M; p2.out = h_out; p2.out_tf32 = h_out_tf32; p2.out_op_f16 = op_dtype == GNNGLS_FT_F16; p2.v0 = b2; p2.v1 = bn_scale; p2.v2 = bn_shift; p2.skip = h1;
    const float *a1 = h1_tf32 ? h1_tf32 : h1;                 // GEMM operand; the skip path always reads fp32 h1
    if (impl == GNNGLS_DENSE_TCGEN05)
        return launch_ff_fused(a1, h1_tf32 == nullptr, W1, b1, W2, p2, st, false);   // no pre-rounded copy: round while staging
    if (impl == GNNGLS_DENSE_TCGEN05_F16)
    {
        p2.skip_is_a = 1;
        return launch_ff_fused(h1, 0, W1, b1, W2, p2, st, true);                     // fp16 weights; h1 is packed to fp16 while staging
    }
    if (impl == GNNGLS_DENSE_SIMT) {
        if (int rc = launch_simt_gemm<HID_, D_, EPI_FF1>(a1, static_cast<const float *>(W1), p1, st)) return rc;
        return launch_simt_gemm<D_, HID_, EPI_FF2>(hid, static_cast<const float *>(W2), p2, st);
    }
    GNNGLS_REQUIRE(false, GNNGLS_ERR_BAD_ARG, "unknown dense impl %d", impl);
}
