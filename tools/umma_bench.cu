// Micro-benchmark: issue rate / throughput of tcgen05.mma kind::tf32 (A,B in 128B-swizzled smem)
// for several N, measured with clock64 around a commit + mbarrier wait.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, bool BF16, bool TMEM_A = false>
__global__ void __launch_bounds__(128, 1) bench(long long *out, int iters, int kblocks) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint32_t fmt = BF16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
        uint32_t parity = 0;
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            t0 = clock64();
            uint32_t lead;
            asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(lead));
            if (lead) {
                for (int it = 0; it < iters; ++it) {
                    for (int kb = 0; kb < kblocks; ++kb) {      // kb-th 128-byte K block: A box kb (16 KB), B box kb
                        const uint64_t da = desc(a0 + (kb & 3) * 16384), db = desc(b0 + (kb & 3) * (N * 128));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (TMEM_A) mma_ts(tm, tm + 256 + (kb & 3) * 32 + 8 * k, db + 2 * k, idesc, (it | kb | k) != 0);
                            else if (BF16) mma_bf16(tm, da + 2 * k, db + 2 * k, idesc, (it | kb | k) != 0);
                            else mma(tm, da + 2 * k, db + 2 * k, idesc, (it | kb | k) != 0);
                        }
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra.uni D;\nbra.uni W;\nD:\n}\n"
                         ::"r"(smem_u32(&bar)), "r"(parity) : "memory");
            parity ^= 1;
            t1 = clock64();
        }
        if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = (long long)iters * kblocks * 4; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

template <int N, bool BF16, bool TMEM_A = false>
void run(const char *name, int grid) {
    long long *d;
    cudaMalloc(&d, sizeof(long long) * 2 * grid);
    auto k = bench<N, BF16, TMEM_A>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<grid, 128, 200 * 1024>>>(d, 64, 4);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-12s N=%3d grid=%3d: %s  cycles=%lld  mmas=%lld  cycles/mma=%.1f  MAC/clk/SM=%.0f\n", name, N, grid,
           cudaGetErrorString(e), h[0], h[1], (double)h[0] / h[1], 128.0 * N * (BF16 ? 16 : 8) * h[1] / h[0]);
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<32, false>("tf32", grid); run<64, false>("tf32", grid); run<128, false>("tf32", grid); run<256, false>("tf32", grid);
        run<32, true>("bf16", grid); run<128, true>("bf16", grid); run<256, true>("bf16", grid);
        run<32, false, true>("tf32 A=tmem", grid); run<64, false, true>("tf32 A=tmem", grid); run<128, false, true>("tf32 A=tmem", grid);
    }
    return 0;
}
