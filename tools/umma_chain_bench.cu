// Latency of one operand-in-TMEM tcgen05.mma chain as gat_kn_tc.cu issues it (M=128, N=48, K = 16 x nk, A in tensor
// memory, B in shared memory MN-major without swizzle), and of T such chains issued back to back by T different
// threads of one CTA (4 teams issue theirs at about the same time).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_chain_bench tools/umma_chain_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N>
__global__ void __launch_bounds__(512, 1) k(long long *out, int nk, int teams, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, team = tid >> 7;
    constexpr int KB = (N / 8) * 128;
    for (int i = tid; i < 4 * 16 * KB / 4; i += 512) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        for (int t = 0; t < 4; ++t) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[t])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tslot + team * 128;
    const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long total = 0;
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        const long long t0 = clock64();
        if ((tid & 127) == 0 && team < teams) {
            for (int ks = 0; ks < nk; ++ks) {
                const uint32_t addr = smem_u32(smem) + team * 16 * KB + ks * 2 * KB;
                const uint64_t desc = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(KB >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
                const uint32_t acc = ks != 0;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                             ::"r"(tbase + 64), "r"(tbase + ks * 8), "l"(desc), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[team])) : "memory");
        }
        __syncwarp();
        if (team < teams) {
            asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra.uni DN;\nbra.uni W;\nDN:\n}\n" ::"r"(smem_u32(&bar[team])), "r"(parity) : "memory");
        }
        parity ^= 1;
        const long long t1 = clock64();
        if (tid == 128 * (teams - 1)) total += t1 - t0;
    }
    if (tid == 128 * (teams - 1)) out[blockIdx.x] = total / reps;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tslot) : "memory");
}

template <int N>
void run(int nk, int teams) {
    long long *d, h[148];
    cudaMalloc(&d, 148 * 8);
    const int smem = 4 * 16 * (N / 8) * 128;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N><<<148, 512, smem>>>(d, nk, teams, 200);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
    printf("N=%3d nk=%d teams=%d: %7.0f cycles from first issue to the last team's completion  (%s)\n", N, nk, teams, m, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int teams : {1, 2, 4}) { run<48>(7, teams); run<48>(8, teams); run<64>(7, teams); run<32>(7, teams); run<128>(7, teams); }
    run<48>(1, 1); run<48>(2, 1); run<48>(4, 1);
    return 0;
}
