// Correctness probe: tcgen05.mma kind::f16 with A in tensor memory (packed fp16 pairs, lane = row) and B in shared
// memory in the MN-major, no-swizzle canonical layout  [k/8][n/8][k%8][n%8]  (core matrix = 8 k-rows x 16 bytes).
// Which of the two descriptor strides is the K-block stride is tried both ways.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_mn_test tools/umma_mn_test.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int M = 128, N = 48, K = 128;

__global__ void __launch_bounds__(128, 1) probe(const __half *A, const __half *B, float *D, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tslot;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    // B -> shared memory, element (k, n) at  (k/8)*KB + (n/8)*128 + (k%8)*16 + (n%8)*2
    constexpr int KB = (N / 8) * 128;
    for (int idx = tid; idx < K * N; idx += 128) {
        const int k = idx / N, n = idx % N;
        *reinterpret_cast<__half *>(smem + (k / 8) * KB + (n / 8) * 128 + (k % 8) * 16 + (n % 8) * 2) = B[k * N + n];
    }
    // A row `tid` -> tensor memory columns [64, 64 + K/2): one 32-bit column per pair of k
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
        uint32_t v[16];
        for (int u = 0; u < 16; ++u) {
            const __half2 h2 = __halves2half2(A[tid * K + 2 * (c0 + u)], A[tid * K + 2 * (c0 + u) + 1]);
            v[u] = *reinterpret_cast<const uint32_t *>(&h2);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(tbase + lane_sel + 64 + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                       "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        // instruction descriptor: F32 accumulate, F16 x F16, A K-major (tensor memory), B MN-major (bit 16)
        const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t kstride = KB, nstride = 128;
        const uint32_t lbo = mode == 0 ? kstride : nstride, sbo = mode == 0 ? nstride : kstride;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t addr = smem_u32(smem) + ks * 2 * KB;
            const uint64_t desc = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
            const uint32_t acc = ks != 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                         ::"r"(tbase), "r"(tbase + 64 + ks * 8), "l"(desc), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra.uni DN;\nbra.uni W;\nDN:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tbase + lane_sel + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int u = 0; u < 16; ++u) D[tid * N + c0 + u] = __uint_as_float(r[u]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}

int main() {
    std::vector<__half> A(M * K), B(K * N);
    srand(1);
    for (auto &x : A) x = __float2half((float)(rand() % 2));                  // 0/1 indicator
    for (auto &x : B) x = __float2half((float)(rand() % 17 - 8) / 8.f);       // exactly representable
    std::vector<float> ref(M * N, 0.f);
    for (int i = 0; i < M; ++i)
        for (int k = 0; k < K; ++k)
            for (int n = 0; n < N; ++n) ref[i * N + n] += __half2float(A[i * K + k]) * __half2float(B[k * N + n]);
    __half *dA, *dB;
    float *dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    const int smem = K * N * 2;
    int ok_mode = -1;
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dD, 0xFF, M * N * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> D(M * N);
        cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        int bad = 0;
        for (int i = 0; i < M * N; ++i) { const double d = fabs((double)D[i] - ref[i]); if (!(d <= 1e-3)) ++bad; if (d > err) err = d; }
        printf("mode %d (%s): cuda=%s max err %.4g bad %d / %d   D[0..3]=%g %g %g %g ref=%g %g %g %g\n", mode,
               mode == 0 ? "LBO=K-block stride, SBO=N-block stride" : "LBO=N-block stride, SBO=K-block stride", cudaGetErrorString(e), err, bad,
               M * N, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
        if (bad == 0) ok_mode = mode;
    }
    printf("RESULT ok_mode=%d\n", ok_mode);
    return ok_mode >= 0 ? 0 : 1;
}
