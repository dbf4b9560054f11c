// Stand-alone timing of the fp16 star kernel's main loop (the real device code, synthetic shared-memory contents).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I gnngls_b200/csrc -o tools/star_loop_bench tools/star_loop_bench.cu gnngls_b200/csrc/common.cu
#include "../gnngls_b200/csrc/gat.cu"
#include <cstdio>

template <int NTMAX, int MINB>
__global__ void __launch_bounds__(STAR_THREADS, MINB) loop_kernel(int n, int reps, long long *cyc, float *part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int KP = round_up(n, 8), KE = round_up(n, 16);
    __half *Fh = reinterpret_cast<__half *>(smem_raw);
    float *ELt = reinterpret_cast<float *>(Fh + (size_t)KP * FH_LD);
    float *ERs = ELt + (size_t)H_ * KE;
    float2 *EA = reinterpret_cast<float2 *>(ERs + (size_t)KE * H_);
    for (int i = threadIdx.x; i < KP * FH_LD; i += STAR_THREADS) Fh[i] = __float2half((float)((i * 31) % 17) * 0.1f - 0.8f);
    for (int i = threadIdx.x; i < H_ * KE; i += STAR_THREADS) ELt[i] = (i % KE) < n ? (float)((i * 13) % 23) * 0.2f - 2.f : -INFINITY;
    for (int i = threadIdx.x; i < KE * H_; i += STAR_THREADS) ERs[i] = (float)((i * 7) % 19) * 0.2f - 2.f;
    for (int i = threadIdx.x; i < H_ * KE; i += STAR_THREADS) EA[i] = make_float2(exp2f(ELt[i] - 2.4f), exp2f(0.2f * (ELt[i] - 2.4f)));
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Star16Ctx c;
    c.base.n = n; c.base.i = 0; c.base.b = 0; c.base.hd = warp; c.base.g = lane >> 2; c.base.t = lane & 3;
    c.base.m1 = 2.4f; c.base.m2 = 2.2f; c.base.a1 = 5;
    c.base.Fh = nullptr; c.base.ELs = nullptr; c.base.ERs = ERs; c.base.ksteps = 0;
    c.base.part = part + (size_t)blockIdx.x * n * (n - 1) * REC;
    c.ELt = ELt + warp * KE;
    c.EA = EA + warp * KE;
    c.base.skip_row = 5;
    const uint32_t fh = (uint32_t)__cvta_generic_to_shared(Fh);
    const int q = lane >> 3, r = lane & 7;
    c.b4_addr = fh + (uint32_t)(((r + 8 * (q & 1)) * FH_LD + warp * 16 + 8 * (q >> 1)) * 2);
    c.kfull = KP / 16;
    c.tail = (KP & 8) != 0;
    c.b2_addr = fh + (uint32_t)(((c.kfull * 16 + r) * FH_LD + warp * 16 + 8 * (q & 1)) * 2);
    const int MT = KE / 16;
    __syncthreads();
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        int mt = 0;
        if (NTMAX >= 2)
            for (; mt + 2 <= MT; mt += 2) star16_tiles<2>(c, mt);
        for (; mt < MT; ++mt) star16_tiles<1>(c, mt);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NTMAX, int MINB>
void run(int n, int ctas_per_sm) {
    const int reps = 8, grid = 148 * ctas_per_sm;
    long long *cyc; float *pn;
    cudaMalloc(&cyc, grid * 8);
    cudaMalloc(&pn, (size_t)grid * n * (n - 1) * REC * 4);
    const size_t smem = star16_smem_bytes(n);
    cudaFuncSetAttribute(loop_kernel<NTMAX, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int w = 0; w < 2; ++w) loop_kernel<NTMAX, MINB><<<grid, STAR_THREADS, smem>>>(n, reps, cyc, pn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long *h = new long long[grid];
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    const double tile_steps = 7 * 6.5;     // n = 100
    printf("n=%d NTMAX=%d minb=%d ctas/SM=%d: %8.0f cycles per main loop  (%.0f per CTA-throughput; %.1f cycles per tile-step per SMSP-slot)\n", n, NTMAX, MINB,
           ctas_per_sm, mean / reps, mean / reps / ctas_per_sm, mean / reps / ctas_per_sm / (tile_steps * 2));
    cudaFree(cyc); cudaFree(pn); delete[] h;
}

int main() {
    run<2, 3>(100, 1); run<2, 3>(100, 2); run<2, 3>(100, 3);
    return 0;
}
