// Issue-rate microbenchmark for the instruction classes of the star kernel's main loop on sm_100a:
// per-SM throughput (thread-ops / clock) of MUFU.EX2, FMNMX, FADD/FFMA, F2FP.PACK and HMMA.16816
// with 8 and 32 resident warps.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench tools/pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, U = 8;

template <int OP>
__global__ void k(float *out, long long *cyc, float seed) {
    float x[U];
    uint32_t p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { x[u] = seed + threadIdx.x * 1e-3f + u; p[u] = threadIdx.x + u; }
    float c[4][4] = {};
    unsigned long long q2[U], c2;                       // packed fp32 pairs
#pragma unroll
    for (int u = 0; u < U; ++u) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q2[u]) : "f"(x[u]), "f"(x[u] + 1.f));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(seed));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[u]));
            if (OP == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[u]) : "f"(seed));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[u]) : "f"(seed));
            if (OP == 3) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p[u]) : "f"(x[u]), "f"(x[(u + 1) % U]));
            if (OP == 4)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[u & 3][0]), "+f"(c[u & 3][1]), "+f"(c[u & 3][2]), "+f"(c[u & 3][3])
                             : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]));
            if (OP == 5)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[u & 3][0]), "+f"(c[u & 3][1]), "+f"(c[u & 3][2]), "+f"(c[u & 3][3])
                             : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]));
            if (OP == 6) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[u]) : "f"(seed));
            if (OP == 11) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p[u]));
            if (OP == 12) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p[u]));
            if (OP == 13) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(p[u]));
            if (OP == 14) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(p[u]) : "r"(p[(u + 1) % U]));
            if (OP == 15) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(p[u]) : "r"(p[(u + 1) % U]));
            if (OP == 18) asm volatile("set.ge.f16x2.f16x2 %0, %0, %1;" : "+r"(p[u]) : "r"(p[(u + 1) % U]));
            if (OP == 19) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(p[u]) : "r"(p[(u + 1) % U]));
            if (OP == 20) asm volatile("set.ge.f32.f32 %0, %0, %1;" : "+f"(x[u]) : "f"(seed));
            if (OP == 21) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(p[u]) : "r"(p[(u + 1) % U]), "r"(p[(u + 2) % U]));
            if (OP == 22) asm volatile("sub.u32 %0, %0, %1;" : "+r"(p[u]) : "r"(p[(u + 1) % U]));
            if (OP == 16) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q2[u]) : "l"(c2));
            if (OP == 17) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q2[u]) : "l"(c2));
        }
        if (OP >= 7 && OP <= 10) {
            // one tile-step of the star main loop: 8 weights = max(el+c1, 0.2*el+c2) -> ex2 -> pack -> 3 HMMA
            float w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float e = x[u] * 1e-3f;
                float a, b2;
                asm volatile("add.f32 %0, %1, %2;" : "=f"(a) : "f"(e), "f"(seed));
                asm volatile("fma.rn.f32 %0, %1, %2, %2;" : "=f"(b2) : "f"(e), "f"(seed));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a) : "f"(b2));
                if (OP != 9) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(w[u]) : "f"(a));
                else w[u] = a;
            }
            uint32_t q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q[u]) : "f"(w[2 * u]), "f"(w[2 * u + 1]));
            const int nm = OP == 8 ? 0 : (OP == 10 ? 2 : 3);
#pragma unroll
            for (int m = 0; m < 3; ++m)
                if (m < nm)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c[m][0]), "+f"(c[m][1]), "+f"(c[m][2]), "+f"(c[m][3])
                                 : "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(p[4]), "r"(p[5]));
            if (OP == 8) { p[0] ^= q[0] ^ q[1] ^ q[2] ^ q[3]; }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) s += x[u] + __uint_as_float(p[u]) + (float)(q2[u] & 0xffff);
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) s += c[a][b];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int warps) {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, warps * 32>>>(out, cyc, 0.5f);
    k<OP><<<148, warps * 32>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
    const double ops = (double)ITERS * U * warps * 32;
    printf("%-18s warps=%2d  %8.0f cycles  %7.2f thread-ops/clk/SM  (%.2f clk per warp-instr per SMSP)\n", name, warps, mean, ops / mean,
           mean / ((double)ITERS * U * warps / 4));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {8, 32}) {
        run<0>("MUFU.EX2", w); run<1>("FMNMX", w); run<2>("FFMA", w); run<6>("FADD", w); run<3>("F2FP.PACK_AB", w);
        run<4>("HMMA.16816.F32", w); run<5>("HMMA.1688.TF32", w);
        run<12>("EX2.f16x2", w); run<13>("TANH.f16x2", w); run<14>("HMNMX2", w); run<15>("HFMA2", w); run<16>("FMUL2 (f32x2)", w); run<17>("FFMA2 (f32x2)", w);
        run<18>("HSET2.BF.GE", w); run<19>("HMUL2", w); run<20>("FSET.BF.GE", w); run<21>("LOP3", w); run<22>("IADD", w);
    }
    return 0;
}
