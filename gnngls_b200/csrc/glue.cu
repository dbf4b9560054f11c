// Glue between the model and the search: input features and regret post-processing.
// Compiled with -fmad=false: each step below is one IEEE operation, rounded as the reference's
// numpy/sklearn code rounds it.
#include <cstdint>
#include "common.h"

namespace {

// datasets.py:14-20 (feature = float32(weight)) then MinMaxScaler.transform (datasets.py:85):
// sklearn runs `X *= scale_; X += min_` in place on the float32 array with float64 scalars, i.e.
// each step is computed in fp64 and rounded to fp32.
__global__ void edge_features_kernel(const double *__restrict__ D, int B, int n, double scale, double min_,
                                     float *__restrict__ x) {
    const int64_t N = (int64_t)n * (n - 1) / 2;
    const int64_t total = (int64_t)B * n * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = idx / ((int64_t)n * n);
        const int rem = (int)(idx - b * n * n);
        const int i = rem / n, j = rem - i * n;
        if (j <= i) continue;
        const int64_t v = (int64_t)i * (2 * n - i - 1) / 2 + (j - i - 1);   // line-graph node id of (i,j)
        const float x0 = __double2float_rn(D[idx]);
        const float x1 = __double2float_rn(__dmul_rn((double)x0, scale));
        x[b * N + v] = __double2float_rn(__dadd_rn((double)x1, min_));
    }
}

// scripts/test.py:79-83: MinMaxScaler.inverse_transform (`X -= min_; X /= scale_` on float32 with
// float64 scalars) followed by np.maximum(., 0)
__global__ void regret_post_kernel(const float *__restrict__ y, int64_t M, double scale, double min_,
                                   float *__restrict__ r) {
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
        const float r1 = __double2float_rn(__dsub_rn((double)y[m], min_));
        const float r2 = __double2float_rn(__ddiv_rn((double)r1, scale));
        r[m] = (r2 < 0.f) ? 0.f : r2;
    }
}

int grid_1d(int64_t total, int threads) {
    const int64_t blocks = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)gnngls::device_sm_count() * 32;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

extern "C" int gnngls_edge_features(const double *D, int B, int n, double scale, double min_, float *x, void *stream) {
    GNNGLS_REQUIRE(D && x, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(n >= 2 && n <= 32768, GNNGLS_ERR_UNSUPPORTED, "n=%d unsupported", n);
    if (B <= 0) return GNNGLS_OK;
    edge_features_kernel<<<grid_1d((int64_t)B * n * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(D, B, n, scale,
                                                                                                      min_, x);
    GNNGLS_LAUNCH_OK("edge_features_kernel");
    return GNNGLS_OK;
}

extern "C" int gnngls_regret_postprocess(const float *y, int64_t M, double scale, double min_, float *regret,
                                         void *stream) {
    GNNGLS_REQUIRE(y && regret, GNNGLS_ERR_BAD_ARG, "null pointer argument");
    GNNGLS_REQUIRE(scale != 0.0, GNNGLS_ERR_BAD_ARG, "scale must be non-zero");
    if (M <= 0) return GNNGLS_OK;
    regret_post_kernel<<<grid_1d(M, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, M, scale, min_, regret);
    GNNGLS_LAUNCH_OK("regret_post_kernel");
    return GNNGLS_OK;
}
