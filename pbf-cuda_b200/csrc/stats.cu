// stats.cu — run statistics of SURVEY.md A.9 (density error, kinetic energy, max speed, mean z).
// The reference has no diagnostics (SURVEY.md section 5 "Metrics / logging: printf only"); the
// 1000-step comparison BASELINE.json asks for needs them. Deterministic: fixed block count, fixed
// tree order, f64 accumulation; the host adds the per-block partials in index order.
#include "pbf_internal.h"

namespace pbf {

constexpr int ST_THREADS = 256;

// partial[b*5 + k]: k = 0 sum|rho/rho0-1|, 1 max(rho/rho0-1), 2 sum 0.5|v|^2, 3 max |v|^2, 4 sum z
__global__ void __launch_bounds__(ST_THREADS)
stats_kernel(const float* __restrict__ rho, const float* __restrict__ npos, const float* __restrict__ nvel,
             int64_t n, float pho0, double* __restrict__ partial) {
    __shared__ double s[5][ST_THREADS];
    double e_sum = 0, e_max = -1e300, ke = 0, v_max = 0, z_sum = 0;
    for (int64_t i = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * ST_THREADS) {
        const double cdev = (double)rho[i] / (double)pho0 - 1.0;
        e_sum += fabs(cdev);
        e_max = fmax(e_max, cdev);
        const double vx = nvel[3 * i], vy = nvel[3 * i + 1], vz = nvel[3 * i + 2];
        const double v2 = vx * vx + vy * vy + vz * vz;
        ke += 0.5 * v2;
        v_max = fmax(v_max, v2);
        z_sum += (double)npos[3 * i + 2];
    }
    s[0][threadIdx.x] = e_sum; s[1][threadIdx.x] = e_max; s[2][threadIdx.x] = ke;
    s[3][threadIdx.x] = v_max; s[4][threadIdx.x] = z_sum;
    __syncthreads();
    for (int off = ST_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s[0][threadIdx.x] += s[0][threadIdx.x + off];
            s[1][threadIdx.x] = fmax(s[1][threadIdx.x], s[1][threadIdx.x + off]);
            s[2][threadIdx.x] += s[2][threadIdx.x + off];
            s[3][threadIdx.x] = fmax(s[3][threadIdx.x], s[3][threadIdx.x + off]);
            s[4][threadIdx.x] += s[4][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 5; k++) partial[blockIdx.x * 5 + k] = s[k][0];
}

cudaError_t preload_stats() {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, stats_kernel);
}

cudaError_t launch_stats(const float* rho, const float* npos, const float* nvel, int64_t n, float pho0,
                         double* partial, int nblocks, cudaStream_t st) {
    stats_kernel<<<nblocks, ST_THREADS, 0, st>>>(rho, npos, nvel, n, pho0, partial);
    return cudaGetLastError();
}

}  // namespace pbf
