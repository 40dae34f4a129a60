// stats.cu — run statistics of SURVEY.md A.9 (density error, kinetic energy, max speed, mean z).
// The reference has no diagnostics (SURVEY.md section 5 "Metrics / logging: printf only"); the
// 1000-step comparison BASELINE.json asks for needs them. Deterministic: fixed block count, fixed
// tree order, f64 accumulation; the host adds the per-block partials in index order.
#include <string.h>

#include "pbf_math.cuh"

namespace pbf {

constexpr int ST_THREADS = 256;

// splitmix64's finaliser; particle_hash chains it over (iid, the six state words as three 64-bit words)
__host__ __device__ inline uint64_t digest_mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t particle_hash(uint32_t iid, uint32_t px, uint32_t py, uint32_t pz, uint32_t vx,
                                                  uint32_t vy, uint32_t vz) {
    uint64_t h = digest_mix64((uint64_t)iid);
    h = digest_mix64(h ^ ((uint64_t)px | ((uint64_t)py << 32)));
    h = digest_mix64(h ^ ((uint64_t)pz | ((uint64_t)vx << 32)));
    return digest_mix64(h ^ ((uint64_t)vy | ((uint64_t)vz << 32)));
}

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

// ---- exhaustive verification of the constant-divisor sequence (pbf_math.cuh div_pho0) ------------------
// Every float bit pattern a: does fma(fma(-a*y, d, a), y, a*y) equal a / d (div.rn) bit for bit? Mismatches are
// reduced to: the largest |a| < 1 that fails and the smallest |a| >= 1 that fails (as positive-float bit
// patterns, which order like the values). The verified interval is what lies strictly between them.
__global__ void __launch_bounds__(256) const_div_check_kernel(float d, float y, uint32_t* fail_below, uint32_t* fail_above) {
    uint32_t lo = 0, hi = 0xffffffffu;
    for (uint64_t b = (uint64_t)blockIdx.x * 256 + threadIdx.x; b < (1ull << 32); b += (uint64_t)gridDim.x * 256) {
        const float a = __uint_as_float((uint32_t)b);
        const float q = __fmul_rn(a, y);
        const float fast = __fmaf_rn(__fmaf_rn(-q, d, a), y, q);
        const float exact = __fdiv_rn(a, d);
        if (__float_as_uint(fast) != __float_as_uint(exact)) {
            const uint32_t m = (uint32_t)b & 0x7fffffffu;
            if (m < 0x3f800000u) lo = max(lo, m); else hi = min(hi, m);
        }
    }
    if (lo) atomicMax(fail_below, lo);
    if (hi != 0xffffffffu) atomicMin(fail_above, hi);
}

// (`scratch`: 8 bytes of device memory owned by the caller — no allocation, hence no device-wide synchronisation,
//  on this path; the kernels run on `st` and only `st` is waited for)
cudaError_t verify_const_div(float d, float rcp, float* lo, float* hi, void* scratch, cudaStream_t st) {
    uint32_t* dev = (uint32_t*)scratch;
    cudaError_t e = cudaSuccess;
    const uint32_t init[2] = {0u, 0xffffffffu};
    uint32_t out[2] = {0, 0};
    e = cudaMemcpyAsync(dev, init, 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        const_div_check_kernel<<<148 * 16, 256, 0, st>>>(d, rcp, dev, dev + 1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dev, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    // strictly inside the failures; +0 (pattern 0) "fails below" only if 0/d mismatched, which it does not,
    // but zero is excluded anyway: lo is at least the smallest normal number
    uint32_t lo_bits = out[0] + 1, hi_bits = out[1] == 0xffffffffu ? 0x7f7fffffu : out[1] - 1;
    if (lo_bits < 0x00800000u) lo_bits = 0x00800000u;
    if (hi_bits > 0x7f7fffffu) hi_bits = 0x7f7fffffu;
    memcpy(lo, &lo_bits, 4);
    memcpy(hi, &hi_bits, 4);
    return cudaSuccess;
}

// ---- exhaustive verification of the branch-free spiky scale (pbf_math.cuh spiky_scale_fast) -------------
// Every float r2 from +0 up to `top` (bit patterns 0 .. bits(top), which order like the values): does the fast
// sequence give the bits of the exact one? Counts the mismatches; the caller uses the fast sequence only if
// there are none.
__global__ void __launch_bounds__(256) spiky_check_kernel(uint32_t top_bits, SolverConsts c, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (uint64_t b = (uint64_t)blockIdx.x * 256 + threadIdx.x; b <= top_bits; b += (uint64_t)gridDim.x * 256) {
        const float r2 = __uint_as_float((uint32_t)b);
        if (__float_as_uint(spiky_scale_fast(r2, c)) != __float_as_uint(spiky_scale(r2, c))) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t verify_spiky(const SolverConsts& c, float top, unsigned long long* mismatches, void* scratch, cudaStream_t st) {
    unsigned long long* dev = (unsigned long long*)scratch;
    cudaError_t e = cudaSuccess;
    uint32_t top_bits;
    memcpy(&top_bits, &top, 4);
    e = cudaMemsetAsync(dev, 0, 8, st);
    if (e == cudaSuccess) {
        spiky_check_kernel<<<148 * 16, 256, 0, st>>>(top_bits, c, dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(mismatches, dev, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
}

// ---- the same for the trimmed powf(w, 4.0f) of the delta-p pass (pbf_math.cuh pow4_trim) ------------------
__global__ void __launch_bounds__(256) pow4_check_kernel(uint32_t top_bits, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (uint64_t b = (uint64_t)blockIdx.x * 256 + threadIdx.x; b <= top_bits; b += (uint64_t)gridDim.x * 256) {
        const float w = __uint_as_float((uint32_t)b);
        const uint32_t want = __float_as_uint(powf(w, 4.0f));
        if (__float_as_uint(pow4_trim(w)) != want) bad++;
        // the two-lane form (pow4_trim2): this w in lane 0 next to a different argument, and in lane 1
        const float other = __uint_as_float(top_bits - (uint32_t)b);
        float a0, a1, b0, b1;
        unpack2(pow4_trim2(pack2(w, other)), a0, a1);
        unpack2(pow4_trim2(pack2(other, w)), b0, b1);
        if (__float_as_uint(a0) != want || __float_as_uint(b1) != want) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t verify_pow4(float top, unsigned long long* mismatches, void* scratch, cudaStream_t st) {
    unsigned long long* dev = (unsigned long long*)scratch;
    cudaError_t e = cudaSuccess;
    uint32_t top_bits;
    memcpy(&top_bits, &top, 4);
    e = cudaMemsetAsync(dev, 0, 8, st);
    if (e == cudaSuccess) {
        pow4_check_kernel<<<148 * 16, 256, 0, st>>>(top_bits, dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(mismatches, dev, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
}

// ---- order-independent digest of a particle state (include/pbf.h pbf_state_digest_*) ---------------------
// One 64-bit hash per particle over (iid, pos bits, vel bits); the digest is {sum, xor of a second mix} of the
// hashes, both commutative: two states holding the same particles in ANY order (one GPU's cell-sorted order, the
// concatenation of G slabs) have the same digest, and digests of disjoint parts combine by + and ^.
__global__ void __launch_bounds__(256)
digest_kernel(const float* __restrict__ pos, const float* __restrict__ vel, const uint32_t* __restrict__ iid, int64_t n,
              unsigned long long* __restrict__ out) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(pos);
    const uint32_t* v = reinterpret_cast<const uint32_t*>(vel);
    unsigned long long sum = 0, x = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const uint64_t h = particle_hash(iid[i], p[3 * i], p[3 * i + 1], p[3 * i + 2], v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        sum += h;
        x ^= digest_mix64(h);
    }
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        x ^= __shfl_xor_sync(0xffffffffu, x, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, sum);
        atomicXor(out + 1, x);
    }
}

cudaError_t launch_digest(const float* pos, const float* vel, const uint32_t* iid, int64_t n, unsigned long long* out,
                          cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, 16, st);
    if (e != cudaSuccess || n <= 0) return e;
    int64_t nb = (n + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    digest_kernel<<<(unsigned)nb, 256, 0, st>>>(pos, vel, iid, n, out);
    return cudaGetLastError();
}

void digest_host(const float* pos, const float* vel, const uint32_t* iid, int64_t n, uint64_t out[2]) {
    uint64_t sum = 0, x = 0;
    for (int64_t i = 0; i < n; i++) {
        uint32_t w[6];
        memcpy(w, pos + 3 * i, 12);
        memcpy(w + 3, vel + 3 * i, 12);
        const uint64_t h = particle_hash(iid[i], w[0], w[1], w[2], w[3], w[4], w[5]);
        sum += h;
        x ^= digest_mix64(h);
    }
    out[0] = sum;
    out[1] = x;
}

cudaError_t preload_stats() {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, digest_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, const_div_check_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, spiky_check_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, pow4_check_kernel);
    return e != cudaSuccess ? e : cudaFuncGetAttributes(&a, stats_kernel);
}

cudaError_t launch_stats(const float* rho, const float* npos, const float* nvel, int64_t n, float pho0,
                         double* partial, int nblocks, cudaStream_t st) {
    stats_kernel<<<nblocks, ST_THREADS, 0, st>>>(rho, npos, nvel, n, pho0, partial);
    return cudaGetLastError();
}

}  // namespace pbf
