// advect_key.cu — stage 1 of the step, in the caller's particle order.
//
// Fuses the reference's advect_kernel (Simulator_kernel.cuh:7-19) with the getGridId transform
// (Simulator.cu:56-73, called at :190-193) and with the digit histograms the onesweep sort needs
// for all of its passes (CUB runs that as a separate DeviceRadixSortHistogramKernel over the
// keys). The advected velocity / position are NOT written here: the reorder pass recomputes them
// from pos/vel with the same two fma (bit-identical), which saves 24 B/particle of HBM writes and
// 24 B/particle of reads.
//
// HBM traffic: R 24 B (pos, vel) + W 4 B (key) per particle, + W 8 B per cell (the cell table is emptied here).
//
// Slab mode (multi-GPU, slab.cu): the caller's arrays hold [own | from left | from right]; keys are
// written in the sort's logical order [from left | own | from right] (SlabInput), particles that
// land outside the planes this rank stores get the discard key (one past the last local cell, so
// the sort parks them behind everything else), and an own particle that a neighbour rank needed
// but was not in the range sent to it raises PBF_SLAB_FLAG_MIGRATION.
#include "launch.cuh"
#include "pbf_math.cuh"

namespace pbf {

constexpr int AK_THREADS = 256;

__global__ void __launch_bounds__(AK_THREADS)
advect_key_kernel(const float* __restrict__ pos, const float* __restrict__ vel,
                  uint32_t* __restrict__ keys, uint32_t* __restrict__ hist, uint4* __restrict__ cell_clear,
                  int64_t n, int npass,
                  const __grid_constant__ SlabInput si, const __grid_constant__ GridConsts g,
                  const __grid_constant__ SolverConsts c) {
    pdl_wait();   // (launch.cuh: nothing of the previous kernel is touched before this)
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    for (int k = threadIdx.x; k < npass * RADIX; k += AK_THREADS) s_hist[k] = 0;
    __syncthreads();

    const int64_t i = (int64_t)blockIdx.x * AK_THREADS + threadIdx.x;
    // the cell table the reorder pass fills next is emptied here (the reference's two cudaMemset, Simulator.cu:
    // 201-203): its last reader was the previous step's XSPH sweep. Two cells per 16-byte store.
    if (cell_clear)
        for (int64_t k = i; k < (int64_t)(g.ncell + 1) / 2; k += (int64_t)gridDim.x * AK_THREADS) cell_clear[k] = make_uint4(0u, 0u, 0u, 0u);
    const bool valid = i < n;
    uint32_t key = 0;
    if (valid) {
        float3 p = load_f3(pos, i), v = load_f3(vel, i);
        float3 q = advect_pos(p, v, c);
        int3 cc = cell_of(q.x, q.y, q.z, g);
        const int lx = cc.x - g.xoff;
        key = (lx >= 0 && lx < g.nxl) ? (uint32_t)cell_id(cc.x, cc.y, cc.z, g) : (uint32_t)g.ncell;
        if (si.flags && i < si.n_own &&
            ((cc.x < si.need_left_below && i >= si.send_left_end) ||
             (cc.x >= si.need_right_from && i < si.send_right_begin)))
            atomicOr(si.flags, (uint32_t)PBF_SLAB_FLAG_MIGRATION);
        keys[slab_logical(si, i)] = key;
    }
    // warp-aggregated shared-memory histogram: particles of one block share their high digits,
    // so a plain atomicAdd per thread would serialise on one bank word.
    const unsigned lane = threadIdx.x & 31;
    for (int p = 0; p < npass; p++) {
        uint32_t d = valid ? ((key >> (p * RADIX_BITS)) & (RADIX - 1)) : (uint32_t)RADIX;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_hist[p * RADIX + d], __popc(peers));
    }
    __syncthreads();
    for (int k = threadIdx.x; k < npass * RADIX; k += AK_THREADS) {
        uint32_t v = s_hist[k];
        if (v) atomicAdd(&hist[k], v);
    }
}

// Lazy module loading (the CUDA default) loads a kernel at its first launch, and that load can wait for
// running kernels to finish: if the running kernel is a neighbour rank's flag wait (several ranks in one
// process), the two deadlock until the time-out. pbf_create therefore loads every kernel up front.
cudaError_t preload_advect_key() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, advect_key_kernel);
    return e;
}

cudaError_t launch_advect_key(const float* pos, const float* vel, uint32_t* keys, uint32_t* hist, uint2* cell_clear,
                              int64_t n, int npass, const SlabInput& si, const GridConsts& g,
                              const SolverConsts& c, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    unsigned blocks = (unsigned)((n + AK_THREADS - 1) / AK_THREADS);
    PBF_LAUNCH((advect_key_kernel), blocks, AK_THREADS, 0, st, pos, vel, keys, hist, reinterpret_cast<uint4*>(cell_clear), n, npass, si, g, c);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace pbf
