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
constexpr int AK_PER_WIDE = 4;                   // particles per thread: 48 B of pos and of vel = three 16-byte loads each
constexpr int64_t AK_SMALL_N = 256 * 1024;       // below: one particle per thread (a 32 000-particle scene is 125 blocks, not 32)
constexpr int AK_BLOCKS_PER_SM = 8;              // persistent grid: a block keeps its digit counts over many tiles
constexpr int AK_BINS = RADIX + 1;               // (one more bin per pass for the slots past n of the last tile)

// VEC: pos / vel are 16-byte aligned (every cudaMalloc'ed or torch buffer is): a thread's four particles are six
// 128-bit loads instead of 24 scalar ones.
template <bool VEC, int AK_PER>
__global__ void __launch_bounds__(AK_THREADS)
advect_key_kernel(const float* __restrict__ pos, const float* __restrict__ vel,
                  uint32_t* __restrict__ keys, uint32_t* __restrict__ hist, uint4* __restrict__ cell_clear,
                  int64_t n, int npass,
                  const __grid_constant__ SlabInput si, const __grid_constant__ GridConsts g,
                  const __grid_constant__ SolverConsts c) {
    pdl_wait();   // (launch.cuh: nothing of the previous kernel is touched before this)
    constexpr int AK_TILE = AK_THREADS * AK_PER;
    __shared__ uint32_t s_hist[MAX_PASSES * AK_BINS];
    for (int k = threadIdx.x; k < npass * AK_BINS; k += AK_THREADS) s_hist[k] = 0;
    __syncthreads();

    // the cell table the reorder pass fills next is emptied here (the reference's two cudaMemset, Simulator.cu:
    // 201-203): its last reader was the previous step's XSPH sweep. Two cells per 16-byte store.
    if (cell_clear)
        for (int64_t k = (int64_t)blockIdx.x * AK_THREADS + threadIdx.x; k < (int64_t)(g.ncell + 1) / 2; k += (int64_t)gridDim.x * AK_THREADS)
            cell_clear[k] = make_uint4(0u, 0u, 0u, 0u);

    const unsigned lane = threadIdx.x & 31;
    const bool slab = si.flags != nullptr || si.m_left != 0 || si.n_own != n;
    const int64_t ntiles = (n + AK_TILE - 1) / AK_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * AK_TILE + (int64_t)threadIdx.x * AK_PER;
        float pv[2][3 * AK_PER];
        const bool full = base + AK_PER <= n;
        if (VEC && AK_PER == 4 && full) {
            const float4* p4 = reinterpret_cast<const float4*>(pos + 3 * base);
            const float4* v4 = reinterpret_cast<const float4*>(vel + 3 * base);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float4 a = __ldg(p4 + k), b = __ldg(v4 + k);
                pv[0][4 * k] = a.x; pv[0][4 * k + 1] = a.y; pv[0][4 * k + 2] = a.z; pv[0][4 * k + 3] = a.w;
                pv[1][4 * k] = b.x; pv[1][4 * k + 1] = b.y; pv[1][4 * k + 2] = b.z; pv[1][4 * k + 3] = b.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3 * AK_PER; k++) {
                const bool in = base + k / 3 < n;
                pv[0][k] = in ? pos[3 * base + k] : 0.f;
                pv[1][k] = in ? vel[3 * base + k] : 0.f;
            }
        }
        uint32_t key[AK_PER];
#pragma unroll
        for (int j = 0; j < AK_PER; j++) {
            const int64_t i = base + j;
            const float3 q = advect_pos(make_float3(pv[0][3 * j], pv[0][3 * j + 1], pv[0][3 * j + 2]),
                                        make_float3(pv[1][3 * j], pv[1][3 * j + 1], pv[1][3 * j + 2]), c);
            const int3 cc = cell_of(q.x, q.y, q.z, g);
            const int lx = cc.x - g.xoff;
            key[j] = (lx >= 0 && lx < g.nxl) ? (uint32_t)cell_id(cc.x, cc.y, cc.z, g) : (uint32_t)g.ncell;
            if (si.flags && i < si.n_own &&
                ((cc.x < si.need_left_below && i >= si.send_left_end) ||
                 (cc.x >= si.need_right_from && i < si.send_right_begin)))
                atomicOr(si.flags, (uint32_t)PBF_SLAB_FLAG_MIGRATION);
        }
        if (AK_PER == 4 && full && !slab) {
            *reinterpret_cast<uint4*>(keys + base) = make_uint4(key[0], key[1 % AK_PER], key[2 % AK_PER], key[3 % AK_PER]);
        } else {
#pragma unroll
            for (int j = 0; j < AK_PER; j++)
                if (base + j < n) keys[slab ? slab_logical(si, base + j) : base + j] = key[j];
        }
        // Digit counts by run length. The caller's order is last step's sorted order, so the keys of a warp's 128
        // consecutive particles form a few runs of equal digits (one or two for the high digits, ~16 for the
        // lowest). A run from element a to element b adds b - a to its bin: every BOUNDARY between elements e - 1
        // and e adds e to the bin of the digit that ends and subtracts e from the bin of the digit that starts
        // (unsigned wrap-around makes the partial sums harmless), and the warp's last element closes the last run
        // with + 128 — two shared-memory atomics per boundary instead of one match_any per element and pass.
        for (int p = 0; p < npass; p++) {
            uint32_t d[AK_PER];
#pragma unroll
            for (int j = 0; j < AK_PER; j++) d[j] = base + j < n ? ((key[j] >> (p * RADIX_BITS)) & (RADIX - 1)) : (uint32_t)RADIX;
            uint32_t prev = __shfl_up_sync(0xffffffffu, d[AK_PER - 1], 1);
            if (lane == 0) prev = d[0];
            uint32_t* bins = s_hist + p * AK_BINS;
#pragma unroll
            for (int j = 0; j < AK_PER; j++) {
                if (d[j] != prev) {
                    const uint32_t e = lane * AK_PER + j;
                    atomicAdd(&bins[prev], e);
                    atomicSub(&bins[d[j]], e);
                }
                prev = d[j];
            }
            if (lane == 31) atomicAdd(&bins[prev], 32u * AK_PER);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < npass * RADIX; k += AK_THREADS) {
        const uint32_t v = s_hist[(k / RADIX) * AK_BINS + (k % RADIX)];
        if (v) atomicAdd(&hist[k], v);
    }
}

// Lazy module loading (the CUDA default) loads a kernel at its first launch, and that load can wait for
// running kernels to finish: if the running kernel is a neighbour rank's flag wait (several ranks in one
// process), the two deadlock until the time-out. pbf_create therefore loads every kernel up front.
cudaError_t preload_advect_key() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, advect_key_kernel<true, AK_PER_WIDE>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, advect_key_kernel<false, AK_PER_WIDE>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, advect_key_kernel<false, 1>);
    return e;
}

cudaError_t launch_advect_key(const float* pos, const float* vel, uint32_t* keys, uint32_t* hist, uint2* cell_clear,
                              int64_t n, int npass, const SlabInput& si, const GridConsts& g,
                              const SolverConsts& c, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    static int sms = 0;   // (same for every device of a B200 box; a wrong count only changes the grid size)
    if (sms == 0 && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0) != cudaSuccess) { cudaGetLastError(); sms = 148; }
    const int per = n < AK_SMALL_N ? 1 : AK_PER_WIDE;
    const int64_t tiles = (n + (int64_t)AK_THREADS * per - 1) / ((int64_t)AK_THREADS * per);
    const unsigned blocks = (unsigned)(tiles < (int64_t)sms * AK_BLOCKS_PER_SM ? tiles : (int64_t)sms * AK_BLOCKS_PER_SM);
    if (per == 1)
        PBF_LAUNCH((advect_key_kernel<false, 1>), blocks, AK_THREADS, 0, st, pos, vel, keys, hist, reinterpret_cast<uint4*>(cell_clear), n, npass, si, g, c);
    else if ((((uintptr_t)pos | (uintptr_t)vel) & 15u) == 0)
        PBF_LAUNCH((advect_key_kernel<true, AK_PER_WIDE>), blocks, AK_THREADS, 0, st, pos, vel, keys, hist, reinterpret_cast<uint4*>(cell_clear), n, npass, si, g, c);
    else
        PBF_LAUNCH((advect_key_kernel<false, AK_PER_WIDE>), blocks, AK_THREADS, 0, st, pos, vel, keys, hist, reinterpret_cast<uint4*>(cell_clear), n, npass, si, g, c);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace pbf
