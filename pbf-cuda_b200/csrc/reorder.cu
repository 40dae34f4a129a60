// reorder.cu — gather the particle state into cell-sorted SoA and build the cell table.
//
// Replaces (a) the payload movement inside the reference's sort_by_key (Simulator.cu:196-198),
// (b) computeGridRange (Simulator.cu:204, Simulator_kernel.cuh:21-50) — the two cudaMemset in front of it
//     (Simulator.cu:201-203) are folded into advect_key.cu.
// For sorted slot s with source index j:
//   x0[s]      = (advect(pos[j], vel[j]), 0)      float4, the iterate the solver works on
//   xs/ys/zs[s] = the same coordinates as three arrays, what the sweeps' cull reads
//   pos0[s]    = pos[j]                           tight float3, parked in the caller's npos buffer
//   iid_s[s]   = iid[j]
//   cell_range[key] = {first slot, one past last slot}; empty cells stay {0,0} (reference
//   semantics: gridStart == gridEnd == 0).
// The advected position is recomputed with the same two fma as in advect_key.cu, so the key the
// particle was sorted by is exactly the cell of x0[s].
//
// Slab mode: all `n` slots this rank stores are gathered (ghost planes included: a ghost's advected
// position is recomputed here from the raw state the owner sent, bit-identical to the owner's);
// pos0 is parked for the owned slots only.
//
// HBM traffic: R 8 (pair) + 28 (gathered pos, vel, iid; near-sequential because the input is
// last step's sorted order) ; W 16 + 12 + 12 + 4 per particle, + 8 B per occupied cell.
#include "launch.cuh"
#include "pbf_math.cuh"

namespace pbf {

constexpr int RO_THREADS = 256;

__global__ void __launch_bounds__(RO_THREADS)
reorder_kernel(const KeyIdx* __restrict__ sorted, const float* __restrict__ pos,
               const float* __restrict__ vel, const uint32_t* __restrict__ iid,
               float4* __restrict__ x0, float* __restrict__ xs, float* __restrict__ ys, float* __restrict__ zs,
               float* __restrict__ pos0_out, uint32_t* __restrict__ iid_sorted,
               uint2* __restrict__ cell_range, uint4* __restrict__ sort_zero, int64_t sort_zero_quads,
               int64_t n, int64_t own_first, int64_t own_count, const int64_t* __restrict__ plane_start, int gl, int nx,
               const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    pdl_wait();   // (launch.cuh: nothing of the previous kernel is touched before this)
    const int64_t s = (int64_t)blockIdx.x * RO_THREADS + threadIdx.x;
    // the sort is done: leave its histograms, tickets and look-back descriptors zeroed for the next one
    // (the invariant of pbf_sim::sort_zero), instead of a memset in front of every step
    for (int64_t k = s; k < sort_zero_quads; k += (int64_t)gridDim.x * RO_THREADS) sort_zero[k] = make_uint4(0u, 0u, 0u, 0u);
    // slab mode: how many of the sorted entries this rank keeps and which of them it owns is read from the plane
    // table on the DEVICE (slab.cu plane_table_kernel) — the host is still downloading that table while this runs
    if (plane_start) {
        n = plane_start[g.nxl];
        own_first = plane_start[gl];
        own_count = plane_start[gl + nx] - own_first;
    }
    if (s >= n) return;
    const KeyIdx e = sorted[s];
    const uint32_t prev = s == 0 ? 0xffffffffu : sorted[s - 1].key;
    const float3 p = load_f3(pos, e.idx), v = load_f3(vel, e.idx);
    const float3 q = advect_pos(p, v, c);
    x0[s] = make_float4(q.x, q.y, q.z, 0.f);
    xs[s] = q.x; ys[s] = q.y; zs[s] = q.z;   // the cull's copy of the coordinates (solver.cu CullSoA)
    if (s >= own_first && s - own_first < own_count) store_f3(pos0_out, s - own_first, p.x, p.y, p.z);
    iid_sorted[s] = iid[e.idx];
    // Green-style range detection (reference computeGridRange)
    if (e.key != prev) {
        cell_range[e.key].x = (uint32_t)s;
        if (prev != 0xffffffffu) cell_range[prev].y = (uint32_t)s;
    }
    if (s == n - 1) cell_range[e.key].y = (uint32_t)n;
}

// Lazy module loading (the CUDA default) loads a kernel at its first launch, and that load can wait for
// running kernels to finish: if the running kernel is a neighbour rank's flag wait (several ranks in one
// process), the two deadlock until the time-out. pbf_create therefore loads every kernel up front.
cudaError_t preload_reorder() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, reorder_kernel);
    return e;
}

cudaError_t launch_reorder(const KeyIdx* sorted, const float* pos, const float* vel, const uint32_t* iid,
                           float4* x0, CullScratch& cs, float* pos0_out, uint32_t* iid_sorted, uint2* cell_range,
                           uint32_t* sort_zero, size_t sort_zero_bytes, int64_t n, int64_t own_first, int64_t own_count,
                           const int64_t* plane_start, int gl, int nx, const GridConsts& g,
                           const SolverConsts& c, cudaStream_t st, int64_t* launches) {
    cs.holds = nullptr;
    if (n <= 0) return cudaSuccess;   // (nothing was sorted and nothing will be read)
    unsigned blocks = (unsigned)((n + RO_THREADS - 1) / RO_THREADS);
    PBF_LAUNCH((reorder_kernel), blocks, RO_THREADS, 0, st, sorted, pos, vel, iid, x0, cs.xs[0], cs.ys[0], cs.zs[0], pos0_out, iid_sorted, cell_range, reinterpret_cast<uint4*>(sort_zero), (int64_t)((sort_zero_bytes + 15) / 16), n, own_first, own_count, plane_start, gl, nx, g, c);
    if (launches) (*launches)++;
    cs.cur = 0;
    cs.holds = x0;   // every stored slot
    return cudaGetLastError();
}

}  // namespace pbf
