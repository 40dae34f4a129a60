// slab.cu — device helpers of the multi-GPU x-slab decomposition (the reference is
// single-GPU, so there is no reference code for this file — what it must preserve is the
// reference's RESULT: every rank reproduces, for the particles it owns, the bits the single-GPU
// step produces for them).
//
// Why x-slabs are plain ranges here: the sort key is the reference's x-major cell id
// (Simulator.cu:45-53), so in the sorted arrays every cell plane x = const is one contiguous slot
// range. A rank stores planes [xoff, xoff + nxl) = ghost | owned | ghost; after its local stable
// sort the ghost particles sit at the two ends, the planes a neighbour needs as ITS ghosts are a
// prefix / suffix of the owned range, and every halo refresh (lambda, positions, velocity+rho)
// is a copy of one contiguous float4 range straight out of / into the solver's own arrays.
#include <stdio.h>

#include "launch.cuh"
#include "pbf_internal.h"

namespace pbf {

namespace {

// plane_start[p] = first sorted slot whose key is >= p*dyz, p = 0..nxl. Keys ascend; the discard
// key (== ncell == nxl*dyz) is behind every local cell, so plane_start[nxl] = particles kept.
__global__ void plane_table_kernel(const KeyIdx* __restrict__ sorted, int64_t n,
                                   int64_t* __restrict__ plane_start, int nxl, int dyz) {
    pdl_wait();   // (launch.cuh: nothing of the previous kernel is touched before this)
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nxl) return;
    const uint32_t want = (uint32_t)p * (uint32_t)dyz;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid].key < want) lo = mid + 1; else hi = mid;
    }
    plane_start[p] = lo;
}

__global__ void __launch_bounds__(256)
gather_state_kernel(const KeyIdx* __restrict__ sorted, const float* __restrict__ pos,
                    const float* __restrict__ vel, const uint32_t* __restrict__ iid,
                    float* __restrict__ npos, float* __restrict__ nvel, uint32_t* __restrict__ iid_out,
                    int64_t n) {
    pdl_wait();
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (s >= n) return;
    const uint32_t j = sorted[s].idx;
    npos[3 * s] = pos[3 * (int64_t)j]; npos[3 * s + 1] = pos[3 * (int64_t)j + 1]; npos[3 * s + 2] = pos[3 * (int64_t)j + 2];
    nvel[3 * s] = vel[3 * (int64_t)j]; nvel[3 * s + 1] = vel[3 * (int64_t)j + 1]; nvel[3 * s + 2] = vel[3 * (int64_t)j + 2];
    iid_out[s] = iid[j];
}

// ---- flag handshake of the fused halo refresh ------------------------------------------------
// After a pass whose kernel pushed its boundary values into the neighbours' ghost slots, a rank
// tells both neighbours "my pushes of refresh #seq are complete" and waits for theirs. Both are
// one-thread kernels in stream order: the signal runs after the pushing kernel has retired (its
// remote stores are performed), the wait blocks the stream until the neighbours' words arrive.
__global__ void halo_signal_kernel(const int64_t* tail_src, int64_t* peer_right_tail, uint32_t* peer_word_left,
                                   uint32_t* peer_word_right, uint32_t seq) {
    pdl_wait();
    if (peer_right_tail) *(volatile int64_t*)peer_right_tail = *tail_src;
    __threadfence_system();
    if (peer_word_left) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_word_left), "r"(seq) : "memory");
    if (peer_word_right) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_word_right), "r"(seq) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// (int32 difference: the sequence number may wrap)
__global__ void halo_wait_kernel(const uint32_t* word_left, const uint32_t* word_right, uint32_t seq,
                                 uint64_t timeout_ns, uint32_t* flags) {
    pdl_wait();
    const uint64_t t0 = global_ns();
    for (int side = 0; side < 2; side++) {
        const uint32_t* w = side == 0 ? word_left : word_right;
        if (!w) continue;
        while ((int32_t)(ld_acquire_sys(w) - seq) < 0) {
            if (global_ns() - t0 > timeout_ns) {   // a dead neighbour must not hang the device
                if (flags) atomicOr(flags, (uint32_t)PBF_SLAB_FLAG_TIMEOUT);
                printf("pbf halo wait timed out: side %d, waiting for handshake %u, word holds %u\n", side, seq, ld_acquire_sys(w));
                return;
            }
            __nanosleep(64);
        }
    }
}

}  // namespace

// Lazy module loading (the CUDA default) loads a kernel at its first launch, and that load can wait for
// running kernels to finish: if the running kernel is a neighbour rank's flag wait (several ranks in one
// process), the two deadlock until the time-out. pbf_create therefore loads every kernel up front.
cudaError_t preload_slab() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, plane_table_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, gather_state_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, halo_signal_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, halo_wait_kernel);
    return e;
}

cudaError_t launch_halo_signal(uint32_t* peer_word_left, uint32_t* peer_word_right, uint32_t seq, cudaStream_t st,
                               int64_t* launches) {
    if (!peer_word_left && !peer_word_right) return cudaSuccess;
    PBF_LAUNCH((halo_signal_kernel), 1, 1, 0, st, nullptr, nullptr, peer_word_left, peer_word_right, seq);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_halo_publish(const int64_t* tail_src, int64_t* peer_right_tail, uint32_t* peer_word_left,
                                uint32_t* peer_word_right, uint32_t seq, cudaStream_t st, int64_t* launches) {
    if (!peer_word_left && !peer_word_right) return cudaSuccess;
    PBF_LAUNCH((halo_signal_kernel), 1, 1, 0, st, tail_src, peer_right_tail, peer_word_left, peer_word_right, seq);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_halo_wait(const uint32_t* word_left, const uint32_t* word_right, uint32_t seq,
                             uint64_t timeout_ns, uint32_t* flags, cudaStream_t st, int64_t* launches) {
    if (!word_left && !word_right) return cudaSuccess;
    PBF_LAUNCH((halo_wait_kernel), 1, 1, 0, st, word_left, word_right, seq, timeout_ns, flags);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_plane_table(const KeyIdx* sorted, int64_t n, int64_t* plane_start, const GridConsts& g,
                               cudaStream_t st, int64_t* launches) {
    const int threads = 128;
    PBF_LAUNCH((plane_table_kernel), (g.nxl + 1 + threads - 1) / threads, threads, 0, st, sorted, n, plane_start, g.nxl, g.dyz);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_gather_state(const KeyIdx* sorted, const float* pos, const float* vel, const uint32_t* iid,
                                float* npos, float* nvel, uint32_t* iid_out, int64_t n, cudaStream_t st,
                                int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    PBF_LAUNCH((gather_state_kernel), (unsigned)((n + 255) / 256), 256, 0, st, sorted, pos, vel, iid, npos, nvel, iid_out, n);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace pbf
