// launch.cuh — how the step's kernels are launched: programmatic dependent launch (PDL).
//
// A step is a chain of 20-odd short kernels on one stream (the reference: 6 + 3*niter kernel / Thrust calls with a
// cudaDeviceSynchronize after every stage, Simulator.cpp:44-78). Back to back on a stream, kernel k+1's CTAs are
// only dispatched after kernel k has drained completely: a few microseconds of empty machine per boundary — a
// fifth of the reference's own 32 000-particle scene. With the programmatic-stream-serialization attribute the
// next grid is dispatched while the previous one is still finishing and its CTAs block in `griddepcontrol.wait`
// (the FIRST statement of every kernel, before any global memory access: no read-after-write or
// write-after-read hazard can arise) until the previous grid has completed and flushed its writes. What
// overlaps is launch latency, CTA dispatch and the parameter / index prologue. The same edges survive stream
// capture into a CUDA graph (pbf_capi.cu).
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace pbf {

// blocks until the preceding kernel of the stream (if this one was launched as its programmatic dependent) has
// completed and its memory operations are visible; a no-op otherwise
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// set by the C-ABI's stage functions from the handle's PBF_OPT_PDL before they call the launchers (thread-local:
// handles used from different host threads do not see each other's setting)
extern thread_local bool tl_pdl;

template <typename... Params, typename... Args>
inline cudaError_t launch_k(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tl_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace pbf

// PBF_LAUNCH((kernel<...>), grid, block, smem, stream, args...) == kernel<...><<<grid, block, smem, stream>>>(args...)
#define PBF_LAUNCH(kernel, grid, block, smem, st, ...) \
    (void)::pbf::launch_k(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)
