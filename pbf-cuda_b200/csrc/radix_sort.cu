// radix_sort.cu — hand-written onesweep LSD radix sort of (cell key, source index) pairs.
//
// Replaces the reference's thrust::sort_by_key over a 5-way zip of float3 payloads
// (Simulator.cu:196-198), which moves 56 B per particle per pass plus a pack and an unpack
// (~0.66 KB/particle). Here only 8 B pairs move; the payload is gathered once
// afterwards (reorder.cu).
//
// Algorithm (Adinets & Merrill "Onesweep", restated from the paper, not from CUB's sources):
//   - digit histograms of ALL passes are produced up front by advect_key.cu (raw counts; every tile scans the
//     256 counts of its pass itself);
//   - one kernel per 8-bit digit: each CTA takes a tile (4096 keys; 1024 below 128 K keys) through an atomic ticket (so that every
//     tile it may wait on is already resident), ranks its keys with warp-level match_any
//     (stable: items are visited in memory order), publishes its per-digit counts, resolves its
//     per-digit exclusive prefix by decoupled look-back over the preceding tiles, reorders the
//     tile through shared memory and writes coalesced runs.
//   - stable, deterministic, no temporary allocation.
//
// HBM traffic per pass: R 8 B + W 8 B per particle (pass 0 reads the 4 B key only).
#include "launch.cuh"
#include "pbf_internal.h"

namespace pbf {

namespace {

constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t FLAG_AGGREGATE = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;
constexpr int SORT_WARPS = SORT_THREADS / 32;
static_assert(SORT_THREADS == RADIX, "one thread per digit in the look-back phase");

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ITEMS keys per thread: 16 (tiles of 4096 keys) for large inputs, 4 (tiles of 1024) below SORT_SMALL_N keys, where
// the eight-tile look-back chain of a 32 K-particle scene was most of the pass (20 us for two passes).
template <bool FIRST, int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS)
onesweep_kernel(const uint32_t* __restrict__ keys_in, const KeyIdx* __restrict__ pairs_in,
                KeyIdx* __restrict__ out, const uint32_t* __restrict__ hist,
                uint32_t* __restrict__ tile_counter, uint32_t* __restrict__ tile_desc, int64_t n,
                int shift, const __grid_constant__ SlabInput si) {
    pdl_wait();
    constexpr int TILE = SORT_THREADS * ITEMS;
    __shared__ KeyIdx s_pairs[TILE];
    __shared__ uint32_t s_warp_hist[SORT_WARPS][RADIX];
    __shared__ uint32_t s_digit_start[RADIX];
    __shared__ uint32_t s_global_base[RADIX];
    __shared__ uint32_t s_wsum[SORT_WARPS];
    __shared__ uint32_t s_gsum[SORT_WARPS];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t total = hist[tid];   // (global count of digit `tid` in this pass; needed after the look-back)
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) s_warp_hist[w][tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int tile_n = (int)min((int64_t)TILE, n - tile_base);

    // ---- load, warp-striped so that (item k, lane l) is memory order within the warp's chunk
    KeyIdx item[ITEMS];
    const int warp_base = warp * 32 * ITEMS;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const int local = warp_base + k * 32 + lane;
        if (local < tile_n) {
            if (FIRST) {
                item[k].key = keys_in[tile_base + local];
                // the keys are in the sort's logical order; the index names the caller's slot
                item[k].idx = (uint32_t)slab_physical(si, tile_base + local);
            } else {
                item[k] = pairs_in[tile_base + local];
            }
        } else {
            item[k].key = 0xffffffffu;
            item[k].idx = 0xffffffffu;
        }
    }

    // ---- rank within the warp (stable)
    uint32_t rank[ITEMS];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const bool valid = (warp_base + k * 32 + lane) < tile_n;
        const uint32_t d = valid ? ((item[k].key >> shift) & (RADIX - 1)) : (uint32_t)RADIX;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t old = 0;
        if (valid) old = s_warp_hist[warp][d];
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) s_warp_hist[warp][d] = old + __popc(peers);
        __syncwarp();
        rank[k] = old + __popc(peers & lt_mask);
    }
    __syncthreads();

    // ---- per digit (thread == digit): exclusive scan over the warps, tile count
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
        uint32_t t = s_warp_hist[w][tid];
        s_warp_hist[w][tid] = count;
        count += t;
    }

    // ---- publish, then decoupled look-back for the exclusive prefix over preceding tiles
    uint32_t* my_desc = tile_desc + (size_t)tile * RADIX;
    uint32_t excl = 0;
    if (tile == 0) {
        st_volatile(my_desc + tid, count | FLAG_PREFIX);
    } else {
        st_volatile(my_desc + tid, count | FLAG_AGGREGATE);
        int64_t t = (int64_t)tile - 1;
        while (true) {
            uint32_t v = ld_volatile(tile_desc + (size_t)t * RADIX + tid);
            if ((v & FLAG_MASK) == 0) {
                __nanosleep(32);
                continue;
            }
            excl += v & VALUE_MASK;
            if (v & FLAG_PREFIX) break;
            t--;
        }
        st_volatile(my_desc + tid, (excl + count) | FLAG_PREFIX);
    }

    // ---- exclusive scans over the digits, two at once: of the tile counts (-> position in the tile-sorted
    // order) and of the pass's global digit counts (-> first global slot of a digit; every tile redoes this
    // 256-element scan, which is cheaper than a separate kernel in front of every sort)
    uint32_t v = count, gv = total;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, off), gt = __shfl_up_sync(0xffffffffu, gv, off);
        if (lane >= off) { v += t; gv += gt; }
    }
    if (lane == 31) { s_wsum[warp] = v; s_gsum[warp] = gv; }
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SORT_WARPS ? s_wsum[lane] : 0, iw = w;
        uint32_t gw = lane < SORT_WARPS ? s_gsum[lane] : 0, igw = gw;
#pragma unroll
        for (int off = 1; off < SORT_WARPS; off <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, iw, off), gt = __shfl_up_sync(0xffffffffu, igw, off);
            if (lane >= off) { iw += t; igw += gt; }
        }
        if (lane < SORT_WARPS) { s_wsum[lane] = iw - w; s_gsum[lane] = igw - gw; }
    }
    __syncthreads();
    const uint32_t dstart = v - count + s_wsum[warp];
    s_digit_start[tid] = dstart;
    s_global_base[tid] = (gv - total + s_gsum[warp]) + excl - dstart;  // + position in tile order = global slot
    __syncthreads();

    // ---- reorder the tile through shared memory
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        if ((warp_base + k * 32 + lane) < tile_n) {
            const uint32_t d = (item[k].key >> shift) & (RADIX - 1);
            s_pairs[s_digit_start[d] + s_warp_hist[warp][d] + rank[k]] = item[k];
        }
    }
    __syncthreads();

    // ---- coalesced runs out
    for (int p = tid; p < tile_n; p += SORT_THREADS) {
        const KeyIdx e = s_pairs[p];
        const uint32_t d = (e.key >> shift) & (RADIX - 1);
        out[s_global_base[d] + (uint32_t)p] = e;
    }
}

}  // namespace

// Lazy module loading (the CUDA default) loads a kernel at its first launch, and that load can wait for
// running kernels to finish: if the running kernel is a neighbour rank's flag wait (several ranks in one
// process), the two deadlock until the time-out. pbf_create therefore loads every kernel up front.
cudaError_t preload_sort() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, onesweep_kernel<true, SORT_ITEMS>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, onesweep_kernel<false, SORT_ITEMS>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, onesweep_kernel<true, SORT_ITEMS_SMALL>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, onesweep_kernel<false, SORT_ITEMS_SMALL>);
    return e;
}

static inline int sort_tile(int64_t n) { return SORT_THREADS * (n < SORT_SMALL_N ? SORT_ITEMS_SMALL : SORT_ITEMS); }
size_t sort_scratch_zero_bytes(int64_t n, int npass) {
    const int64_t tiles = (n + sort_tile(n) - 1) / sort_tile(n);
    return sizeof(uint32_t) * ((size_t)MAX_PASSES * RADIX + MAX_PASSES + (size_t)npass * tiles * RADIX);
}
size_t sort_scratch_capacity_bytes(int64_t max_n) {   // the largest of the above over every n <= max_n
    const size_t big = sort_scratch_zero_bytes(max_n, MAX_PASSES);
    const int64_t small_n = max_n < SORT_SMALL_N ? max_n : SORT_SMALL_N - 1;
    const size_t small = sort_scratch_zero_bytes(small_n, MAX_PASSES);
    return big > small ? big : small;
}

cudaError_t launch_sort(const uint32_t* keys, SortScratch& s, int64_t n, int npass, const SlabInput& si,
                        int* result_buf, cudaStream_t st, int64_t* launches) {
    if (n <= 0) { *result_buf = 0; return cudaSuccess; }
    const bool small = n < SORT_SMALL_N;
    const int64_t tiles = (n + sort_tile(n) - 1) / sort_tile(n);
    for (int p = 0; p < npass; p++) {
        uint32_t* desc = s.tile_desc + (size_t)p * tiles * RADIX;
        const uint32_t* h = s.hist + p * RADIX;
        if (p == 0 && small)
            PBF_LAUNCH((onesweep_kernel<true, SORT_ITEMS_SMALL>), (unsigned)tiles, SORT_THREADS, 0, st, keys, nullptr, s.bufs[0], h, s.tile_counter + p, desc, n, p * RADIX_BITS, si);
        else if (p == 0)
            PBF_LAUNCH((onesweep_kernel<true, SORT_ITEMS>), (unsigned)tiles, SORT_THREADS, 0, st, keys, nullptr, s.bufs[0], h, s.tile_counter + p, desc, n, p * RADIX_BITS, si);
        else if (small)
            PBF_LAUNCH((onesweep_kernel<false, SORT_ITEMS_SMALL>), (unsigned)tiles, SORT_THREADS, 0, st, nullptr, s.bufs[(p - 1) & 1], s.bufs[p & 1], h, s.tile_counter + p, desc, n, p * RADIX_BITS, si);
        else
            PBF_LAUNCH((onesweep_kernel<false, SORT_ITEMS>), (unsigned)tiles, SORT_THREADS, 0, st, nullptr, s.bufs[(p - 1) & 1], s.bufs[p & 1], h, s.tile_counter + p, desc, n, p * RADIX_BITS, si);
        if (launches) (*launches)++;
    }
    *result_buf = (npass - 1) & 1;
    return cudaGetLastError();
}

}  // namespace pbf
