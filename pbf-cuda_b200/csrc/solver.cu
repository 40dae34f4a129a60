// solver.cu — the constraint-solver passes of the step on the cell-sorted float4 SoA.
//
//   lambda pass   <- computeLambda  (reference Simulator_kernel.cuh:52-129)
//   delta-p pass  <- computetpos    (reference Simulator_kernel.cuh:131-194) + the Jacobi commit
//                    thrust::copy_n (Simulator.cu:247-248), which becomes a ping-pong
//   velocity      <- h_updateVelocity (Simulator.cu:127-137, 267-274)
//   XSPH          <- computeXSPH    (reference Simulator_kernel.cuh:196-239)
//
// Parity design: a particle's sums are accumulated by ONE thread in the reference's visiting
// order (dx, dy, dz nested, ascending slot inside a cell) with the reference's exact fp32
// operation sequence (pbf_math.cuh), so rho / lambda / positions reproduce the reference's CUDA
// build bit for bit (with exact_pow) instead of merely within tolerance. What changes is the
// data path: the three z-cells of a column visited as one contiguous slot run (9 runs instead
// of 27 cells); a cull that tests four candidates with three 16-byte loads from coordinate
// arrays and 14 two-lane FP32 instructions and records the hits as bits; and the expensive part
// (sqrt, 4 IEEE divisions, pow) only for the pairs that passed — candidates outside h contribute
// exact zeros in the reference, so skipping them does not change a single bit — with the
// neighbour list of the lambda pass handed to the delta-p pass of the same iteration.
#include "launch.cuh"
#include "solver_common.cuh"

#ifndef PBF_CULL_UNROLL
#define PBF_CULL_UNROLL 1
#endif

namespace pbf {

constexpr int CULL_UNROLL = PBF_CULL_UNROLL;
// PBF_CULL_WIDE: the cull takes EIGHT slots per trip with three 256-bit loads (LDG.E.256, sm_100) instead of four with
// three 128-bit ones: half the load instructions — and, where the lanes of a warp read the same few lines, half the L1
// tag requests — per candidate. Words then start at multiples of eight slots.
#ifndef PBF_CULL_WIDE
#define PBF_CULL_WIDE 0
#endif
constexpr bool CULL_WIDE = PBF_CULL_WIDE != 0;
struct float8 { float4 lo, hi; };
__device__ __forceinline__ float8 ldg256(const float* p) {
    float8 v;
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y), "=f"(v.hi.z), "=f"(v.hi.w) : "l"(p));
    return v;
}

// Two-phase gather of one particle (one thread), the core of all three neighbour sweeps.
//
// Phase 1 (cull) walks the candidates in the reference's visiting order — dx, dy, dz nested,
// ascending slot inside a cell; the three dz cells of a column are consecutive keys, hence ONE
// contiguous slot run per (dx, dy), and the runs are visited in ascending slot order — with
// three 16-byte loads, 14 two-lane flops and four funnel shifts per FOUR candidates and nothing
// else: the sign bit of RN(r2 - limit) (set exactly when r2 < limit; IEEE subtraction with
// denormals never rounds a non-zero difference to zero) is shifted into a hit word, 32 slots per
// word, first slot in the top bit. Non-empty words go to a per-thread list in shared memory as
// (first slot, hits) (entry k of thread t at [k][t]: conflict-free). Runs are read in aligned
// groups of four slots, up to three before their start and after their end (the arrays are
// padded); those bits are masked off.
// Phase 2 (`heavy`) then runs the expensive exact arithmetic over the set bits only — count
// leading zeros, clear, next word when empty — every lane busy, still in visiting order.
// Without the split a warp executes the heavy path for nearly every candidate, because some
// lane is almost always in range (~15 % of the candidates are), at ~15 % lane utilisation.
// The list is drained once per particle — one heavy phase over ~33 neighbours keeps the lanes of a
// warp busier than three phases over ~11 (measured: 3.78 -> 3.01 ms per step) — and whenever it
// is full, so any neighbour count stays correct and ordered; 15 words per thread = 15 KB per CTA
// keep 8 CTAs inside the 132 KB carve-out step and leave ~124 KB of L1.
// (Measured and dropped: issuing the next group's loads before this group's arithmetic, and loading the
//  next neighbour ahead of the heavy arithmetic — no change in either case. The sweeps are bound by the L1
//  wavefront rate (70-80 % of peak, ncu): lanes of a warp sit in ~3 cells, and after the first Jacobi
//  iteration their home cells drift apart, so one warp load touches 3-9 different lines.)
// STAGED (the TMA A/B of DESIGN.md 3.3, lambda_staged_kernel below): the block has copied, per (dx, dy) run, the
// union of its threads' runs into shared memory with cp.async.bulk; the cull then reads from there —
// `st_base` = shared address of the staged coordinates [run][x|y|z][STAGE_CAP], `st_ubase[run]` = first staged slot.
constexpr int STAGE_CAP = 192;   // slots staged per run (9 runs x 3 arrays x 192 x 4 B = 20.25 KB per CTA)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// MORTON (the A/B of DESIGN.md 3.1): keys are bit-interleaved, the three z cells of a column are not neighbours in the
// sorted order any more — 27 one-cell runs per particle instead of 9 three-cell runs, in the same dx, dy, dz order.
template <bool SKIP_SELF, bool STAGED = false, bool MORTON = false, typename Heavy>
__device__ __forceinline__ void gather(const float4 p, const uint32_t self, const float limit,
                                       const float4* __restrict__ x, const CullSoA soa,
                                       const uint2* __restrict__ cell_range, const GridConsts& g,
                                       uint2* __restrict__ my_words, Heavy&& heavy, uint32_t st_base = 0,
                                       const uint32_t* st_ubase = nullptr) {
    const int3 cc = cell_of(p.x, p.y, p.z, g);
    const bool has_below = cc.z > 0, has_above = cc.z + 1 < g.dim[2];
    uint2* const words_end = my_words + WORD_CAP * GATHER_THREADS;
    uint2* tail = my_words;  // next free entry of this thread's list
    int k_total = 0;         // in-range neighbours handed to `heavy` so far (its third argument)
    const f32x2 px = pack2(p.x, p.x), py = pack2(p.y, p.y), pz = pack2(p.z, p.z), lim = pack2(limit, limit);
    auto flush = [&]() {
        const uint2* e = my_words;
        uint32_t first = 0, hits = 0;
        for (;;) {
            if (hits == 0) {  // next word of the list
                if (e == tail) break;
                const uint2 w = *e;
                e += GATHER_THREADS;
                first = w.x;
                hits = w.y;
            }
            const int lead = __clz((int)hits);
            hits &= ~(0x80000000u >> lead);
            const uint32_t j = first + (uint32_t)lead;
            if (SKIP_SELF && j == self) continue;
            heavy(j, __ldg(&x[j]), k_total);
            k_total++;
        }
        tail = my_words;
    };
#pragma unroll 1
    for (int dx = -1; dx <= 1; dx++) {
        const int cx = cc.x + dx;
        const int lx = cx - g.xoff;  // plane in this handle's table (slab mode; lx == cx on one GPU)
        if (cx < 0 || cx >= g.dim[0]) continue;
        if (lx < 0 || lx >= g.nxl) {
            // the particle drifted so far from its stored cell that its search leaves the ghost
            // planes: the result would silently miss neighbours, so say so
            if (g.flags) atomicOr(g.flags, (uint32_t)PBF_SLAB_FLAG_GHOST);
            continue;
        }
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++) {
            const int cy = cc.y + dy;
            if (cy < 0 || cy >= g.dim[1]) continue;
            const int cbase = lx * g.dyz + cy * g.dim[2];
            const uint32_t mxy = MORTON ? (__ldg(g.morton + cx) | __ldg(g.morton + 1024 + cy)) : 0u;
#pragma unroll 1
            for (int zi = 0; zi < (MORTON ? 3 : 1); zi++) {
                uint32_t start, end;
                if (MORTON) {   // one cell per run
                    const int cz = cc.z + zi - 1;
                    if (cz < 0 || cz >= g.dim[2]) continue;
                    const uint2 r = __ldg(&cell_range[mxy | __ldg(g.morton + 2048 + cz)]);
                    start = r.x;
                    end = r.y;
                } else {
                    // the run = from the first slot of the first non-empty cell of the column's (up to) three to the
                    // end of the last non-empty one; empty and out-of-range cells read {0, 0}, so an empty column gives
                    // start == end == 0. Straight-line on purpose: as a loop over z (1-3 trips) this was ~80 instructions.
                    const uint2 zero = make_uint2(0u, 0u);
                    const uint2 r0 = has_below ? __ldg(&cell_range[cbase + cc.z - 1]) : zero;
                    const uint2 r1 = __ldg(&cell_range[cbase + cc.z]);
                    const uint2 r2 = has_above ? __ldg(&cell_range[cbase + cc.z + 1]) : zero;
                    const bool e0 = r0.y > r0.x, e1 = r1.y > r1.x, e2 = r2.y > r2.x;
                    start = e0 ? r0.x : e1 ? r1.x : r2.x;
                    end = e2 ? r2.y : e1 ? r1.y : r0.y;
                }
#pragma unroll 1
                for (uint32_t b = start & ((CULL_WIDE && !STAGED) ? ~7u : ~3u); b < end; b += 32) {   // words start at multiples of four (eight) slots
                    const uint32_t cnt = min(end - b, 32u);   // slots of this word up to the end of the run
                    const uint32_t groups = (CULL_WIDE && !STAGED) ? ((cnt + 7) >> 3) << 1 : (cnt + 3) >> 2;   // (in units of four slots)
                    uint32_t hits = 0;
                    if (CULL_WIDE && !STAGED) {
#pragma unroll 1
                        for (uint32_t gi = 0; gi < groups; gi += 2) {   // eight candidates: 3 loads, 28 packed flops, 8 shifts
                            const float8 X = ldg256(soa.xs + b + 4 * gi), Y = ldg256(soa.ys + b + 4 * gi), Z = ldg256(soa.zs + b + 4 * gi);
                            hits = push_hits2(hits, px, py, pz, lim, X.lo.x, X.lo.y, Y.lo.x, Y.lo.y, Z.lo.x, Z.lo.y);
                            hits = push_hits2(hits, px, py, pz, lim, X.lo.z, X.lo.w, Y.lo.z, Y.lo.w, Z.lo.z, Z.lo.w);
                            hits = push_hits2(hits, px, py, pz, lim, X.hi.x, X.hi.y, Y.hi.x, Y.hi.y, Z.hi.x, Z.hi.y);
                            hits = push_hits2(hits, px, py, pz, lim, X.hi.z, X.hi.w, Y.hi.z, Y.hi.w, Z.hi.z, Z.hi.w);
                        }
                    } else if (STAGED) {
                        const int run = (dx + 1) * 3 + (dy + 1);
                        uint32_t a = st_base + ((uint32_t)run * 3u * STAGE_CAP + (b - st_ubase[run])) * 4u;
#pragma unroll 1
                        for (uint32_t gi = 0; gi < groups; gi++, a += 16u) {  // the same four candidates out of shared memory
                            const float4 X = lds128(a), Y = lds128(a + STAGE_CAP * 4u), Z = lds128(a + 2u * STAGE_CAP * 4u);
                            hits = push_hits2(hits, px, py, pz, lim, X.x, X.y, Y.x, Y.y, Z.x, Z.y);
                            hits = push_hits2(hits, px, py, pz, lim, X.z, X.w, Y.z, Y.w, Z.z, Z.w);
                        }
                    } else {
                        const float4* xp = reinterpret_cast<const float4*>(soa.xs + b);
                        const float4* yp = reinterpret_cast<const float4*>(soa.ys + b);
                        const float4* zp = reinterpret_cast<const float4*>(soa.zs + b);
#pragma unroll CULL_UNROLL
                        for (uint32_t gi = 0; gi < groups; gi++) {  // four candidates: 3 loads, 14 packed flops, 4 shifts
                            const float4 X = __ldg(xp + gi), Y = __ldg(yp + gi), Z = __ldg(zp + gi);
                            hits = push_hits2(hits, px, py, pz, lim, X.x, X.y, Y.x, Y.y, Z.x, Z.y);
                            hits = push_hits2(hits, px, py, pz, lim, X.z, X.w, Y.z, Y.w, Z.z, Z.w);
                        }
                    }
                    // first slot to the top bit; drop the slots before the run and what was read past its end
                    hits = (hits << (32 - 4 * groups)) & (0xffffffffu << (32 - cnt)) & (0xffffffffu >> (b < start ? start - b : 0));
                    *tail = make_uint2(b, hits);
                    tail += hits ? GATHER_THREADS : 0;
                    if (tail == words_end) flush();
                }
            }
        }
    }
    flush();
}

// float4 iterate -> the cull's three coordinate arrays, every stored slot (ghosts included); ~8 us per
// million particles, HBM bound (16 B read, 12 B written per particle)
__global__ void __launch_bounds__(256)
pack_kernel(const float4* __restrict__ x, float* __restrict__ xs, float* __restrict__ ys, float* __restrict__ zs, int64_t n) {
    pdl_wait();   // (launch.cuh: nothing of the previous kernel is touched before this)
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 q = x[i];
    xs[i] = q.x;
    ys[i] = q.y;
    zs[i] = q.z;
}

// the same for the GHOST slots only — [0, own_first) and [own_first + own_count, n) — once the neighbours have
// reported that their pushes into them are complete (HaloSync; every block waits for both sides: the signal
// left the neighbours' kernels long ago, see the block order there)
__global__ void __launch_bounds__(256)
pack_ghosts_kernel(const float4* __restrict__ x, float* __restrict__ xs, float* __restrict__ ys, float* __restrict__ zs,
                   int64_t own_first, int64_t own_count, int64_t n, const __grid_constant__ HaloSync hs) {
    pdl_wait();
    if (threadIdx.x == 0) {
        if (hs.wait_left) halo_spin(hs.wait_left, hs.wait_seq, hs.timeout_ns, hs.flags);
        if (hs.wait_right) halo_spin(hs.wait_right, hs.wait_seq, hs.timeout_ns, hs.flags);
    }
    __syncthreads();
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i = k < own_first ? k : k + own_count;
    if (i >= n) return;
    const float4 q = x[i];
    xs[i] = q.x;
    ys[i] = q.y;
    zs[i] = q.z;
}

// ---- re-binning a block's particles by their CURRENT home cell -------------------------------------------
// The reference re-derives a particle's home cell from the current iterate (Simulator_kernel.cuh:70,148,212), so
// after the first Jacobi iteration 12-17 % of the particles per axis search the neighbourhood of a cell that is
// not their sorted cell any more. Threads of a warp that handle consecutive sorted slots then walk DIFFERENT
// runs, and one cull load of the warp touches ~8.5 cache lines instead of ~3 (ncu, profiles/r01h: the sweeps
// sit on the L1 wavefront rate). Which thread computes which particle is free, though: every particle is still
// accumulated by ONE thread in ascending slot order and written to its own slot, so no bit changes. The block
// therefore re-deals its GATHER_THREADS particles to its threads in the order of (current home cell, slot):
// lanes of a warp share home cells again, as in the first iteration.
// Rank by counting over 32-bit words (cell key relative to the block's smallest, clamped to 25 bits | local
// slot): ~290 instructions per thread against ~5000 of the sweep. A key beyond the clamp only groups worse.
// Returns the local index (0..GATHER_THREADS-1) of the particle this thread takes; all threads of the block call.
__device__ __forceinline__ uint32_t rebin_block(const float4* __restrict__ x, int64_t first, int64_t n, uint32_t block,
                                                const GridConsts& g, uint32_t* __restrict__ s_re /* GATHER_THREADS + 4 words */) {
    const uint32_t tid = threadIdx.x;
    const int64_t t = (int64_t)block * GATHER_THREADS + tid;
    uint32_t key = 0x3fffffffu;   // past the end of the range: sorts last
    if (t < n) {
        const float4 p = x[first + t];
        const int3 cc = cell_of(p.x, p.y, p.z, g);
        key = (uint32_t)(cc.x * g.dyz + cc.y * g.dim[2] + cc.z);
    }
    const uint32_t wmin = __reduce_min_sync(0xffffffffu, key);
    if ((tid & 31u) == 0) s_re[GATHER_THREADS + (tid >> 5)] = wmin;
    __syncthreads();
    uint32_t kmin = s_re[GATHER_THREADS];
#pragma unroll
    for (int w = 1; w < GATHER_THREADS / 32; w++) kmin = min(kmin, s_re[GATHER_THREADS + w]);
    const uint32_t v = (min(key - kmin, 0x1ffffffu) << 7) | tid;
    s_re[tid] = v;
    __syncthreads();
    uint32_t rank = 0;
    const uint4* s4 = reinterpret_cast<const uint4*>(s_re);
#pragma unroll 8
    for (int j = 0; j < GATHER_THREADS / 4; j++) {
        const uint4 q = s4[j];
        rank += (q.x < v) + (q.y < v) + (q.z < v) + (q.w < v);
    }
    __syncthreads();
    s_re[rank] = tid;
    __syncthreads();
    return s_re[tid];
}

// ---- the TMA A/B: the block's candidate coordinates staged in shared memory ---------------------
// North-star item 2 asks for the 27-cell neighbourhood staged in shared memory by TMA bulk copies of contiguous
// sorted cell ranges. In the FIRST Jacobi iteration a block's 128 particles sit in consecutive cells, so for each of
// the nine (dx, dy) runs the union of its threads' runs is one short contiguous slot range: thread 0 issues 27
// cp.async.bulk copies (9 runs x {xs, ys, zs}) that complete on one mbarrier, and the cull reads shared memory
// instead of global memory / L1 (gather<.., STAGED>). Returns false (all threads) when a union does not fit
// STAGE_CAP slots: the block then takes the plain path. s_u: 20 words (9 union starts, 9 union ends, ok, bytes).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool stage_runs(const float4* __restrict__ x, const CullSoA soa, const uint2* __restrict__ cell_range,
                                           const GridConsts& g, int64_t first, int64_t n, uint32_t block, float* s_xyz,
                                           uint32_t* s_u, unsigned long long* s_mbar) {
    const uint32_t tid = threadIdx.x;
    const int64_t t = (int64_t)block * GATHER_THREADS + tid;
    const bool valid = t < n;
    if (tid < 9) { s_u[tid] = 0xffffffffu; s_u[9 + tid] = 0u; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(s_mbar)));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the async proxy (TMA) sees the initialised barrier
    }
    __syncthreads();
    int3 cc = make_int3(0, 0, 0);
    if (valid) {
        const float4 p = x[first + t];
        cc = cell_of(p.x, p.y, p.z, g);
    }
    const bool has_below = cc.z > 0, has_above = cc.z + 1 < g.dim[2];
#pragma unroll
    for (int run = 0; run < 9; run++) {
        const int cx = cc.x + run / 3 - 1, cy = cc.y + run % 3 - 1, lx = cx - g.xoff;
        uint32_t lo = 0xffffffffu, hi = 0u;
        if (valid && cx >= 0 && cx < g.dim[0] && lx >= 0 && lx < g.nxl && cy >= 0 && cy < g.dim[1]) {
            const int cbase = lx * g.dyz + cy * g.dim[2];
            const uint2 zero = make_uint2(0u, 0u);
            const uint2 r0 = has_below ? __ldg(&cell_range[cbase + cc.z - 1]) : zero;
            const uint2 r1 = __ldg(&cell_range[cbase + cc.z]);
            const uint2 r2 = has_above ? __ldg(&cell_range[cbase + cc.z + 1]) : zero;
            const bool e0 = r0.y > r0.x, e1 = r1.y > r1.x, e2 = r2.y > r2.x;
            const uint32_t start = e0 ? r0.x : e1 ? r1.x : r2.x, end = e2 ? r2.y : e1 ? r1.y : r0.y;
            if (end > start) { lo = start; hi = end; }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if ((tid & 31u) == 0 && hi > lo) { atomicMin(&s_u[run], lo); atomicMax(&s_u[9 + run], hi); }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = 0, ok = 1;
        for (int run = 0; run < 9; run++) {
            if (s_u[9 + run] <= s_u[run]) continue;
            const uint32_t b = s_u[run] & ~3u, len = (s_u[9 + run] - b + 3u) & ~3u;   // whole groups of four slots
            if (len > (uint32_t)STAGE_CAP) ok = 0;
            bytes += 3u * len * 4u;
        }
        if (ok && bytes) {
            const uint32_t mb = smem_addr(s_mbar);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            for (int run = 0; run < 9; run++) {
                if (s_u[9 + run] <= s_u[run]) continue;
                const uint32_t b = s_u[run] & ~3u, nbytes = ((s_u[9 + run] - b + 3u) & ~3u) * 4u;
                const float* src[3] = {soa.xs + b, soa.ys + b, soa.zs + b};
                for (int a = 0; a < 3; a++)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_addr(s_xyz + ((size_t)run * 3 + a) * STAGE_CAP)), "l"(src[a]), "r"(nbytes), "r"(mb) : "memory");
                s_u[run] = b;   // what the cull subtracts
            }
        }
        s_u[18] = ok;
        s_u[19] = bytes;
    }
    __syncthreads();
    const bool ok = s_u[18] != 0;
    if (ok && s_u[19]) {
        const uint32_t mb = smem_addr(s_mbar);
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
    }
    return ok;
}

// ---- neighbour-list reuse between the two passes of one Jacobi iteration ------------------------
// The lambda and delta-p passes of an iteration read the SAME positions (the reference runs
// computeLambda and computetpos on the same dc_npos, Simulator.cu:222-245), so their in-range
// neighbour sets and the per-pair kernel values coincide. The lambda pass therefore saves, per
// particle and in visiting order, the slot of every in-range neighbour (itself excluded, as in
// computetpos) plus the one expensive value the delta-p pass needs from the pair geometry: the
// spiky scale s (sqrt + IEEE division). The delta-p pass replays the list: no cull, no sqrt, no
// division — the poly6 weight is four multiplies from the r2 it forms anyway, and w^n_corr it can
// afford (it is HBM bound) — and produces the same bits, because it consumes the very values its
// own evaluation would have produced.
// Layout (block b of 128 threads, entry k, thread t): [(b*PAIR_CAP + k)*128 + t] — a warp's k-th
// entries are contiguous 8-byte (slot, s) records. A particle with more than PAIR_CAP neighbours
// is flagged in its count word and handled by the delta-p pass's full gather instead.

template <bool SAVE_PAIRS, bool FAST_SPIKY, bool REBIN, bool STAGED = false, bool MORTON = false>
__global__ void __launch_bounds__(GATHER_THREADS, STAGED ? 6 : PBF_GATHER_MINBLOCKS)
lambda_kernel(const float4* __restrict__ x, const CullSoA soa, float4* __restrict__ xl, float* __restrict__ rho_out,
              const uint2* __restrict__ cell_range, int64_t first, int64_t n,
              uint2* __restrict__ pair_js, uint32_t* __restrict__ pair_cnt,
              const __grid_constant__ HaloPush hp, const __grid_constant__ HaloSync hs,
              const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    pdl_wait();
    extern __shared__ uint2 s_words[];
    __shared__ __align__(16) uint32_t s_re[REBIN ? GATHER_THREADS + 4 : 4];
    const uint32_t lb = halo_block(hs);   // (slab mode: the edge blocks first; blockIdx.x otherwise)
    const uint32_t local = REBIN ? rebin_block(x, first, n, lb, g, s_re) : threadIdx.x;   // which particle of the block
    // STAGED (first iteration only): the block's candidates by TMA bulk copies into shared memory, see stage_runs
    __shared__ __align__(16) float s_xyz[STAGED ? 9 * 3 * STAGE_CAP : 4];
    __shared__ uint32_t s_u[STAGED ? 20 : 1];
    __shared__ __align__(8) unsigned long long s_mbar;
    const bool staged = STAGED && stage_runs(x, soa, cell_range, g, first, n, lb, s_xyz, s_u, &s_mbar);
    const int64_t t = (int64_t)lb * GATHER_THREADS + local;
    if (t >= n) {   // a column of the last block without a particle: its word names a particle >= n, the replay skips it
        if (SAVE_PAIRS) pair_cnt[(int64_t)lb * GATHER_THREADS + threadIdx.x] = pair_word(0, local);
        return;
    }
    const int64_t i = first + t;
    const float4 p = x[i];
    float rho = 0.f, gradj_l2 = 0.f, gix = 0.f, giy = 0.f, giz = 0.f;
    // the particle itself: r2 = 0 adds poly6(0) to rho at its place in the visiting order and
    // nothing else (spiky is 0 below KERNAL_EPS, and gradj skips j == i). Taking it out of the
    // general path keeps three 0/rho0 divisions off IEEE division's slow path in every warp.
    const float w_self = poly6_in(0.f, c);
    const size_t pair0 = (size_t)lb * PAIR_CAP * GATHER_THREADS + threadIdx.x;
    int n_pairs = 0;
    auto heavy = [&](uint32_t j, float4 q, int) {
        if (j == (uint32_t)i) {
            rho = __fadd_rn(rho, w_self);
        } else {
            const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
            const float r2 = sumsq(dx, dy, dz);
            rho = __fadd_rn(rho, poly6(r2, c));
            const float s = FAST_SPIKY ? spiky_scale_fast(r2, c) : spiky_scale(r2, c);
            float gx = __fmul_rn(dx, s), gy = __fmul_rn(dy, s), gz = __fmul_rn(dz, s);
            div3_pho0(gx, gy, gz, c);
            gix = __fadd_rn(gix, gx);
            giy = __fadd_rn(giy, gy);
            giz = __fadd_rn(giz, gz);
            gradj_l2 = __fadd_rn(gradj_l2, sumsq(gx, gy, gz));
            // The delta-p pass needs s, poly6(r2) and its n_corr-th power for this pair. The lambda pass
            // is instruction-issue bound and the delta-p replay is HBM bound with idle issue slots, so
            // only s travels: the replay re-forms r2 from the same positions (the same bits) and
            // evaluates poly6 and the ~45-instruction powf there.
            if (SAVE_PAIRS) {
                if (n_pairs < PAIR_CAP) list_store(&pair_js[pair0 + (size_t)n_pairs * GATHER_THREADS], j, __float_as_uint(s));
                n_pairs++;
            }
        }
    };
    if (STAGED && staged)
        gather<false, true>(p, (uint32_t)i, c.h2_cull, x, soa, cell_range, g, s_words + threadIdx.x, heavy, smem_addr(s_xyz), s_u);
    else
        gather<false, false, MORTON>(p, (uint32_t)i, c.h2_cull, x, soa, cell_range, g, s_words + threadIdx.x, heavy);
    if (c.k_boundary != 0.f) rho = __fmaf_rn(c.k_boundary, boundary_density(p.x, p.y, p.z, g), rho);
    const float grad_l2 = __fmaf_rn(giz, giz, __fmaf_rn(giy, giy, __fmaf_rn(gix, gix, gradj_l2)));
    const float lambda = __fdiv_rn(-__fadd_rn(__fdiv_rn(rho, c.pho0), -1.f), __fadd_rn(grad_l2, c.lambda_eps));
    const float4 out = make_float4(p.x, p.y, p.z, lambda);
    xl[i] = out;
    halo_push(hp, t, out);
    rho_out[i] = rho;
    if (SAVE_PAIRS) {   // the list lives in THIS THREAD's column (coalesced records); the word says whose it is
        pair_cnt[(int64_t)lb * GATHER_THREADS + threadIdx.x] = pair_word(n_pairs, local);
    }
    halo_exit(hs, lb);
}

// The delta-p pass comes as two kernels. The REPLAY kernel walks the neighbour list the lambda pass saved:
// no cell table, no shared-memory list, few registers — it is latency / HBM bound, so it is compiled for
// 16 CTAs per SM instead of 8 (32 registers, no spills; measured 0.28 -> 0.20 ms in the compressed state). A particle whose list overflowed
// (more than PAIR_CAP neighbours) takes delta_p_one (solver_common.cuh) inside the replay kernel; the GATHER kernel
// below runs only when there is no list at all.
#ifndef PBF_REPLAY_MINBLOCKS
#define PBF_REPLAY_MINBLOCKS 16
#endif
// (the branch-free POW = 3 body lets the compiler overlap more pairs than 32 registers hold: unrolled by 4 it
//  spills two values per trip of four pairs — and is still the faster one: 2.456 vs 2.494 ms per step unrolled by 2)
#ifndef PBF_REPLAY_UNROLL3
#define PBF_REPLAY_UNROLL3 4
#endif
constexpr int REPLAY_UNROLL3 = PBF_REPLAY_UNROLL3;
// POW = 3: two list entries per trip through the two-lane fp32 instructions (pow4_trim2, pbf_math.cuh)
#ifndef PBF_REPLAY_PACKED
#define PBF_REPLAY_PACKED 1
#endif
#ifndef PBF_REPLAY_MINBLOCKS_PACKED
#define PBF_REPLAY_MINBLOCKS_PACKED 10
#endif
#ifndef PBF_REPLAY_UNROLL2
#define PBF_REPLAY_UNROLL2 1
#endif
// list records this many TRIPS (of two records) ahead are prefetched into L1: the list streams from HBM, and one trip
// of look-ahead in registers (nx0 / nx1 below) does not cover that latency — half of the replay's stall samples were
// waits for a record (profiles/r03_delta_p_ncu_per_instruction.txt). 0: off. Measured (dam_1m, ms per launch, early /
// compressed state): off 0.1324 / 0.1746, 2 trips 0.1283 / 0.1713, 4 trips 0.1291 / 0.1717, 8 trips 0.1320 / 0.1737,
// 4 trips into L2 only 0.1288 / 0.1717.
#ifndef PBF_REPLAY_PREFETCH
#define PBF_REPLAY_PREFETCH 2
#endif
constexpr int REPLAY_PREFETCH = PBF_REPLAY_PREFETCH;
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#ifndef PBF_REPLAY_PREFETCH_L2
#define PBF_REPLAY_PREFETCH_L2 0
#endif
constexpr bool REPLAY_PACKED = PBF_REPLAY_PACKED != 0;
constexpr int REPLAY_UNROLL2 = PBF_REPLAY_UNROLL2;
// (the packed POW = 3 loop keeps two pairs in flight: 48 registers, 10 CTAs per SM — measured against 16 (spills),
//  12 (spills) and 8: 0.1360 ms per launch vs 0.1626 / 0.1422 / 0.1704, scalar loop at 16 CTAs 0.1483; dam_1m, early state)
template <int POW>
__global__ void __launch_bounds__(GATHER_THREADS, ((POW == 3 || POW == 0) && REPLAY_PACKED) ? PBF_REPLAY_MINBLOCKS_PACKED : PBF_REPLAY_MINBLOCKS)
delta_p_replay_kernel(const float4* __restrict__ xl, float4* __restrict__ x_out, const CullOut co, int64_t first, int64_t n,
                      const uint2* __restrict__ pair_js,
                      const uint32_t* __restrict__ pair_cnt, const uint2* __restrict__ cell_range,
                      const __grid_constant__ HaloPush hp, const __grid_constant__ HaloSync hs,
                      const __grid_constant__ VelTail vt, const __grid_constant__ GridConsts g,
                      const __grid_constant__ SolverConsts c, const uint32_t block0) {
    pdl_wait();
    // (block0: this launch is a SLICE of the pass, its blocks are block0 .. of the whole — pbf_step_host runs the
    //  last pass in slices so that the final positions of a finished slice go home while the next one is computed)
    const uint32_t lb = halo_block(hs) + block0;
    halo_enter(hs, lb);   // (edge blocks: the neighbours' lambdas of this iteration are in the ghost slots of xl)
    const int64_t col = (int64_t)lb * GATHER_THREADS + threadIdx.x;   // list column; its word names the particle
    const uint32_t cw = pair_cnt[col];   // (the lambda pass writes the word of EVERY column of its blocks)
    const int64_t t = (int64_t)lb * GATHER_THREADS + pair_local(cw);
    if (t >= n) return;   // a column of the last block that holds no particle (the paired sweeps' columns are not 0..m-1)
    const int64_t i = first + t;
    float4 out;
    if (cw & PAIR_OVERFLOW) {   // more neighbours than the list holds: the plain pass for this one particle
        out = delta_p_one<POW>(xl, (uint32_t)i, cell_range, g, c);
    } else {
        const uint32_t cnt = pair_count(cw);
        const float4 p = xl[i];
        float ax = 0.f, ay = 0.f, az = 0.f;
        const size_t pair0 = (size_t)lb * PAIR_CAP * GATHER_THREADS + threadIdx.x;
        if ((POW == 3 || POW == 0) && REPLAY_PACKED) {
            // two list entries per trip, their arithmetic in the two lanes of the packed fp32 instructions (each lane
            // the scalar IEEE operation, pow4_trim2); the sums are still accumulated one pair after the other, in
            // list order. An odd list ends with its last entry in both lanes, the second one not accumulated.
            const f32x2 px = splat2(p.x), py = splat2(p.y), pz = splat2(p.z), pw_ = splat2(p.w);
            const f32x2 h2 = splat2(c.h2), coef = splat2(c.poly6_coef), corr = splat2(c.coef_corr);
            // (the records of the NEXT trip are requested before this trip's gathers: record -> gather -> arithmetic
            //  is a chain of two long-latency loads otherwise)
            uint2 nx0 = make_uint2(0u, 0u), nx1 = nx0;
            if (cnt > 0) {
                nx0 = list_load(&pair_js[pair0]);
                nx1 = list_load(&pair_js[pair0 + (size_t)(cnt > 1 ? 1 : 0) * GATHER_THREADS]);
            }
#pragma unroll REPLAY_UNROLL2
            for (uint32_t k = 0; k < cnt; k += 2) {
                const bool two = k + 1 < cnt;
                const uint2 js0 = nx0, js1 = nx1;
                if (k + 2 < cnt) {
                    nx0 = list_load(&pair_js[pair0 + (size_t)(k + 2) * GATHER_THREADS]);
                    nx1 = list_load(&pair_js[pair0 + (size_t)(k + 3 < cnt ? k + 3 : k + 2) * GATHER_THREADS]);
                }
                if (REPLAY_PREFETCH > 0 && k + 2 * REPLAY_PREFETCH + 2 < cnt) {   // (two rows of the list: one line each per 16 lanes)
                    const uint2* pf = &pair_js[pair0 + (size_t)(k + 2 * REPLAY_PREFETCH + 2) * GATHER_THREADS];
                    if (PBF_REPLAY_PREFETCH_L2) { prefetch_l2(pf); prefetch_l2(pf + GATHER_THREADS); }
                    else { prefetch_l1(pf); prefetch_l1(pf + GATHER_THREADS); }
                }
                const float4 q0 = __ldg(&xl[js0.x]), q1 = __ldg(&xl[js1.x]);
                const f32x2 dx = sub2(px, pack2(q0.x, q1.x)), dy = sub2(py, pack2(q0.y, q1.y)), dz = sub2(pz, pack2(q0.z, q1.z));
                const f32x2 r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
                const f32x2 t = sub2(h2, r2);
                float w0, w1, r20, r21;
                unpack2(mul2(mul2(mul2(coef, t), t), t), w0, w1);      // poly6_in ...
                unpack2(r2, r20, r21);
                w0 = r20 >= c.h2 ? 0.f : w0;                           // ... and poly6's range test
                w1 = r21 >= c.h2 ? 0.f : w1;
                const f32x2 w2 = pack2(w0, w1), ww = mul2(w2, w2);
                const f32x2 sc = fma2(corr, POW == 3 ? pow4_trim2(w2) : mul2(ww, ww), add2(pw_, pack2(q0.w, q1.w)));
                const f32x2 sj = pack2(__uint_as_float(js0.y), __uint_as_float(js1.y));
                float sc0, sc1, tx0, tx1, ty0, ty1, tz0, tz1;
                unpack2(sc, sc0, sc1);
                unpack2(mul2(dx, sj), tx0, tx1);
                unpack2(mul2(dy, sj), ty0, ty1);
                unpack2(mul2(dz, sj), tz0, tz1);
                ax = __fmaf_rn(sc0, tx0, ax);
                ay = __fmaf_rn(sc0, ty0, ay);
                az = __fmaf_rn(sc0, tz0, az);
                if (two) {
                    ax = __fmaf_rn(sc1, tx1, ax);
                    ay = __fmaf_rn(sc1, ty1, ay);
                    az = __fmaf_rn(sc1, tz1, az);
                }
            }
        } else {
#pragma unroll (POW == 3 ? REPLAY_UNROLL3 : 4)
            for (uint32_t k = 0; k < cnt; k++) {
                const size_t e = pair0 + (size_t)k * GATHER_THREADS;
                const uint2 js = list_load(&pair_js[e]);
                const float4 q = __ldg(&xl[js.x]);
                const float sj = __uint_as_float(js.y);  // spiky scale of the pair, saved by the lambda pass
                const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
                const float pw = pow_ncorr<POW>(poly6(sumsq(dx, dy, dz), c), c);
                const float sc = __fmaf_rn(c.coef_corr, pw, __fadd_rn(p.w, q.w));
                ax = __fmaf_rn(sc, __fmul_rn(dx, sj), ax);
                ay = __fmaf_rn(sc, __fmul_rn(dy, sj), ay);
                az = __fmaf_rn(sc, __fmul_rn(dz, sj), az);
            }
        }
        out = delta_p_finish(p, ax, ay, az, c);
    }
    x_out[i] = out;
    co.store(i, out);
    halo_push(hp, t, out);
    if (vt.v4) velocity_tail(vt, t, i, out);   // the step's last pass: the velocity update rides along
    halo_exit(hs, lb);
}

// The delta-p pass WITHOUT a neighbour list (the list is optional scratch: PBF_NO_PAIR_REUSE=1, or a handle too
// large for it): the full two-phase gather for every particle.
template <int POW, bool REBIN, bool MORTON = false>
__global__ void __launch_bounds__(GATHER_THREADS, PBF_GATHER_MINBLOCKS)
delta_p_kernel(const float4* __restrict__ xl, const CullSoA soa, float4* __restrict__ x_out, const CullOut co,
               const uint2* __restrict__ cell_range, int64_t first, int64_t n, const __grid_constant__ HaloPush hp,
               const __grid_constant__ HaloSync hs, const __grid_constant__ VelTail vt,
               const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    pdl_wait();
    extern __shared__ uint2 s_words[];
    __shared__ __align__(16) uint32_t s_re[REBIN ? GATHER_THREADS + 4 : 4];
    const uint32_t lb = halo_block(hs);
    halo_enter(hs, lb);
    const uint32_t local = REBIN ? rebin_block(xl, first, n, lb, g, s_re) : threadIdx.x;
    const int64_t t = (int64_t)lb * GATHER_THREADS + local;
    if (t >= n) return;
    const int64_t i = first + t;
    const float4 p = xl[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    gather<true, false, MORTON>(p, (uint32_t)i, c.h2_cull, xl, soa, cell_range, g, s_words + threadIdx.x, [&](uint32_t, float4 q, int) {
        const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
        const float r2 = sumsq(dx, dy, dz);
        const float pw = pow_ncorr<POW>(poly6(r2, c), c);
        const float sc = __fmaf_rn(c.coef_corr, pw, __fadd_rn(p.w, q.w));
        const float s = spiky_scale(r2, c);
        ax = __fmaf_rn(sc, __fmul_rn(dx, s), ax);
        ay = __fmaf_rn(sc, __fmul_rn(dy, s), ay);
        az = __fmaf_rn(sc, __fmul_rn(dz, s), az);
    });
    const float4 out = delta_p_finish(p, ax, ay, az, c);
    x_out[i] = out;
    co.store(i, out);   // (reads of this iteration's coordinates go to `soa` = the OTHER set of arrays)
    halo_push(hp, t, out);
    if (vt.v4) velocity_tail(vt, t, i, out);
    halo_exit(hs, lb);
}

// vel = (npos - pos) * inv_dt, plus everything the caller-facing buffers need from this point:
// pos <- step-input position (parked in npos by the reorder pass), npos <- final iterate.
__global__ void __launch_bounds__(256)
update_velocity_kernel(const float4* __restrict__ x, const float* __restrict__ rho,
                       float* __restrict__ pos_out, float* __restrict__ npos_io,
                       float* __restrict__ vel_out, float4* __restrict__ v4, int64_t first, int64_t n,
                       const __grid_constant__ HaloPush hp, const __grid_constant__ HaloSync hs,
                       const __grid_constant__ StatePush sp, const __grid_constant__ SolverConsts c) {
    pdl_wait();
    const uint32_t lb = halo_block(hs);
    halo_enter(hs, lb);   // (edge blocks: the neighbours are done with what this kernel's pushes overwrite)
    const int64_t t = (int64_t)lb * 256 + threadIdx.x;
    if (t >= n) return;
    const int64_t i = first + t;
    const float4 q = x[i];
    const float3 p0 = load_f3(npos_io, t);
    const float vx = __fmul_rn(__fsub_rn(q.x, p0.x), c.inv_dt);
    const float vy = __fmul_rn(__fsub_rn(q.y, p0.y), c.inv_dt);
    const float vz = __fmul_rn(__fsub_rn(q.z, p0.z), c.inv_dt);
    const float4 out = make_float4(vx, vy, vz, rho[i]);
    v4[i] = out;
    halo_push(hp, t, out);
    store_f3(vel_out, t, vx, vy, vz);
    store_f3(pos_out, t, p0.x, p0.y, p0.z);
    store_f3(npos_io, t, q.x, q.y, q.z);
    push_state_pos(sp, t, q.x, q.y, q.z);
    halo_exit(hs, lb);
}

template <bool REBIN, bool MORTON = false>
__global__ void __launch_bounds__(GATHER_THREADS, PBF_GATHER_MINBLOCKS)
xsph_kernel(const float4* __restrict__ x, const CullSoA soa, const float4* __restrict__ v4,
            const uint2* __restrict__ cell_range, float* __restrict__ nvel_out,
            const uint32_t* __restrict__ iid_sorted, uint32_t* __restrict__ iid_out, int64_t first, int64_t n,
            const __grid_constant__ HaloSync hs, const __grid_constant__ StatePush sp, const __grid_constant__ GridConsts g,
            const __grid_constant__ SolverConsts c) {
    pdl_wait();
    extern __shared__ uint2 s_words[];
    __shared__ __align__(16) uint32_t s_re[REBIN ? GATHER_THREADS + 4 : 4];
    const uint32_t lb = halo_block(hs);
    halo_enter(hs, lb);   // (edge blocks: the neighbours' velocities are in the ghost slots of v4)
    const uint32_t local = REBIN ? rebin_block(x, first, n, lb, g, s_re) : threadIdx.x;
    const int64_t t = (int64_t)lb * GATHER_THREADS + local;
    if (t >= n) return;
    const int64_t i = first + t;
    const float4 p = x[i];
    const float4 vi = v4[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    // The particle itself (computeXSPH visits it) contributes 0 / (2 rho_i) = +0 per component, and an
    // accumulator that starts at +0 never becomes -0 (x + y = -0 only for x = y = -0), so adding +0 changes
    // no bit: it is skipped, which keeps three zero dividends off IEEE division's slow path in every warp.
    gather<true, false, MORTON>(p, (uint32_t)i, c.h2, x, soa, cell_range, g, s_words + threadIdx.x, [&](uint32_t j, float4 q, int) {
        const float r2 = sumsq(__fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z));
        const float4 vj = __ldg(&v4[j]);
        const float w = poly6_in(r2, c);
        const float den = __fadd_rn(vi.w, vj.w);
        const float tx = __fsub_rn(vj.x, vi.x), ty = __fsub_rn(vj.y, vi.y), tz = __fsub_rn(vj.z, vi.z);
        ax = __fadd_rn(ax, __fdiv_rn(__fmul_rn(__fadd_rn(tx, tx), w), den));
        ay = __fadd_rn(ay, __fdiv_rn(__fmul_rn(__fadd_rn(ty, ty), w), den));
        az = __fadd_rn(az, __fdiv_rn(__fmul_rn(__fadd_rn(tz, tz), w), den));
    });
    const float ox = __fmaf_rn(c.c_xsph, ax, vi.x), oy = __fmaf_rn(c.c_xsph, ay, vi.y), oz = __fmaf_rn(c.c_xsph, az, vi.z);
    const uint32_t id = iid_sorted[i];
    store_f3(nvel_out, t, ox, oy, oz);
    iid_out[t] = id;
    push_state_vel(sp, t, ox, oy, oz, id);
}

// ---- two particles per thread: the paired sweeps (PBF_OPT_PAIRED) ---------------------------------------------------
// The cull is where the sweeps sit on the L1 data pipe: every thread streams its particle's ~216-390 candidates
// through three 16-byte loads per four slots, and two particles in neighbouring slots stream almost the same
// candidates. A thread of the paired kernels therefore takes TWO consecutive slots A and B and walks the UNION of
// their candidate runs once: one set of loads feeds both particles' tests (half the L1 wavefronts, a quarter fewer
// instructions per test). Each particle keeps its own hit words, masked to ITS run ([start, end) of the three cells
// around ITS current home cell in that column), so each still sees exactly its candidates in ascending slot order,
// and each is accumulated by this one thread in that order: no bit changes. Columns are visited over the bounding
// box of the two 3 x 3 column neighbourhoods in ascending (x, y) = ascending slot order; a column only one of the
// two searches is culled for that one alone. A pair whose home cells are far apart (a slot pair that straddles the
// end of a z column) is simply done one after the other.
constexpr int PAIR_THREADS = GATHER_THREADS / 2;   // 64 threads = 128 particles per block, like the other sweeps:
                                                   // the neighbour list's layout and the halo's edge blocks stay
#ifndef PBF_PAIRED_MINBLOCKS
#define PBF_PAIRED_MINBLOCKS 10
#endif

struct SlotRun { uint32_t start, end; };
// the run of a column for a particle whose home cell has z index cz: first slot of the first non-empty cell of
// (cz - 1, cz, cz + 1) to the end of the last non-empty one; {0, 0} if all are empty (see gather())
__device__ __forceinline__ SlotRun column_run(const uint2* __restrict__ cell_range, int cbase, int cz, int dimz) {
    const uint2 zero = make_uint2(0u, 0u);
    const uint2 r0 = cz > 0 ? __ldg(&cell_range[cbase + cz - 1]) : zero;
    const uint2 r1 = __ldg(&cell_range[cbase + cz]);
    const uint2 r2 = cz + 1 < dimz ? __ldg(&cell_range[cbase + cz + 1]) : zero;
    const bool e0 = r0.y > r0.x, e1 = r1.y > r1.x, e2 = r2.y > r2.x;
    SlotRun r;
    r.start = e0 ? r0.x : e1 ? r1.x : r2.x;
    r.end = e2 ? r2.y : e1 ? r1.y : r0.y;
    return r;
}
// bits of the 32-slot word that starts at slot b (first slot in the top bit) whose slots lie in [s, e)
__device__ __forceinline__ uint32_t word_mask(uint32_t b, uint32_t s, uint32_t e) {
    if (e <= b || s >= b + 32u || e <= s) return 0u;
    const uint32_t lo = s > b ? s - b : 0u;          // 0..31
    const uint32_t hi = min(e - b, 32u);             // 1..32
    return (0xffffffffu >> lo) & (0xffffffffu << (32u - hi));
}
__device__ __forceinline__ uint32_t push_hits2p(uint32_t hits, f32x2 px, f32x2 py, f32x2 pz, f32x2 lim, f32x2 cx, f32x2 cy, f32x2 cz) {
    const f32x2 dx = sub2(px, cx), dy = sub2(py, cy), dz = sub2(pz, cz);
    const f32x2 t = sub2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), lim);
    uint32_t t0, t1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(t));
    hits = __funnelshift_l(t0, hits, 1);
    return __funnelshift_l(t1, hits, 1);
}
// phase 2 of one particle: the exact arithmetic over the set bits of its list, in slot order (see gather())
template <bool SKIP_SELF, typename Heavy>
__device__ __forceinline__ void drain_words(const uint2* col, const uint2* tail, uint32_t self, const float4* __restrict__ x, Heavy& heavy) {
    const uint2* e = col;
    uint32_t first = 0, hits = 0;
    for (;;) {
        if (hits == 0) {
            if (e == tail) break;
            const uint2 w = *e;
            e += GATHER_THREADS;
            first = w.x;
            hits = w.y;
        }
        const int lead = __clz((int)hits);
        hits &= ~(0x80000000u >> lead);
        const uint32_t j = first + (uint32_t)lead;
        if (SKIP_SELF && j == self) continue;
        heavy(j, __ldg(&x[j]));
    }
}

// wordsA / wordsB: the two particles' columns of the block's shared-memory list (entry k at [k * GATHER_THREADS])
template <bool SKIP_SELF, typename HeavyA, typename HeavyB>
__device__ __forceinline__ void gather2(const float4 pA, const float4 pB, const bool validB, const uint32_t selfA, const uint32_t selfB,
                                        const float limit, const float4* __restrict__ x, const CullSoA soa,
                                        const uint2* __restrict__ cell_range, const GridConsts& g,
                                        uint2* __restrict__ wordsA, uint2* __restrict__ wordsB, HeavyA& heavyA, HeavyB& heavyB) {
    const int3 ca = cell_of(pA.x, pA.y, pA.z, g);
    const int3 cb = validB ? cell_of(pB.x, pB.y, pB.z, g) : ca;
    // one walk over the union of the two neighbourhoods when they are close (runs of three cells up to three cells
    // apart in z still form one gap-free slot range), else A's walk, then B's
    const bool joint = validB && abs(ca.x - cb.x) <= 2 && abs(ca.y - cb.y) <= 2 && abs(ca.z - cb.z) <= 3;
    uint2* tailA = wordsA;
    uint2* tailB = wordsB;
    uint2* const endA = wordsA + WORD_CAP * GATHER_THREADS;
    uint2* const endB = wordsB + WORD_CAP * GATHER_THREADS;
    const f32x2 pxA = pack2(pA.x, pA.x), pyA = pack2(pA.y, pA.y), pzA = pack2(pA.z, pA.z);
    const f32x2 pxB = pack2(pB.x, pB.x), pyB = pack2(pB.y, pB.y), pzB = pack2(pB.z, pB.z);
    const f32x2 lim = pack2(limit, limit);
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const bool actA = joint || pass == 0, actB = validB && (joint || pass == 1);
        if (pass == 1 && (joint || !validB)) break;
        const int xlo = actA && actB ? min(ca.x, cb.x) : actA ? ca.x : cb.x, xhi = actA && actB ? max(ca.x, cb.x) : actA ? ca.x : cb.x;
        const int ylo = actA && actB ? min(ca.y, cb.y) : actA ? ca.y : cb.y, yhi = actA && actB ? max(ca.y, cb.y) : actA ? ca.y : cb.y;
#pragma unroll 1
        for (int cx = xlo - 1; cx <= xhi + 1; cx++) {
            const int lx = cx - g.xoff;
            if (cx < 0 || cx >= g.dim[0]) continue;
            const bool ax = actA && abs(cx - ca.x) <= 1, bx = actB && abs(cx - cb.x) <= 1;
            if (!(ax || bx)) continue;
            if (lx < 0 || lx >= g.nxl) {   // the search leaves the stored planes (slab mode): say so, see gather()
                if (g.flags) atomicOr(g.flags, (uint32_t)PBF_SLAB_FLAG_GHOST);
                continue;
            }
#pragma unroll 1
            for (int cy = ylo - 1; cy <= yhi + 1; cy++) {
                if (cy < 0 || cy >= g.dim[1]) continue;
                const bool a = ax && abs(cy - ca.y) <= 1, b = bx && abs(cy - cb.y) <= 1;
                if (!(a || b)) continue;
                const int cbase = lx * g.dyz + cy * g.dim[2];
                SlotRun ra = {0u, 0u}, rb = {0u, 0u};
                if (a) ra = column_run(cell_range, cbase, ca.z, g.dim[2]);
                if (b) rb = (a && cb.z == ca.z) ? ra : column_run(cell_range, cbase, cb.z, g.dim[2]);
                const bool ea = ra.end > ra.start, eb = rb.end > rb.start;
                if (!(ea || eb)) continue;
                const uint32_t start = ea && eb ? min(ra.start, rb.start) : ea ? ra.start : rb.start;
                const uint32_t end = ea && eb ? max(ra.end, rb.end) : ea ? ra.end : rb.end;
#pragma unroll 1
                for (uint32_t w0 = start & ~3u; w0 < end; w0 += 32) {   // words start at multiples of four slots
                    const uint32_t cnt = min(end - w0, 32u);
                    const uint32_t groups = (cnt + 3) >> 2;
                    const float4* xp = reinterpret_cast<const float4*>(soa.xs + w0);
                    const float4* yp = reinterpret_cast<const float4*>(soa.ys + w0);
                    const float4* zp = reinterpret_cast<const float4*>(soa.zs + w0);
                    uint32_t hA = 0, hB = 0;
                    if (ea && eb) {
#pragma unroll 1
                        for (uint32_t gi = 0; gi < groups; gi++) {   // four candidates, both particles: 3 loads, 28 packed flops, 8 shifts
                            const float4 X = __ldg(xp + gi), Y = __ldg(yp + gi), Z = __ldg(zp + gi);
                            const f32x2 x01 = pack2(X.x, X.y), y01 = pack2(Y.x, Y.y), z01 = pack2(Z.x, Z.y);
                            const f32x2 x23 = pack2(X.z, X.w), y23 = pack2(Y.z, Y.w), z23 = pack2(Z.z, Z.w);
                            hA = push_hits2p(hA, pxA, pyA, pzA, lim, x01, y01, z01);
                            hB = push_hits2p(hB, pxB, pyB, pzB, lim, x01, y01, z01);
                            hA = push_hits2p(hA, pxA, pyA, pzA, lim, x23, y23, z23);
                            hB = push_hits2p(hB, pxB, pyB, pzB, lim, x23, y23, z23);
                        }
                    } else {
                        const f32x2 px = ea ? pxA : pxB, py = ea ? pyA : pyB, pz = ea ? pzA : pzB;
                        uint32_t h = 0;
#pragma unroll 1
                        for (uint32_t gi = 0; gi < groups; gi++) {
                            const float4 X = __ldg(xp + gi), Y = __ldg(yp + gi), Z = __ldg(zp + gi);
                            h = push_hits2p(h, px, py, pz, lim, pack2(X.x, X.y), pack2(Y.x, Y.y), pack2(Z.x, Z.y));
                            h = push_hits2p(h, px, py, pz, lim, pack2(X.z, X.w), pack2(Y.z, Y.w), pack2(Z.z, Z.w));
                        }
                        hA = ea ? h : 0u;
                        hB = ea ? 0u : h;
                    }
                    // first slot to the top bit; keep each particle's own run only
                    const uint32_t sh = 32 - 4 * groups;
                    hA = (hA << sh) & word_mask(w0, ra.start, ra.end);
                    hB = (hB << sh) & word_mask(w0, rb.start, rb.end);
                    *tailA = make_uint2(w0, hA);
                    tailA += hA ? GATHER_THREADS : 0;
                    *tailB = make_uint2(w0, hB);
                    tailB += hB ? GATHER_THREADS : 0;
                    if (tailA == endA) { drain_words<SKIP_SELF>(wordsA, tailA, selfA, x, heavyA); tailA = wordsA; }
                    if (tailB == endB) { drain_words<SKIP_SELF>(wordsB, tailB, selfB, x, heavyB); tailB = wordsB; }
                }
            }
        }
    }
    drain_words<SKIP_SELF>(wordsA, tailA, selfA, x, heavyA);
    if (validB) drain_words<SKIP_SELF>(wordsB, tailB, selfB, x, heavyB);
}

// one particle's sums of the lambda pass (computeLambda, Simulator_kernel.cuh:52-129), see lambda_kernel
template <bool SAVE_PAIRS, bool FAST_SPIKY>
struct LambdaAcc {
    float4 p;
    uint32_t self;
    float rho = 0.f, gradj_l2 = 0.f, gix = 0.f, giy = 0.f, giz = 0.f;
    int n_pairs = 0;
    float w_self;
    uint2* pair_col;   // this particle's column of the neighbour list (entry k at [k * GATHER_THREADS])
    const SolverConsts& c;
    __device__ __forceinline__ LambdaAcc(const float4 p_, uint32_t self_, uint2* col, const SolverConsts& c_)
        : p(p_), self(self_), w_self(poly6_in(0.f, c_)), pair_col(col), c(c_) {}
    __device__ __forceinline__ void operator()(uint32_t j, const float4 q) {
        if (j == self) {
            rho = __fadd_rn(rho, w_self);
        } else {
            const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
            const float r2 = sumsq(dx, dy, dz);
            rho = __fadd_rn(rho, poly6(r2, c));
            const float s = FAST_SPIKY ? spiky_scale_fast(r2, c) : spiky_scale(r2, c);
            float gx = __fmul_rn(dx, s), gy = __fmul_rn(dy, s), gz = __fmul_rn(dz, s);
            div3_pho0(gx, gy, gz, c);
            gix = __fadd_rn(gix, gx);
            giy = __fadd_rn(giy, gy);
            giz = __fadd_rn(giz, gz);
            gradj_l2 = __fadd_rn(gradj_l2, sumsq(gx, gy, gz));
            if (SAVE_PAIRS) {
                if (n_pairs < PAIR_CAP) list_store(&pair_col[(size_t)n_pairs * GATHER_THREADS], j, __float_as_uint(s));
                n_pairs++;
            }
        }
    }
    __device__ __forceinline__ float4 finish(float& rho_out, const GridConsts& g) {
        if (c.k_boundary != 0.f) rho = __fmaf_rn(c.k_boundary, boundary_density(p.x, p.y, p.z, g), rho);
        const float grad_l2 = __fmaf_rn(giz, giz, __fmaf_rn(giy, giy, __fmaf_rn(gix, gix, gradj_l2)));
        const float lambda = __fdiv_rn(-__fadd_rn(__fdiv_rn(rho, c.pho0), -1.f), __fadd_rn(grad_l2, c.lambda_eps));
        rho_out = rho;
        return make_float4(p.x, p.y, p.z, lambda);
    }
};

// Thread u of block lb takes the block's particles 2u (A) and 2u + 1 (B); their list columns are u and 64 + u (the
// k-th records of a warp's particles stay contiguous), and the count word of a column names its particle, which is
// all the delta-p replay needs to know (pair_word).
template <bool SAVE_PAIRS, bool FAST_SPIKY>
__global__ void __launch_bounds__(PAIR_THREADS, PBF_PAIRED_MINBLOCKS)
lambda2_kernel(const float4* __restrict__ x, const CullSoA soa, float4* __restrict__ xl, float* __restrict__ rho_out,
               const uint2* __restrict__ cell_range, int64_t first, int64_t n,
               uint2* __restrict__ pair_js, uint32_t* __restrict__ pair_cnt,
               const __grid_constant__ HaloPush hp, const __grid_constant__ HaloSync hs,
               const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    pdl_wait();
    extern __shared__ uint2 s_words[];
    const uint32_t lb = halo_block(hs);
    const uint32_t colA = threadIdx.x, colB = PAIR_THREADS + threadIdx.x;
    const int64_t tA = (int64_t)lb * GATHER_THREADS + 2 * threadIdx.x, tB = tA + 1;
    if (tA >= n) {   // columns without a particle: their words name particles >= n, the replay skips them
        if (SAVE_PAIRS) {
            pair_cnt[(int64_t)lb * GATHER_THREADS + colA] = pair_word(0, 2 * threadIdx.x);
            pair_cnt[(int64_t)lb * GATHER_THREADS + colB] = pair_word(0, 2 * threadIdx.x + 1);
        }
        return;
    }
    const bool validB = tB < n;
    const int64_t iA = first + tA, iB = iA + 1;
    const float4 pA = x[iA], pB = validB ? x[iB] : pA;
    uint2* const list0 = SAVE_PAIRS ? pair_js + (size_t)lb * PAIR_CAP * GATHER_THREADS : nullptr;
    LambdaAcc<SAVE_PAIRS, FAST_SPIKY> accA(pA, (uint32_t)iA, list0 + colA, c), accB(pB, (uint32_t)iB, list0 + colB, c);
    gather2<false>(pA, pB, validB, (uint32_t)iA, (uint32_t)iB, c.h2_cull, x, soa, cell_range, g, s_words + colA, s_words + colB, accA, accB);
    float rho;
    const float4 outA = accA.finish(rho, g);
    xl[iA] = outA;
    halo_push(hp, tA, outA);
    rho_out[iA] = rho;
    if (validB) {
        const float4 outB = accB.finish(rho, g);
        xl[iB] = outB;
        halo_push(hp, tB, outB);
        rho_out[iB] = rho;
    }
    if (SAVE_PAIRS) {
        const int64_t c0 = (int64_t)lb * GATHER_THREADS;
        pair_cnt[c0 + colA] = pair_word(accA.n_pairs, 2 * threadIdx.x);
        pair_cnt[c0 + colB] = pair_word(validB ? accB.n_pairs : 0, 2 * threadIdx.x + 1);
    }
    halo_exit(hs, lb);
}

struct XsphAcc {
    float4 p, vi;
    float ax = 0.f, ay = 0.f, az = 0.f;
    const float4* __restrict__ v4;
    const SolverConsts& c;
    __device__ __forceinline__ XsphAcc(const float4 p_, const float4 vi_, const float4* v4_, const SolverConsts& c_) : p(p_), vi(vi_), v4(v4_), c(c_) {}
    __device__ __forceinline__ void operator()(uint32_t j, const float4 q) {
        const float r2 = sumsq(__fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z));
        const float4 vj = __ldg(&v4[j]);
        const float w = poly6_in(r2, c);
        const float den = __fadd_rn(vi.w, vj.w);
        const float tx = __fsub_rn(vj.x, vi.x), ty = __fsub_rn(vj.y, vi.y), tz = __fsub_rn(vj.z, vi.z);
        ax = __fadd_rn(ax, __fdiv_rn(__fmul_rn(__fadd_rn(tx, tx), w), den));
        ay = __fadd_rn(ay, __fdiv_rn(__fmul_rn(__fadd_rn(ty, ty), w), den));
        az = __fadd_rn(az, __fdiv_rn(__fmul_rn(__fadd_rn(tz, tz), w), den));
    }
};

__global__ void __launch_bounds__(PAIR_THREADS, PBF_PAIRED_MINBLOCKS)
xsph2_kernel(const float4* __restrict__ x, const CullSoA soa, const float4* __restrict__ v4,
             const uint2* __restrict__ cell_range, float* __restrict__ nvel_out,
             const uint32_t* __restrict__ iid_sorted, uint32_t* __restrict__ iid_out, int64_t first, int64_t n,
             const __grid_constant__ HaloSync hs, const __grid_constant__ StatePush sp, const __grid_constant__ GridConsts g,
             const __grid_constant__ SolverConsts c) {
    pdl_wait();
    extern __shared__ uint2 s_words[];
    const uint32_t lb = halo_block(hs);
    halo_enter(hs, lb);   // (edge blocks: the neighbours' velocities are in the ghost slots of v4)
    const int64_t tA = (int64_t)lb * GATHER_THREADS + 2 * threadIdx.x, tB = tA + 1;
    if (tA >= n) return;
    const bool validB = tB < n;
    const int64_t iA = first + tA, iB = iA + 1;
    const float4 pA = x[iA], pB = validB ? x[iB] : pA;
    XsphAcc accA(pA, v4[iA], v4, c), accB(pB, validB ? v4[iB] : v4[iA], v4, c);
    gather2<true>(pA, pB, validB, (uint32_t)iA, (uint32_t)iB, c.h2, x, soa, cell_range, g, s_words + threadIdx.x,
                  s_words + PAIR_THREADS + threadIdx.x, accA, accB);
    {
        const float ox = __fmaf_rn(c.c_xsph, accA.ax, accA.vi.x), oy = __fmaf_rn(c.c_xsph, accA.ay, accA.vi.y), oz = __fmaf_rn(c.c_xsph, accA.az, accA.vi.z);
        const uint32_t id = iid_sorted[iA];
        store_f3(nvel_out, tA, ox, oy, oz);
        iid_out[tA] = id;
        push_state_vel(sp, tA, ox, oy, oz, id);
    }
    if (validB) {
        const float ox = __fmaf_rn(c.c_xsph, accB.ax, accB.vi.x), oy = __fmaf_rn(c.c_xsph, accB.ay, accB.vi.y), oz = __fmaf_rn(c.c_xsph, accB.az, accB.vi.z);
        const uint32_t id = iid_sorted[iB];
        store_f3(nvel_out, tB, ox, oy, oz);
        iid_out[tB] = id;
        push_state_vel(sp, tB, ox, oy, oz, id);
    }
}

__global__ void __launch_bounds__(GATHER_THREADS)
neighbor_count_kernel(const float4* __restrict__ x, const CullSoA soa, const uint2* __restrict__ cell_range,
                      uint32_t* __restrict__ count, int64_t n, const __grid_constant__ GridConsts g,
                      const __grid_constant__ SolverConsts c) {
    pdl_wait();
    extern __shared__ uint2 s_words[];
    const int64_t i = (int64_t)blockIdx.x * GATHER_THREADS + threadIdx.x;
    if (i >= n) return;
    uint32_t cnt = 0;
    gather<false>(x[i], (uint32_t)i, c.h2, x, soa, cell_range, g, s_words + threadIdx.x, [&](uint32_t, float4, int) { cnt++; });
    count[i] = cnt;
}

// Lazy module loading (the CUDA default) loads a kernel at its first launch, and that load can wait for
// running kernels to finish: if the running kernel is a neighbour rank's flag wait (several ranks in one
// process), the two deadlock until the time-out. pbf_create therefore loads every kernel up front.
cudaError_t preload_solver() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, false, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, false, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, true, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, true, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, true, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, true, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, false, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, false, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, true, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, true, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_kernel<0>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_kernel<1>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_kernel<2>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_kernel<3>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<0, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<1, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<2, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<3, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<0, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<1, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<2, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<3, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, update_velocity_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, xsph_kernel<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, xsph_kernel<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, neighbor_count_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda2_kernel<false, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda2_kernel<true, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda2_kernel<false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda2_kernel<true, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, xsph2_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, false, false, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, false, false, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, true, false, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, true, false, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<0, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<1, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<2, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<3, false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, xsph_kernel<false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, pack_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, pack_ghosts_kernel);
    if (e == cudaSuccess) e = preload_solver_team();
    return e;
}

static inline unsigned nblocks(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// Small scenes take the four-lanes-per-particle kernels of solver_team.cu (latency bound, not throughput bound).
// The handle's PBF_OPT_TEAM forces the choice (tests run the golden scenes through both, tuning experiments).
static bool use_team(const SweepMode& mode, int64_t n) {
    if (mode.morton) return false;   // (the Morton A/B exists for the thread kernels only)
    return mode.team < 0 ? n < TEAM_MAX_PARTICLES : mode.team == 1;
}

// the cull's coordinate arrays must mirror `x`: nothing to do if the producer of `x` wrote them along
static cudaError_t launch_pack(const float4* x, CullScratch& cs, int64_t n_slots, cudaStream_t st, int64_t* launches) {
    if (cs.holds == x) return cudaSuccess;
    PBF_LAUNCH((pack_kernel), nblocks(n_slots, 256), 256, 0, st, x, cs.xs[cs.cur], cs.ys[cs.cur], cs.zs[cs.cur], n_slots);
    if (launches) (*launches)++;
    cs.holds = x;
    return cudaGetLastError();
}
cudaError_t launch_pack_ghosts(const float4* x, CullScratch& cs, int64_t n_slots, int64_t own_first, int64_t own_count,
                               const HaloSync& hs, cudaStream_t st, int64_t* launches) {
    // the pass that produced x wrote the owned slots' coordinates into the OTHER set: switch to it, add the ghosts
    cs.cur ^= 1;
    cs.holds = x;
    // (launched even without ghost slots: its wait is also what lets this rank's NEXT pass overwrite the
    //  neighbours' ghost slots — they have finished reading them)
    const int64_t ghosts = n_slots - own_count > 0 ? n_slots - own_count : 1;
    PBF_LAUNCH((pack_ghosts_kernel), nblocks(ghosts, 256), 256, 0, st, x, cs.xs[cs.cur], cs.ys[cs.cur], cs.zs[cs.cur], own_first, own_count,
               n_slots, hs);
    if (launches) (*launches)++;
    return cudaGetLastError();
}
static inline CullSoA soa_of(const CullScratch& cs) { return CullSoA{cs.xs[cs.cur], cs.ys[cs.cur], cs.zs[cs.cur]}; }
static inline CullOut out_of(const CullScratch& cs) { return CullOut{cs.xs[cs.cur ^ 1], cs.ys[cs.cur ^ 1], cs.zs[cs.cur ^ 1]}; }

bool sweeps_use_team(const SweepMode& mode, int64_t n) { return use_team(mode, n); }
// whether launch_delta_p would take the replay THREAD kernel (the one that can run in slices)
bool delta_p_sliceable(const PairList& pl, const SweepMode& mode, int64_t n) { return pl.js != nullptr && !use_team(mode, n); }

size_t pair_list_bytes(int64_t max_particles, size_t* js_bytes, size_t* cnt_bytes) {
    const size_t blocks = (size_t)((max_particles + GATHER_THREADS - 1) / GATHER_THREADS);
    *js_bytes = blocks * PAIR_CAP * GATHER_THREADS * sizeof(uint2);
    *cnt_bytes = blocks * GATHER_THREADS * sizeof(uint32_t);
    return *js_bytes + *cnt_bytes;
}

cudaError_t launch_lambda(const float4* x, CullScratch& cs, int64_t n_slots, float4* xl, float* rho,
                          const uint2* cell_range, int64_t first, int64_t n, const PairList& pl, const HaloPush& hp, const HaloSync& hs_in,
                          const GridConsts& g, const SolverConsts& c, const SweepMode& mode, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    cudaError_t pe = launch_pack(x, cs, n_slots, st, launches);
    if (pe != cudaSuccess) return pe;
    const CullSoA soa = soa_of(cs);
    if (use_team(mode, n)) {
        launch_lambda_team(x, soa, xl, rho, cell_range, first, n, pl.js, pl.cnt, hp, hs_in, g, c, st);
        if (launches) (*launches)++;
        return cudaGetLastError();
    }
    const unsigned nb = nblocks(n, GATHER_THREADS);
    HaloSync hs = hs_in;
    halo_sync_blocks(hs, n, GATHER_THREADS);
#define PBF_LAMBDA_LAUNCH(SAVE, FAST)                                                                                          \
    do {                                                                                                                       \
        if (mode.morton)                                                                                                       \
            PBF_LAUNCH((lambda_kernel<SAVE, FAST, false, false, true>), nb, GATHER_THREADS, LIST_SMEM, st, x, soa, xl, rho, cell_range, first, n, \
                       pl.js, pl.cnt, hp, hs, g, c);                                                                         \
        else if (mode.paired)                                                                                                  \
            PBF_LAUNCH((lambda2_kernel<SAVE, FAST>), nb, PAIR_THREADS, LIST_SMEM, st, x, soa, xl, rho, cell_range, first, n,   \
                       pl.js, pl.cnt, hp, hs, g, c);                                                                         \
        else if (mode.staged && !mode.moved)                                                                                        \
            PBF_LAUNCH((lambda_kernel<SAVE, FAST, false, true>), nb, GATHER_THREADS, LIST_SMEM, st, x, soa, xl, rho, cell_range, first, n, \
                       pl.js, pl.cnt, hp, hs, g, c);                                                                         \
        else if (mode.rebin && mode.moved)                                                                                     \
            PBF_LAUNCH((lambda_kernel<SAVE, FAST, true>), nb, GATHER_THREADS, LIST_SMEM, st, x, soa, xl, rho, cell_range, first, n,       \
                                                                                 pl.js, pl.cnt, hp, hs, g, c);               \
        else                                                                                                                   \
            PBF_LAUNCH((lambda_kernel<SAVE, FAST, false>), nb, GATHER_THREADS, LIST_SMEM, st, x, soa, xl, rho, cell_range, first, n,      \
                                                                                  pl.js, pl.cnt, hp, hs, g, c);              \
    } while (0)
    if (!pl.js && !c.fast_spiky) PBF_LAMBDA_LAUNCH(false, false);
    else if (!pl.js) PBF_LAMBDA_LAUNCH(false, true);
    else if (!c.fast_spiky) PBF_LAMBDA_LAUNCH(true, false);
    else PBF_LAMBDA_LAUNCH(true, true);
#undef PBF_LAMBDA_LAUNCH
    if (launches) (*launches)++;
    return cudaGetLastError();
}

// (`cs` holds the positions the lambda pass of this iteration packed: the same ones xl carries)
cudaError_t launch_delta_p(const float4* xl, CullScratch& cs, int64_t n_slots, float4* x_out, const uint2* cell_range,
                           int64_t first, int64_t n, const PairList& pl, const HaloPush& hp, const HaloSync& hs_in, const VelTail& vt,
                           const GridConsts& g, const SolverConsts& c, const SweepMode& mode, cudaStream_t st, int64_t* launches,
                           uint32_t block0, uint32_t nblk, bool last_slice) {
    if (n <= 0) return cudaSuccess;
    const CullSoA soa = soa_of(cs);   // this iteration's coordinates: what the overflow kernel culls on
    const CullOut co = out_of(cs);    // the other set receives the coordinates of x_out
    // 3: the verified special-case-free powf(w, 4) (n_corr == 4, the default); 2: the library's powf with the
    // exponent folded (same bits; when 3 did not verify or is switched off); 1: powf, any exponent; 0: (w*w)^2
    const int pow_mode = c.n_corr == 4.0f ? (c.exact_pow ? (c.trim_pow ? 3 : 2) : 0) : 1;
    const unsigned nb = nblk ? nblk : nblocks(n, GATHER_THREADS);   // (a slice: nblk blocks from block0; replay thread kernel only)
    HaloSync hs = hs_in;
    halo_sync_blocks(hs, n, GATHER_THREADS);
#define PBF_DP_LAUNCH(POW)                                                                                                    \
    do {                                                                                                                      \
        if (pl.js) {                                                                                                          \
            if (use_team(mode, n)) launch_delta_p_replay_team(xl, x_out, co, first, n, pl.js, pl.cnt, cell_range, hp, hs_in, vt, g, c, POW, st); \
            else PBF_LAUNCH((delta_p_replay_kernel<POW>), nb, GATHER_THREADS, 0, st, xl, x_out, co, first, n, pl.js, pl.cnt,  \
                            cell_range, hp, hs, vt, g, c, block0);                                                           \
        } else if (mode.morton) {                                                                                             \
            PBF_LAUNCH((delta_p_kernel<POW, false, true>), nb, GATHER_THREADS, LIST_SMEM, st, xl, soa, x_out, co, cell_range, first, n, hp, hs, vt, g, c); \
        } else if (mode.rebin && mode.moved) {                                                                                \
            PBF_LAUNCH((delta_p_kernel<POW, true>), nb, GATHER_THREADS, LIST_SMEM, st, xl, soa, x_out, co, cell_range, first, n, hp, hs, vt, g, c); \
        } else {                                                                                                              \
            PBF_LAUNCH((delta_p_kernel<POW, false>), nb, GATHER_THREADS, LIST_SMEM, st, xl, soa, x_out, co, cell_range, first, n, hp, hs, vt, g, c); \
        }                                                                                                                     \
    } while (0)
    if (pow_mode == 3) PBF_DP_LAUNCH(3);
    else if (pow_mode == 2) PBF_DP_LAUNCH(2);
    else if (pow_mode == 1) PBF_DP_LAUNCH(1);
    else PBF_DP_LAUNCH(0);
#undef PBF_DP_LAUNCH
    if (launches) (*launches)++;
    if (!last_slice) return cudaGetLastError();
    // the other set now mirrors x_out — if the pass covered every stored slot (single GPU); in slab mode the ghost
    // slots of x_out are the neighbours' to fill, and the next sweep packs
    if (first == 0 && n == n_slots) {
        cs.cur ^= 1;
        cs.holds = x_out;
    } else {
        cs.holds = nullptr;   // (slab mode: launch_pack_ghosts, or a full pack, completes the set before the next sweep)
    }
    return cudaGetLastError();
}

cudaError_t launch_update_velocity(const float4* x, const float* rho, float* pos_out, float* npos_io,
                                   float* vel_out, float4* v4, int64_t first, int64_t n, const HaloPush& hp,
                                   const HaloSync& hs_in, const StatePush& sp, const SolverConsts& c, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    HaloSync hs = hs_in;
    halo_sync_blocks(hs, n, 256);
    PBF_LAUNCH((update_velocity_kernel), nblocks(n, 256), 256, 0, st, x, rho, pos_out, npos_io, vel_out, v4, first, n, hp, hs, sp, c);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

// n_slots > 0: refresh the cull's coordinate arrays first; 0: they are current (a further chunk of the same sweep)
cudaError_t launch_xsph(const float4* x, CullScratch& cs, int64_t n_slots, const float4* v4,
                        const uint2* cell_range, float* nvel_out, const uint32_t* iid_sorted, uint32_t* iid_out,
                        int64_t first, int64_t n, const HaloSync& hs_in, const StatePush& sp, const GridConsts& g, const SolverConsts& c,
                        const SweepMode& mode, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    if (n_slots > 0) {
        cudaError_t pe = launch_pack(x, cs, n_slots, st, launches);
        if (pe != cudaSuccess) return pe;
    }
    if (use_team(mode, n)) {
        launch_xsph_team(x, soa_of(cs), v4, cell_range, nvel_out, iid_sorted, iid_out, first, n, hs_in, sp, g, c, st);
        if (launches) (*launches)++;
        return cudaGetLastError();
    }
    HaloSync hs = hs_in;
    halo_sync_blocks(hs, n, GATHER_THREADS);
    if (mode.morton)
        PBF_LAUNCH((xsph_kernel<false, true>), nblocks(n, GATHER_THREADS), GATHER_THREADS, LIST_SMEM, st, x, soa_of(cs), v4, cell_range, nvel_out, iid_sorted, iid_out, first, n, hs, sp, g, c);
    else if (mode.paired)
        PBF_LAUNCH((xsph2_kernel), nblocks(n, GATHER_THREADS), PAIR_THREADS, LIST_SMEM, st, x, soa_of(cs), v4, cell_range, nvel_out, iid_sorted, iid_out, first, n, hs, sp, g, c);
    else if (mode.rebin && mode.moved)
        PBF_LAUNCH((xsph_kernel<true>), nblocks(n, GATHER_THREADS), GATHER_THREADS, LIST_SMEM, st, x, soa_of(cs), v4, cell_range, nvel_out, iid_sorted, iid_out, first, n, hs, sp, g, c);
    else
        PBF_LAUNCH((xsph_kernel<false>), nblocks(n, GATHER_THREADS), GATHER_THREADS, LIST_SMEM, st, x, soa_of(cs), v4, cell_range, nvel_out, iid_sorted, iid_out, first, n, hs, sp, g, c);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_neighbor_count(const float4* x, CullScratch& cs, const uint2* cell_range, uint32_t* count,
                                  int64_t n, const GridConsts& g, const SolverConsts& c, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    cudaError_t pe = launch_pack(x, cs, n, st, nullptr);
    if (pe != cudaSuccess) return pe;
    PBF_LAUNCH((neighbor_count_kernel), nblocks(n, GATHER_THREADS), GATHER_THREADS, LIST_SMEM, st, x, soa_of(cs), cell_range, count, n, g, c);
    return cudaGetLastError();
}

}  // namespace pbf
