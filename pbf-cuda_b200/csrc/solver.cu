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
// data path: one 16-byte load per candidate from the sorted float4 array instead of three
// scalar loads from AoS float3 (+1 for lambda), the three z-cells of a column visited as one
// contiguous slot run (9 runs instead of 27 cells), and the expensive part (sqrt, 4 IEEE
// divisions, pow) only for pairs that pass an r2 cull — candidates outside h contribute exact
// zeros in the reference, so skipping them does not change a single bit.
#include "pbf_math.cuh"

namespace pbf {

#ifndef PBF_GATHER_MINBLOCKS
#define PBF_GATHER_MINBLOCKS 8
#endif
#ifndef PBF_WORD_CAP
#define PBF_WORD_CAP 15
#endif
#ifndef PBF_FLUSH_PER_SLAB
#define PBF_FLUSH_PER_SLAB 0
#endif
#ifndef PBF_PAIR_CAP
#define PBF_PAIR_CAP 96
#endif
constexpr int GATHER_THREADS = 128;
constexpr int WORD_CAP = PBF_WORD_CAP;  // hit words (32 candidates each) buffered per thread before a flush
constexpr int PAIR_CAP = PBF_PAIR_CAP;  // neighbours per particle the lambda pass can hand to the delta-p pass
constexpr size_t LIST_SMEM = (size_t)WORD_CAP * GATHER_THREADS * sizeof(uint2);  // 8 KB per CTA

// Two-phase gather of one particle (one thread), the core of all three neighbour sweeps.
//
// Phase 1 (cull) walks the candidates in the reference's visiting order — dx, dy, dz nested,
// ascending slot inside a cell; the three dz cells of a column are consecutive keys, hence ONE
// contiguous slot run per (dx, dy), and the runs are visited in ascending slot order — with one
// 16-byte load, 7 flops and one funnel shift per candidate and nothing else: the sign bit of
// RN(r2 - limit) (set exactly when r2 < limit; IEEE subtraction with denormals never rounds a
// non-zero difference to zero) is shifted into a hit word, 32 candidates per word, first
// candidate in the top bit. Non-empty words go to a per-thread list in shared memory as
// (first slot, hits) (entry k of thread t at [k][t]: conflict-free). Runs are read in groups of
// four, up to three slots past their end (the arrays are padded); those bits are masked off.
// Phase 2 (`heavy`) then runs the expensive exact arithmetic over the set bits only — count
// leading zeros, clear, next word when empty — every lane busy, still in visiting order.
// Without the split a warp executes the heavy path for nearly every candidate, because some
// lane is almost always in range (~15 % of the candidates are), at ~15 % lane utilisation.
// The list is drained after each dx slab (3 runs, ~11 neighbours) and whenever it is full, so
// any neighbour count stays correct and ordered at 8 KB of shared memory per CTA, which leaves
// the rest of the 256 KB for L1.
__device__ __forceinline__ uint32_t push_hit(uint32_t hits, const float4 p, const float4 q, const float limit) {
    const float r2 = sumsq(__fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z));
    return __funnelshift_l(__float_as_uint(__fsub_rn(r2, limit)), hits, 1);
}

// Four consecutive candidates. sm_100 has 256-bit global loads (LDG.E.256): two of them instead of four
// LDG.128 halve the L1 tag look-ups and data wavefronts of the cull, which is what bounds the later Jacobi
// iterations (lanes whose home cell drifted apart read different lines: 8.5 tags per warp load, ncu).
// `xp` is 32-byte aligned: words start at even slots.
#ifndef PBF_CULL_LD256
#define PBF_CULL_LD256 1
#endif
#ifndef PBF_CULL_PIPE
#define PBF_CULL_PIPE 0
#endif
#ifndef PBF_HEAVY_PREFETCH
#define PBF_HEAVY_PREFETCH 0
#endif
__device__ __forceinline__ void load4(const float4* __restrict__ xp, float4& q0, float4& q1, float4& q2, float4& q3) {
#if PBF_CULL_LD256
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w), "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w) : "l"(xp));
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w), "=f"(q3.x), "=f"(q3.y), "=f"(q3.z), "=f"(q3.w) : "l"(xp + 2));
#else
    q0 = __ldg(xp); q1 = __ldg(xp + 1); q2 = __ldg(xp + 2); q3 = __ldg(xp + 3);
#endif
}

template <bool SKIP_SELF, typename Heavy>
__device__ __forceinline__ void gather(const float4 p, const uint32_t self, const float limit,
                                       const float4* __restrict__ x, const uint2* __restrict__ cell_range,
                                       const GridConsts& g, uint2* __restrict__ my_words, Heavy&& heavy) {
    const int3 cc = cell_of(p.x, p.y, p.z, g);
    const int zlo = max(cc.z - 1, 0), zhi = min(cc.z + 1, g.dim[2] - 1);
    uint2* const words_end = my_words + WORD_CAP * GATHER_THREADS;
    uint2* tail = my_words;  // next free entry of this thread's list
    int k_total = 0;         // in-range neighbours handed to `heavy` so far (its third argument)
    auto flush = [&]() {
        const uint2* e = my_words;
        uint32_t first = 0, hits = 0;
        auto next = [&](uint32_t& j) -> bool {  // next set bit of the list, in order
            if (hits == 0) {
                if (e == tail) return false;
                const uint2 w = *e;
                e += GATHER_THREADS;
                first = w.x;
                hits = w.y;
            }
            const int lead = __clz((int)hits);
            hits &= ~(0x80000000u >> lead);
            j = first + (uint32_t)lead;
            return true;
        };
#if PBF_HEAVY_PREFETCH
        // the neighbour after the current one is loaded before the current one's arithmetic
        uint32_t j = 0, jn = 0;
        float4 q = p, qn = p;
        bool have = next(j);
        if (have) q = __ldg(&x[j]);
        while (have) {
            const bool have_n = next(jn);
            if (have_n) qn = __ldg(&x[jn]);
            if (!(SKIP_SELF && j == self)) {
                heavy(j, q, k_total);
                k_total++;
            }
            j = jn;
            q = qn;
            have = have_n;
        }
#else
        uint32_t j;
        while (next(j)) {
            if (SKIP_SELF && j == self) continue;
            heavy(j, __ldg(&x[j]), k_total);
            k_total++;
        }
#endif
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
            uint32_t start = 0, end = 0;
            bool any = false;
            for (int z = zlo; z <= zhi; z++) {
                const uint2 r = __ldg(&cell_range[cbase + z]);
                if (r.y > r.x) {
                    if (!any) { start = r.x; any = true; }
                    end = r.y;
                }
            }
#pragma unroll 1
            for (uint32_t b = start & ~1u; b < end; b += 32) {   // words start at even slots (load4)
                const uint32_t cnt = min(end - b, 32u);   // slots of this word up to the end of the run
                const uint32_t groups = (cnt + 3) >> 2;
                const float4* xp = x + b;
                uint32_t hits = 0;
#if PBF_CULL_PIPE
                // software pipeline: the next group's loads are issued before this group's arithmetic
                // (two register sets A / B in ping-pong, so no register is moved)
                float4 a0, a1, a2, a3, b0, b1, b2, b3;
                load4(xp, a0, a1, a2, a3);
                uint32_t left = groups;  // groups not yet pushed, A holds the first of them
#pragma unroll 1
                for (;;) {
                    if (left > 1) load4(xp + 4, b0, b1, b2, b3);
                    hits = push_hit(hits, p, a0, limit);
                    hits = push_hit(hits, p, a1, limit);
                    hits = push_hit(hits, p, a2, limit);
                    hits = push_hit(hits, p, a3, limit);
                    if (left == 1) break;
                    if (left > 2) load4(xp + 8, a0, a1, a2, a3);
                    hits = push_hit(hits, p, b0, limit);
                    hits = push_hit(hits, p, b1, limit);
                    hits = push_hit(hits, p, b2, limit);
                    hits = push_hit(hits, p, b3, limit);
                    if (left == 2) break;
                    left -= 2;
                    xp += 8;
                }
#else
#pragma unroll 1
                for (uint32_t gi = 0; gi < groups; gi++, xp += 4) {  // four candidates in flight per lane
                    float4 q0, q1, q2, q3;
                    load4(xp, q0, q1, q2, q3);
                    hits = push_hit(hits, p, q0, limit);
                    hits = push_hit(hits, p, q1, limit);
                    hits = push_hit(hits, p, q2, limit);
                    hits = push_hit(hits, p, q3, limit);
                }
#endif
                // first slot to the top bit; drop the slot before the run (odd start) and what was read past its end
                hits = (hits << (32 - 4 * groups)) & (0xffffffffu << (32 - cnt)) & (0xffffffffu >> (b < start ? 1 : 0));
                *tail = make_uint2(b, hits);
                tail += hits ? GATHER_THREADS : 0;
                if (tail == words_end) flush();
            }
        }
        if (PBF_FLUSH_PER_SLAB) flush();
    }
    if (!PBF_FLUSH_PER_SLAB) flush();
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
constexpr uint32_t PAIR_OVERFLOW = 1u << 31;

template <bool EXACT_POW, bool SAVE_PAIRS>
__global__ void __launch_bounds__(GATHER_THREADS, PBF_GATHER_MINBLOCKS)
lambda_kernel(const float4* __restrict__ x, float4* __restrict__ xl, float* __restrict__ rho_out,
              const uint2* __restrict__ cell_range, int64_t first, int64_t n,
              uint2* __restrict__ pair_js, uint32_t* __restrict__ pair_cnt,
              const __grid_constant__ HaloPush hp, const __grid_constant__ GridConsts g,
              const __grid_constant__ SolverConsts c) {
    extern __shared__ uint2 s_words[];
    const int64_t t = (int64_t)blockIdx.x * GATHER_THREADS + threadIdx.x;
    if (t >= n) return;
    const int64_t i = first + t;
    const float4 p = x[i];
    float rho = 0.f, gradj_l2 = 0.f, gix = 0.f, giy = 0.f, giz = 0.f;
    // the particle itself: r2 = 0 adds poly6(0) to rho at its place in the visiting order and
    // nothing else (spiky is 0 below KERNAL_EPS, and gradj skips j == i). Taking it out of the
    // general path keeps three 0/rho0 divisions off IEEE division's slow path in every warp.
    const float w_self = poly6_in(0.f, c);
    const size_t pair0 = (size_t)blockIdx.x * PAIR_CAP * GATHER_THREADS + threadIdx.x;
    int n_pairs = 0;
    gather<false>(p, (uint32_t)i, c.h2_cull, x, cell_range, g, s_words + threadIdx.x, [&](uint32_t j, float4 q, int) {
        if (j == (uint32_t)i) {
            rho = __fadd_rn(rho, w_self);
        } else {
            const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
            const float r2 = sumsq(dx, dy, dz);
            rho = __fadd_rn(rho, poly6(r2, c));
            const float s = spiky_scale(r2, c);
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
                if (n_pairs < PAIR_CAP) pair_js[pair0 + (size_t)n_pairs * GATHER_THREADS] = make_uint2(j, __float_as_uint(s));
                n_pairs++;
            }
        }
    });
    if (c.k_boundary != 0.f) rho = __fmaf_rn(c.k_boundary, boundary_density(p.x, p.y, p.z, g), rho);
    const float grad_l2 = __fmaf_rn(giz, giz, __fmaf_rn(giy, giy, __fmaf_rn(gix, gix, gradj_l2)));
    const float lambda = __fdiv_rn(-__fadd_rn(__fdiv_rn(rho, c.pho0), -1.f), __fadd_rn(grad_l2, c.lambda_eps));
    const float4 out = make_float4(p.x, p.y, p.z, lambda);
    xl[i] = out;
    halo_push(hp, t, out);
    rho_out[i] = rho;
    if (SAVE_PAIRS) pair_cnt[t] = n_pairs <= PAIR_CAP ? (uint32_t)n_pairs : PAIR_OVERFLOW;
}

// shared tail of the delta-p pass: divide, clamp to MAX_DP, add, clamp to the box (f64 like the reference)
__device__ __forceinline__ float4 delta_p_finish(const float4 p, float ax, float ay, float az, const SolverConsts& c) {
    const float max_dp = (float)0.1;  // MAX_DP through clamp3f's float parameters (helper.h:9,26)
    div3_pho0(ax, ay, az, c);
    const float vx = fmaxf(fminf(ax, max_dp), -max_dp);
    const float vy = fmaxf(fminf(ay, max_dp), -max_dp);
    const float vz = fmaxf(fminf(az, max_dp), -max_dp);
    const float qx = (float)fmax(fmin((double)__fadd_rn(p.x, vx), c.lim_hi[0]), c.lim_lo[0]);
    const float qy = (float)fmax(fmin((double)__fadd_rn(p.y, vy), c.lim_hi[1]), c.lim_lo[1]);
    const float qz = (float)fmax(fmin((double)__fadd_rn(p.z, vz), c.lim_hi[2]), c.lim_lo[2]);
    return make_float4(qx, qy, qz, 0.f);
}

// The delta-p pass comes as two kernels. The REPLAY kernel walks the neighbour list the lambda pass saved:
// no cell table, no shared-memory list, few registers — it is latency / HBM bound, so it is compiled for
// 16 CTAs per SM instead of 8 (32 registers, no spills; measured 0.28 -> 0.20 ms in the compressed state). Particles whose list overflowed
// (more than PAIR_CAP neighbours) are left to the GATHER kernel, which re-runs the full two-phase gather
// for them only (ONLY_OVERFLOW) — or for everybody when there is no list at all.
#ifndef PBF_REPLAY_MINBLOCKS
#define PBF_REPLAY_MINBLOCKS 16
#endif
template <bool EXACT_POW>
__global__ void __launch_bounds__(GATHER_THREADS, PBF_REPLAY_MINBLOCKS)
delta_p_replay_kernel(const float4* __restrict__ xl, float4* __restrict__ x_out, int64_t first, int64_t n,
                      const uint2* __restrict__ pair_js,
                      const uint32_t* __restrict__ pair_cnt, const __grid_constant__ HaloPush hp,
                      const __grid_constant__ SolverConsts c) {
    const int64_t t = (int64_t)blockIdx.x * GATHER_THREADS + threadIdx.x;
    if (t >= n) return;
    const uint32_t cnt = pair_cnt[t];
    if (cnt & PAIR_OVERFLOW) return;   // the gather kernel's particle
    const int64_t i = first + t;
    const float4 p = xl[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    const size_t pair0 = (size_t)blockIdx.x * PAIR_CAP * GATHER_THREADS + threadIdx.x;
#pragma unroll 4
    for (uint32_t k = 0; k < cnt; k++) {
        const size_t e = pair0 + (size_t)k * GATHER_THREADS;
        const uint2 js = __ldg(&pair_js[e]);
        const float4 q = __ldg(&xl[js.x]);
        const float sj = __uint_as_float(js.y);  // spiky scale of the pair, saved by the lambda pass
        const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
        const float w = poly6(sumsq(dx, dy, dz), c);
        float pw;
        if (EXACT_POW) {
            pw = powf(w, c.n_corr);
        } else {  // n_corr == 4
            const float w2 = __fmul_rn(w, w);
            pw = __fmul_rn(w2, w2);
        }
        const float sc = __fmaf_rn(c.coef_corr, pw, __fadd_rn(p.w, q.w));
        ax = __fmaf_rn(sc, __fmul_rn(dx, sj), ax);
        ay = __fmaf_rn(sc, __fmul_rn(dy, sj), ay);
        az = __fmaf_rn(sc, __fmul_rn(dz, sj), az);
    }
    const float4 out = delta_p_finish(p, ax, ay, az, c);
    x_out[i] = out;
    halo_push(hp, t, out);
}

template <bool EXACT_POW, bool ONLY_OVERFLOW>
__global__ void __launch_bounds__(GATHER_THREADS, PBF_GATHER_MINBLOCKS)
delta_p_kernel(const float4* __restrict__ xl, float4* __restrict__ x_out,
               const uint2* __restrict__ cell_range, int64_t first, int64_t n,
               const uint32_t* __restrict__ pair_cnt, const __grid_constant__ HaloPush hp,
               const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    extern __shared__ uint2 s_words[];
    const int64_t t = (int64_t)blockIdx.x * GATHER_THREADS + threadIdx.x;
    if (t >= n) return;
    if (ONLY_OVERFLOW && !(pair_cnt[t] & PAIR_OVERFLOW)) return;
    const int64_t i = first + t;
    const float4 p = xl[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    gather<true>(p, (uint32_t)i, c.h2_cull, xl, cell_range, g, s_words + threadIdx.x, [&](uint32_t, float4 q, int) {
        const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
        const float r2 = sumsq(dx, dy, dz);
        const float w = poly6(r2, c);
        float pw;
        if (EXACT_POW) {
            pw = powf(w, c.n_corr);
        } else {  // n_corr == 4
            const float w2 = __fmul_rn(w, w);
            pw = __fmul_rn(w2, w2);
        }
        const float sc = __fmaf_rn(c.coef_corr, pw, __fadd_rn(p.w, q.w));
        const float s = spiky_scale(r2, c);
        ax = __fmaf_rn(sc, __fmul_rn(dx, s), ax);
        ay = __fmaf_rn(sc, __fmul_rn(dy, s), ay);
        az = __fmaf_rn(sc, __fmul_rn(dz, s), az);
    });
    const float4 out = delta_p_finish(p, ax, ay, az, c);
    x_out[i] = out;
    halo_push(hp, t, out);
}

// vel = (npos - pos) * inv_dt, plus everything the caller-facing buffers need from this point:
// pos <- step-input position (parked in npos by the reorder pass), npos <- final iterate.
__global__ void __launch_bounds__(256)
update_velocity_kernel(const float4* __restrict__ x, const float* __restrict__ rho,
                       float* __restrict__ pos_out, float* __restrict__ npos_io,
                       float* __restrict__ vel_out, float4* __restrict__ v4, int64_t first, int64_t n,
                       const __grid_constant__ HaloPush hp, const __grid_constant__ SolverConsts c) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
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
}

__global__ void __launch_bounds__(GATHER_THREADS, PBF_GATHER_MINBLOCKS)
xsph_kernel(const float4* __restrict__ x, const float4* __restrict__ v4,
            const uint2* __restrict__ cell_range, float* __restrict__ nvel_out,
            const uint32_t* __restrict__ iid_sorted, uint32_t* __restrict__ iid_out, int64_t first, int64_t n,
            const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    extern __shared__ uint2 s_words[];
    const int64_t t = (int64_t)blockIdx.x * GATHER_THREADS + threadIdx.x;
    if (t >= n) return;
    const int64_t i = first + t;
    const float4 p = x[i];
    const float4 vi = v4[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    gather<false>(p, (uint32_t)i, c.h2, x, cell_range, g, s_words + threadIdx.x, [&](uint32_t j, float4 q, int) {
        const float r2 = sumsq(__fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z));
        const float4 vj = __ldg(&v4[j]);
        const float w = poly6_in(r2, c);
        const float den = __fadd_rn(vi.w, vj.w);
        const float tx = __fsub_rn(vj.x, vi.x), ty = __fsub_rn(vj.y, vi.y), tz = __fsub_rn(vj.z, vi.z);
        ax = __fadd_rn(ax, __fdiv_rn(__fmul_rn(__fadd_rn(tx, tx), w), den));
        ay = __fadd_rn(ay, __fdiv_rn(__fmul_rn(__fadd_rn(ty, ty), w), den));
        az = __fadd_rn(az, __fdiv_rn(__fmul_rn(__fadd_rn(tz, tz), w), den));
    });
    store_f3(nvel_out, t, __fmaf_rn(c.c_xsph, ax, vi.x), __fmaf_rn(c.c_xsph, ay, vi.y), __fmaf_rn(c.c_xsph, az, vi.z));
    iid_out[t] = iid_sorted[i];
}

__global__ void __launch_bounds__(GATHER_THREADS)
neighbor_count_kernel(const float4* __restrict__ x, const uint2* __restrict__ cell_range,
                      uint32_t* __restrict__ count, int64_t n, const __grid_constant__ GridConsts g,
                      const __grid_constant__ SolverConsts c) {
    extern __shared__ uint2 s_words[];
    const int64_t i = (int64_t)blockIdx.x * GATHER_THREADS + threadIdx.x;
    if (i >= n) return;
    uint32_t cnt = 0;
    gather<false>(x[i], (uint32_t)i, c.h2, x, cell_range, g, s_words + threadIdx.x, [&](uint32_t, float4, int) { cnt++; });
    count[i] = cnt;
}

// Lazy module loading (the CUDA default) loads a kernel at its first launch, and that load can wait for
// running kernels to finish: if the running kernel is a neighbour rank's flag wait (several ranks in one
// process), the two deadlock until the time-out. pbf_create therefore loads every kernel up front.
cudaError_t preload_solver() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<true, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_kernel<false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_kernel<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_kernel<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<true, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<true, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_kernel<false, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, update_velocity_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, xsph_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, neighbor_count_kernel);
    return e;
}

static inline unsigned nblocks(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

size_t pair_list_bytes(int64_t max_particles, size_t* js_bytes, size_t* cnt_bytes) {
    const size_t blocks = (size_t)((max_particles + GATHER_THREADS - 1) / GATHER_THREADS);
    *js_bytes = blocks * PAIR_CAP * GATHER_THREADS * sizeof(uint2);
    *cnt_bytes = blocks * GATHER_THREADS * sizeof(uint32_t);
    return *js_bytes + *cnt_bytes;
}

cudaError_t launch_lambda(const float4* x, float4* xl, float* rho, const uint2* cell_range, int64_t first,
                          int64_t n, const PairList& pl, const HaloPush& hp, const GridConsts& g,
                          const SolverConsts& c, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    const bool exact = c.exact_pow || c.n_corr != 4.0f;
    const unsigned nb = nblocks(n, GATHER_THREADS);
    if (!pl.js)
        lambda_kernel<false, false><<<nb, GATHER_THREADS, LIST_SMEM, st>>>(x, xl, rho, cell_range, first, n, nullptr, nullptr, hp, g, c);
    else if (exact)
        lambda_kernel<true, true><<<nb, GATHER_THREADS, LIST_SMEM, st>>>(x, xl, rho, cell_range, first, n, pl.js, pl.cnt, hp, g, c);
    else
        lambda_kernel<false, true><<<nb, GATHER_THREADS, LIST_SMEM, st>>>(x, xl, rho, cell_range, first, n, pl.js, pl.cnt, hp, g, c);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_delta_p(const float4* xl, float4* x_out, const uint2* cell_range, int64_t first, int64_t n,
                           const PairList& pl, const HaloPush& hp, const GridConsts& g, const SolverConsts& c,
                           cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    const bool exact = c.exact_pow || c.n_corr != 4.0f;
    const unsigned nb = nblocks(n, GATHER_THREADS);
    if (pl.js) {
        if (exact) {
            delta_p_replay_kernel<true><<<nb, GATHER_THREADS, 0, st>>>(xl, x_out, first, n, pl.js, pl.cnt, hp, c);
            delta_p_kernel<true, true><<<nb, GATHER_THREADS, LIST_SMEM, st>>>(xl, x_out, cell_range, first, n, pl.cnt, hp, g, c);
        } else {
            delta_p_replay_kernel<false><<<nb, GATHER_THREADS, 0, st>>>(xl, x_out, first, n, pl.js, pl.cnt, hp, c);
            delta_p_kernel<false, true><<<nb, GATHER_THREADS, LIST_SMEM, st>>>(xl, x_out, cell_range, first, n, pl.cnt, hp, g, c);
        }
        if (launches) (*launches)++;
    } else {
        if (exact) delta_p_kernel<true, false><<<nb, GATHER_THREADS, LIST_SMEM, st>>>(xl, x_out, cell_range, first, n, nullptr, hp, g, c);
        else delta_p_kernel<false, false><<<nb, GATHER_THREADS, LIST_SMEM, st>>>(xl, x_out, cell_range, first, n, nullptr, hp, g, c);
    }
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_update_velocity(const float4* x, const float* rho, float* pos_out, float* npos_io,
                                   float* vel_out, float4* v4, int64_t first, int64_t n, const HaloPush& hp,
                                   const SolverConsts& c, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    update_velocity_kernel<<<nblocks(n, 256), 256, 0, st>>>(x, rho, pos_out, npos_io, vel_out, v4, first, n, hp, c);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_xsph(const float4* x, const float4* v4, const uint2* cell_range, float* nvel_out,
                        const uint32_t* iid_sorted, uint32_t* iid_out, int64_t first, int64_t n,
                        const GridConsts& g, const SolverConsts& c, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return cudaSuccess;
    xsph_kernel<<<nblocks(n, GATHER_THREADS), GATHER_THREADS, LIST_SMEM, st>>>(x, v4, cell_range, nvel_out, iid_sorted, iid_out, first, n, g, c);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_neighbor_count(const float4* x, const uint2* cell_range, uint32_t* count, int64_t n,
                                  const GridConsts& g, const SolverConsts& c, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    neighbor_count_kernel<<<nblocks(n, GATHER_THREADS), GATHER_THREADS, LIST_SMEM, st>>>(x, cell_range, count, n, g, c);
    return cudaGetLastError();
}

}  // namespace pbf
