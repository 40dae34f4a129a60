// solver_team.cu — the neighbour sweeps for SMALL scenes: four lanes per particle.
//
// With one thread per particle (solver.cu) a 32 000-particle scene is 250 CTAs on 148 SMs, ~7 warps per SM, and
// every thread walks ~216 candidates and ~33 neighbours one after the other: the kernels are bound by the latency
// of that one serial chain (ncu: 26 % of issue slots, 9 % occupancy). Here a TEAM of four adjacent lanes shares a
// particle: the cull takes four groups of four slots per step, the exact pair arithmetic four neighbours per
// round — four times the warps, a quarter of the chain — and what has to stay ordered stays ordered:
//   * the hits of a 32-slot word are assembled from the lanes' nibbles (two butterfly shuffles) and expanded, in
//     slot order, into the team's neighbour list in shared memory;
//   * a round evaluates neighbours k, k+1, k+2, k+3 of that list on lanes 0..3 and then ALL four lanes add the
//     four results in list order (broadcast by shuffles), so every lane carries the sums the single thread of
//     solver.cu would carry: same operations, same order, same bits (tests: the golden scenes run through both).
// A slot that is not a neighbour (past the end of the list) enters the sums as an exact +0, which changes no
// bit: the accumulators start at +0 and x + (+0) == x for every x that is not -0, and a sum that started at +0
// never becomes -0.
// More instruction slots per pair than solver.cu (shuffles, the list expansion), so the launchers take these
// kernels only below TEAM_MAX_PARTICLES, where latency, not throughput, is the bound.
#include <cooperative_groups.h>

#include "launch.cuh"
#include "solver_common.cuh"

namespace cg = cooperative_groups;

namespace pbf {

namespace {

constexpr int TEAM = 4;                                  // lanes per particle
constexpr int TEAM_THREADS = 128;
constexpr int TEAM_PARTICLES = TEAM_THREADS / TEAM;      // 32 particles per CTA
constexpr int TEAM_LIST = 128;                           // neighbour slots a team buffers before it evaluates them
constexpr size_t TEAM_SMEM = (size_t)TEAM_PARTICLES * TEAM_LIST * sizeof(uint32_t);   // 16 KB
static_assert(GATHER_THREADS % TEAM_PARTICLES == 0, "a team CTA must not straddle two blocks of the pair list");

struct Team {
    uint32_t lane;   // 0..3 inside the team
    uint32_t mask;   // the team's four lanes of the warp
};
__device__ __forceinline__ Team team_of() {
    const uint32_t l = threadIdx.x & 31u;
    return Team{l & 3u, 0xfu << (l & ~3u)};
}
template <typename T>
__device__ __forceinline__ T team_bcast(const Team& tm, T v, int src) { return __shfl_sync(tm.mask, v, src, TEAM); }

// Cooperative two-phase gather of one particle by its team. `eval(j, q, valid)` computes one neighbour's
// contribution on the calling lane (valid = false: a padding lane of the last round, must yield exact zeros);
// `add(m)` is then executed by all four lanes for m = 0..3 in order and has to fetch lane m's contribution with
// team_bcast and accumulate it. `list` = the team's TEAM_LIST slots of shared memory.
template <typename Eval, typename Add>
__device__ __forceinline__ void team_gather(const Team& tm, const float4 p, const float limit, const float4* __restrict__ x,
                                            const CullSoA soa, const uint2* __restrict__ cell_range, const GridConsts& g,
                                            uint32_t* __restrict__ list, Eval&& eval, Add&& add) {
    const int3 cc = cell_of(p.x, p.y, p.z, g);
    const bool has_below = cc.z > 0, has_above = cc.z + 1 < g.dim[2];
    const f32x2 px = pack2(p.x, p.x), py = pack2(p.y, p.y), pz = pack2(p.z, p.z), lim = pack2(limit, limit);
    uint32_t n_list = 0;   // neighbours buffered (the same value on the four lanes)
    auto drain = [&]() {
        __syncwarp(tm.mask);   // the list is complete
        for (uint32_t k0 = 0; k0 < n_list; k0 += TEAM) {
            const uint32_t k = k0 + tm.lane;
            const bool valid = k < n_list;
            const uint32_t j = list[valid ? k : 0];
            eval(j, __ldg(&x[j]), valid);
#pragma unroll
            for (int m = 0; m < TEAM; m++) add(m);
        }
        __syncwarp(tm.mask);   // everybody has read it: it may be overwritten
        n_list = 0;
    };
#pragma unroll 1
    for (int dx = -1; dx <= 1; dx++) {
        const int cx = cc.x + dx;
        const int lx = cx - g.xoff;
        if (cx < 0 || cx >= g.dim[0]) continue;
        if (lx < 0 || lx >= g.nxl) {
            if (g.flags && tm.lane == 0) atomicOr(g.flags, (uint32_t)PBF_SLAB_FLAG_GHOST);   // see solver.cu gather()
            continue;
        }
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++) {
            const int cy = cc.y + dy;
            if (cy < 0 || cy >= g.dim[1]) continue;
            const int cbase = lx * g.dyz + cy * g.dim[2];
            // the run = from the first slot of the first non-empty cell of the column's (up to) three to the
            // end of the last non-empty one; empty and out-of-range cells read {0, 0}, so an empty column gives
            // start == end == 0. Straight-line on purpose: as a loop over z (1-3 trips) this was ~80 instructions.
            const uint2 zero = make_uint2(0u, 0u);
            const uint2 r0 = has_below ? __ldg(&cell_range[cbase + cc.z - 1]) : zero;
            const uint2 r1 = __ldg(&cell_range[cbase + cc.z]);
            const uint2 r2 = has_above ? __ldg(&cell_range[cbase + cc.z + 1]) : zero;
            const bool e0 = r0.y > r0.x, e1 = r1.y > r1.x, e2 = r2.y > r2.x;
            const uint32_t start = e0 ? r0.x : e1 ? r1.x : r2.x;
            const uint32_t end = e2 ? r2.y : e1 ? r1.y : r0.y;
#pragma unroll 1
            for (uint32_t b = start & ~3u; b < end; b += 32) {   // one 32-slot word per step of the team
                // lane l tests the groups of four slots at b + 4l and b + 16 + 4l (loads only below `end`: the
                // arrays are padded by 8 slots, not by 28)
                uint32_t part = 0;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t s0 = b + 16u * h + 4u * tm.lane;
                    if (s0 < end) {
                        const float4 X = __ldg(reinterpret_cast<const float4*>(soa.xs + s0));
                        const float4 Y = __ldg(reinterpret_cast<const float4*>(soa.ys + s0));
                        const float4 Z = __ldg(reinterpret_cast<const float4*>(soa.zs + s0));
                        uint32_t nib = push_hits2(0u, px, py, pz, lim, X.x, X.y, Y.x, Y.y, Z.x, Z.y);
                        nib = push_hits2(nib, px, py, pz, lim, X.z, X.w, Y.z, Y.w, Z.z, Z.w);
                        part |= (nib & 0xfu) << (28u - 16u * h - 4u * tm.lane);   // first slot of the word in the top bit
                    }
                }
                part |= __shfl_xor_sync(tm.mask, part, 1, TEAM);
                part |= __shfl_xor_sync(tm.mask, part, 2, TEAM);
                const uint32_t cnt = min(end - b, 32u);
                const uint32_t word = part & (0xffffffffu << (32 - cnt)) & (0xffffffffu >> (b < start ? start - b : 0));
                const uint32_t nw = __popc(word);
                if (n_list + nw > TEAM_LIST) drain();
                // expansion in slot order: lane l owns byte l of the word (bits 31-8l .. 24-8l)
                uint32_t mine = (word << (8u * tm.lane)) & 0xff000000u;
                uint32_t at = n_list + (tm.lane ? __popc(word >> (32u - 8u * tm.lane)) : 0u);
                const uint32_t slot0 = b + 8u * tm.lane;
                while (mine) {
                    const int lead = __clz((int)mine);
                    mine &= ~(0x80000000u >> lead);
                    list[at++] = slot0 + (uint32_t)lead;
                }
                n_list += nw;
            }
        }
    }
    drain();
}

}  // namespace

// ---- lambda pass -----------------------------------------------------------------------------------------

// (the bodies of the three team kernels are device functions of a logical block `lb`, so that the persistent cooperative
//  kernel at the end of this file can run the same code for all passes of a step)
template <bool SAVE_PAIRS, bool FAST_SPIKY>
__device__ __forceinline__ void lambda_team_body(const float4* __restrict__ x, const CullSoA soa, float4* __restrict__ xl, float* __restrict__ rho_out,
                   const uint2* __restrict__ cell_range, int64_t first, int64_t n,
                   uint2* __restrict__ pair_js, uint32_t* __restrict__ pair_cnt,
                   const HaloPush& hp, const HaloSync& hs, const GridConsts& g, const SolverConsts& c,
                   uint32_t* __restrict__ s_list, const uint32_t lb) {
    const Team tm = team_of();
    const int64_t t = (int64_t)lb * TEAM_PARTICLES + (threadIdx.x >> 2);
    if (t >= n) return;   // whole teams leave together
    const int64_t i = first + t;
    const float4 p = x[i];
    const float w_self = poly6_in(0.f, c);
    // the particle's column of the pair list (same layout as solver.cu: block of GATHER_THREADS particles, entry k)
    const size_t pair0 = (size_t)(t / GATHER_THREADS) * PAIR_CAP * GATHER_THREADS + (size_t)(t % GATHER_THREADS);
    float rho = 0.f, gradj_l2 = 0.f, gix = 0.f, giy = 0.f, giz = 0.f;
    float w = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;   // this lane's neighbour of the current round
    int n_pairs = 0;
    team_gather(tm, p, c.h2_cull, x, soa, cell_range, g, s_list + (threadIdx.x >> 2) * TEAM_LIST,
        [&](uint32_t j, float4 q, bool valid) {
            const bool other = valid && j != (uint32_t)i;
            w = valid ? w_self : 0.f;   // the particle itself: poly6(0) at its place in the order, nothing else
            gx = gy = gz = 0.f;
            float s = 0.f;
            if (other) {
                const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
                const float r2 = sumsq(dx, dy, dz);
                w = poly6(r2, c);
                s = FAST_SPIKY ? spiky_scale_fast(r2, c) : spiky_scale(r2, c);
                gx = __fmul_rn(dx, s); gy = __fmul_rn(dy, s); gz = __fmul_rn(dz, s);
                div3_pho0(gx, gy, gz, c);
            }
            if (SAVE_PAIRS) {
                // entry index = neighbours other than itself that precede this one in the order
                const uint32_t others = (__ballot_sync(tm.mask, other) >> ((threadIdx.x & 31u) & ~3u)) & 0xfu;
                const int idx = n_pairs + __popc(others & ((1u << tm.lane) - 1u));
                if (other && idx < PAIR_CAP) pair_js[pair0 + (size_t)idx * GATHER_THREADS] = make_uint2(j, __float_as_uint(s));
                n_pairs += __popc(others);
            }
        },
        [&](int m) {
            const float wm = team_bcast(tm, w, m), ax = team_bcast(tm, gx, m), ay = team_bcast(tm, gy, m), az = team_bcast(tm, gz, m);
            rho = __fadd_rn(rho, wm);
            gix = __fadd_rn(gix, ax);
            giy = __fadd_rn(giy, ay);
            giz = __fadd_rn(giz, az);
            gradj_l2 = __fadd_rn(gradj_l2, sumsq(ax, ay, az));
        });
    if (tm.lane != 0) return;
    if (c.k_boundary != 0.f) rho = __fmaf_rn(c.k_boundary, boundary_density(p.x, p.y, p.z, g), rho);
    const float grad_l2 = __fmaf_rn(giz, giz, __fmaf_rn(giy, giy, __fmaf_rn(gix, gix, gradj_l2)));
    const float lambda = __fdiv_rn(-__fadd_rn(__fdiv_rn(rho, c.pho0), -1.f), __fadd_rn(grad_l2, c.lambda_eps));
    const float4 out = make_float4(p.x, p.y, p.z, lambda);
    xl[i] = out;
    halo_push(hp, t, out);
    rho_out[i] = rho;
    if (SAVE_PAIRS) {
        pair_cnt[t] = pair_word(n_pairs, (uint32_t)(t % GATHER_THREADS));
    }
    halo_exit(hs, lb);
}
template <bool SAVE_PAIRS, bool FAST_SPIKY>
__global__ void __launch_bounds__(TEAM_THREADS, 8)
lambda_team_kernel(const float4* __restrict__ x, const CullSoA soa, float4* __restrict__ xl, float* __restrict__ rho_out,
                   const uint2* __restrict__ cell_range, int64_t first, int64_t n,
                   uint2* __restrict__ pair_js, uint32_t* __restrict__ pair_cnt,
                   const __grid_constant__ HaloPush hp, const __grid_constant__ HaloSync hs,
                   const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    pdl_wait();   // (launch.cuh: nothing of the previous kernel is touched before this)
    extern __shared__ uint32_t s_list[];
    // (slab mode: the edge blocks first, see HaloSync)
    lambda_team_body<SAVE_PAIRS, FAST_SPIKY>(x, soa, xl, rho_out, cell_range, first, n, pair_js, pair_cnt, hp, hs, g, c, s_list, halo_block(hs));
}

// ---- delta-p replay ----------------------------------------------------------------------------------------

template <int POW>
__device__ __forceinline__ void delta_p_replay_team_body(const float4* __restrict__ xl, float4* __restrict__ x_out, const CullOut co, int64_t first, int64_t n,
                           const uint2* __restrict__ pair_js, const uint32_t* __restrict__ pair_cnt,
                           const uint2* __restrict__ cell_range, const HaloPush& hp, const HaloSync& hs, const VelTail& vt,
                           const GridConsts& g, const SolverConsts& c, const uint32_t lb) {
    const Team tm = team_of();
    halo_enter(hs, lb);   // (edge blocks: the neighbours' lambdas of this iteration are in the ghost slots of xl)
    const int64_t t = (int64_t)lb * TEAM_PARTICLES + (threadIdx.x >> 2);
    if (t >= n) return;
    const uint32_t cw = pair_cnt[t];   // (the team kernels do not re-bin: column t % GATHER_THREADS is particle t)
    const int64_t i = first + t;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cw & PAIR_OVERFLOW) {          // more neighbours than the list holds: the plain pass, on the team's first lane
        if (tm.lane == 0) out = delta_p_one<POW>(xl, (uint32_t)i, cell_range, g, c);
    } else {
        const uint32_t cnt = pair_count(cw);
        const float4 p = xl[i];
        const size_t pair0 = (size_t)(t / GATHER_THREADS) * PAIR_CAP * GATHER_THREADS + (size_t)(t % GATHER_THREADS);
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (uint32_t k0 = 0; k0 < cnt; k0 += TEAM) {
            const uint32_t k = k0 + tm.lane;
            float sc = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;   // a padding lane adds fma(0, 0, a) == a
            if (k < cnt) {
                const uint2 js = __ldg(&pair_js[pair0 + (size_t)k * GATHER_THREADS]);
                const float4 q = __ldg(&xl[js.x]);
                const float sj = __uint_as_float(js.y);
                const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
                const float pw = pow_ncorr<POW>(poly6(sumsq(dx, dy, dz), c), c);
                sc = __fmaf_rn(c.coef_corr, pw, __fadd_rn(p.w, q.w));
                tx = __fmul_rn(dx, sj); ty = __fmul_rn(dy, sj); tz = __fmul_rn(dz, sj);
            }
#pragma unroll
            for (int m = 0; m < TEAM; m++) {
                if (k0 + m < cnt) {   // (uniform in the team; a skipped fma(sc, t, a) is not the same as fma(0, 0, a) when a == -0)
                    const float scm = team_bcast(tm, sc, m);
                    ax = __fmaf_rn(scm, team_bcast(tm, tx, m), ax);
                    ay = __fmaf_rn(scm, team_bcast(tm, ty, m), ay);
                    az = __fmaf_rn(scm, team_bcast(tm, tz, m), az);
                }
            }
        }
        out = delta_p_finish(p, ax, ay, az, c);
    }
    if (tm.lane != 0) return;
    x_out[i] = out;
    co.store(i, out);
    halo_push(hp, t, out);
    if (vt.v4) velocity_tail(vt, t, i, out);
    halo_exit(hs, lb);
}
template <int POW>
__global__ void __launch_bounds__(TEAM_THREADS, 16)
delta_p_replay_team_kernel(const float4* __restrict__ xl, float4* __restrict__ x_out, const CullOut co, int64_t first, int64_t n,
                           const uint2* __restrict__ pair_js, const uint32_t* __restrict__ pair_cnt,
                           const uint2* __restrict__ cell_range, const __grid_constant__ HaloPush hp,
                           const __grid_constant__ HaloSync hs, const __grid_constant__ VelTail vt,
                           const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    pdl_wait();
    delta_p_replay_team_body<POW>(xl, x_out, co, first, n, pair_js, pair_cnt, cell_range, hp, hs, vt, g, c, halo_block(hs));
}

// ---- XSPH ----------------------------------------------------------------------------------------------------

__device__ __forceinline__ void xsph_team_body(const float4* __restrict__ x, const CullSoA soa, const float4* __restrict__ v4,
                 const uint2* __restrict__ cell_range, float* __restrict__ nvel_out,
                 const uint32_t* __restrict__ iid_sorted, uint32_t* __restrict__ iid_out, int64_t first, int64_t n,
                 const HaloSync& hs, const StatePush& sp, const GridConsts& g, const SolverConsts& c,
                 uint32_t* __restrict__ s_list, const uint32_t lb) {
    const Team tm = team_of();
    halo_enter(hs, lb);   // (edge blocks: the neighbours' velocities are in the ghost slots of v4)
    const int64_t t = (int64_t)lb * TEAM_PARTICLES + (threadIdx.x >> 2);
    if (t >= n) return;
    const int64_t i = first + t;
    const float4 p = x[i];
    const float4 vi = v4[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    float ex = 0.f, ey = 0.f, ez = 0.f;   // this lane's neighbour of the current round
    team_gather(tm, p, c.h2, x, soa, cell_range, g, s_list + (threadIdx.x >> 2) * TEAM_LIST,
        [&](uint32_t j, float4 q, bool valid) {
            ex = ey = ez = 0.f;   // itself (exact +0 per component, see solver.cu xsph_kernel) and padding lanes
            if (valid && j != (uint32_t)i) {
                const float r2 = sumsq(__fsub_rn(p.x, q.x), __fsub_rn(p.y, q.y), __fsub_rn(p.z, q.z));
                const float4 vj = __ldg(&v4[j]);
                const float w = poly6_in(r2, c);
                const float den = __fadd_rn(vi.w, vj.w);
                const float tx = __fsub_rn(vj.x, vi.x), ty = __fsub_rn(vj.y, vi.y), tz = __fsub_rn(vj.z, vi.z);
                ex = __fdiv_rn(__fmul_rn(__fadd_rn(tx, tx), w), den);
                ey = __fdiv_rn(__fmul_rn(__fadd_rn(ty, ty), w), den);
                ez = __fdiv_rn(__fmul_rn(__fadd_rn(tz, tz), w), den);
            }
        },
        [&](int m) {
            ax = __fadd_rn(ax, team_bcast(tm, ex, m));
            ay = __fadd_rn(ay, team_bcast(tm, ey, m));
            az = __fadd_rn(az, team_bcast(tm, ez, m));
        });
    if (tm.lane != 0) return;
    const float ox = __fmaf_rn(c.c_xsph, ax, vi.x), oy = __fmaf_rn(c.c_xsph, ay, vi.y), oz = __fmaf_rn(c.c_xsph, az, vi.z);
    const uint32_t id = iid_sorted[i];
    store_f3(nvel_out, t, ox, oy, oz);
    iid_out[t] = id;
    push_state_vel(sp, t, ox, oy, oz, id);
}
__global__ void __launch_bounds__(TEAM_THREADS, 8)
xsph_team_kernel(const float4* __restrict__ x, const CullSoA soa, const float4* __restrict__ v4,
                 const uint2* __restrict__ cell_range, float* __restrict__ nvel_out,
                 const uint32_t* __restrict__ iid_sorted, uint32_t* __restrict__ iid_out, int64_t first, int64_t n,
                 const __grid_constant__ HaloSync hs, const __grid_constant__ StatePush sp, const __grid_constant__ GridConsts g,
                 const __grid_constant__ SolverConsts c) {
    pdl_wait();
    extern __shared__ uint32_t s_list[];
    xsph_team_body(x, soa, v4, cell_range, nvel_out, iid_sorted, iid_out, first, n, hs, sp, g, c, s_list, halo_block(hs));
}

// ---- the persistent cooperative kernel (north-star item 3, the A/B of DESIGN.md 3.8) -------------------------------
// All solver passes of a single-GPU small-scene step in ONE launch: niter x (lambda, delta-p) — the last delta-p pass
// carrying the velocity update — and the XSPH sweep, as loops over logical blocks with a grid-wide barrier between the
// passes (cooperative groups: every block of the grid is resident). The bodies are the team kernels' own, so the bits
// are the same; what changes is that kernel boundaries become grid barriers.
struct CoopArgs {
    float4* x[2];           // position iterate, ping-pong (x[0] = what the reorder pass built)
    float* cx[2][3];        // the cull's coordinate arrays, two sets (set 0 mirrors x[0])
    float4* xl;
    float* rho;
    const uint2* cell_range;
    uint2* pair_js;
    uint32_t* pair_cnt;
    float* pos_out;  float* npos_io;  float* vel_out;  float* nvel_out;   // the caller's arrays (VelTail, XSPH)
    const uint32_t* iid_sorted;  uint32_t* iid_out;
    int64_t n;
    int32_t niter;
};
template <bool FAST_SPIKY, int POW>
__global__ void __launch_bounds__(TEAM_THREADS, 7)   // (73 registers: 8 blocks per SM spill; 7 x 148 still hold the 1000 blocks of 32 000 particles)
solve_team_coop_kernel(const __grid_constant__ CoopArgs a, const __grid_constant__ GridConsts g, const __grid_constant__ SolverConsts c) {
    pdl_wait();
    extern __shared__ uint32_t s_list[];
    cg::grid_group grid = cg::this_grid();
    const uint32_t nvb = (uint32_t)((a.n + TEAM_PARTICLES - 1) / TEAM_PARTICLES);
    const HaloPush hp;
    const HaloSync hs;
    const StatePush sp;
    int cur = 0;
    for (int it = 0; it < a.niter; it++) {
        const CullSoA soa{a.cx[cur][0], a.cx[cur][1], a.cx[cur][2]};
        for (uint32_t vb = blockIdx.x; vb < nvb; vb += gridDim.x)
            lambda_team_body<true, FAST_SPIKY>(a.x[cur], soa, a.xl, a.rho, a.cell_range, 0, a.n, a.pair_js, a.pair_cnt, hp, hs, g, c, s_list, vb);
        grid.sync();
        VelTail vt;
        if (it == a.niter - 1) {   // (pbf_stage_delta_p, fused: the velocities go to the iterate buffer this pass does not read)
            vt.rho = a.rho; vt.pos_out = a.pos_out; vt.npos_io = a.npos_io; vt.vel_out = a.vel_out; vt.v4 = a.x[cur]; vt.inv_dt = c.inv_dt;
        }
        const CullOut co{a.cx[cur ^ 1][0], a.cx[cur ^ 1][1], a.cx[cur ^ 1][2]};
        for (uint32_t vb = blockIdx.x; vb < nvb; vb += gridDim.x)
            delta_p_replay_team_body<POW>(a.xl, a.x[cur ^ 1], co, 0, a.n, a.pair_js, a.pair_cnt, a.cell_range, hp, hs, vt, g, c, vb);
        grid.sync();
        cur ^= 1;
    }
    const CullSoA soa{a.cx[cur][0], a.cx[cur][1], a.cx[cur][2]};
    for (uint32_t vb = blockIdx.x; vb < nvb; vb += gridDim.x)
        xsph_team_body(a.x[cur], soa, a.x[cur ^ 1], a.cell_range, a.nvel_out, a.iid_sorted, a.iid_out, 0, a.n, hs, sp, g, c, s_list, vb);
}

// ---- launchers -----------------------------------------------------------------------------------------------

static inline unsigned team_blocks(int64_t n) { return (unsigned)((n + TEAM_PARTICLES - 1) / TEAM_PARTICLES); }

cudaError_t preload_solver_team() {
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_team_kernel<false, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_team_kernel<true, false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_team_kernel<false, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, lambda_team_kernel<true, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_team_kernel<0>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_team_kernel<1>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_team_kernel<2>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, delta_p_replay_team_kernel<3>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, xsph_team_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, solve_team_coop_kernel<true, 3>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, solve_team_coop_kernel<false, 3>);
    return e;
}

// One cooperative launch for niter x (lambda, delta-p) + XSPH (see solve_team_coop_kernel). Returns cudaErrorNotSupported
// when the configuration is not the one the kernel is built for (n_corr == 4 with the verified trimmed powf, the
// neighbour list present, niter >= 1): the caller then takes the ordinary launches.
cudaError_t launch_solve_team_coop(float4* const x[2], CullScratch& cs, float4* xl, float* rho, const uint2* cell_range,
                                   const PairList& pl, float* pos_out, float* npos_io, float* vel_out, float* nvel_out,
                                   const uint32_t* iid_sorted, uint32_t* iid_out, int64_t n, int niter,
                                   const GridConsts& g, const SolverConsts& c, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || niter < 1 || !pl.js || !(c.n_corr == 4.0f && c.exact_pow && c.trim_pow) || cs.holds != x[0] || cs.cur != 0)
        return cudaErrorNotSupported;
    CoopArgs a;
    a.x[0] = x[0]; a.x[1] = x[1];
    for (int k = 0; k < 2; k++) { a.cx[k][0] = cs.xs[k]; a.cx[k][1] = cs.ys[k]; a.cx[k][2] = cs.zs[k]; }
    a.xl = xl; a.rho = rho; a.cell_range = cell_range; a.pair_js = pl.js; a.pair_cnt = pl.cnt;
    a.pos_out = pos_out; a.npos_io = npos_io; a.vel_out = vel_out; a.nvel_out = nvel_out;
    a.iid_sorted = iid_sorted; a.iid_out = iid_out; a.n = n; a.niter = niter;
    static int resident[2] = {0, 0};   // blocks of the kernel the device holds at once (per FAST_SPIKY variant)
    const int v = c.fast_spiky ? 1 : 0;
    if (resident[v] == 0) {
        int per_sm = 0, sms = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = c.fast_spiky ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_team_coop_kernel<true, 3>, TEAM_THREADS, TEAM_SMEM)
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_team_coop_kernel<false, 3>, TEAM_THREADS, TEAM_SMEM);
        if (e != cudaSuccess || per_sm < 1 || sms < 1) { cudaGetLastError(); return cudaErrorNotSupported; }
        resident[v] = per_sm * sms;
    }
    const unsigned nvb = team_blocks(n);
    const unsigned grid = nvb < (unsigned)resident[v] ? nvb : (unsigned)resident[v];
    void* args[3] = {(void*)&a, (void*)&g, (void*)&c};
    cudaError_t e = c.fast_spiky ? cudaLaunchCooperativeKernel((const void*)solve_team_coop_kernel<true, 3>, dim3(grid), dim3(TEAM_THREADS), args, TEAM_SMEM, st)
                                 : cudaLaunchCooperativeKernel((const void*)solve_team_coop_kernel<false, 3>, dim3(grid), dim3(TEAM_THREADS), args, TEAM_SMEM, st);
    if (e != cudaSuccess) return e;
    if (launches) (*launches)++;
    // the state the ordinary launches leave behind: the coordinates of the final iterate x[niter & 1] are in set niter & 1
    cs.cur = niter & 1;
    cs.holds = x[niter & 1];
    return cudaSuccess;
}

void launch_lambda_team(const float4* x, const CullSoA soa, float4* xl, float* rho, const uint2* cell_range, int64_t first,
                        int64_t n, uint2* pair_js, uint32_t* pair_cnt, const HaloPush& hp, HaloSync hs,
                        const GridConsts& g, const SolverConsts& c, cudaStream_t st) {
    const unsigned nb = team_blocks(n);
    halo_sync_blocks(hs, n, TEAM_PARTICLES);
    if (!pair_js && !c.fast_spiky)
        PBF_LAUNCH((lambda_team_kernel<false, false>), nb, TEAM_THREADS, TEAM_SMEM, st, x, soa, xl, rho, cell_range, first, n, nullptr, nullptr, hp, hs, g, c);
    else if (!pair_js)
        PBF_LAUNCH((lambda_team_kernel<false, true>), nb, TEAM_THREADS, TEAM_SMEM, st, x, soa, xl, rho, cell_range, first, n, nullptr, nullptr, hp, hs, g, c);
    else if (!c.fast_spiky)
        PBF_LAUNCH((lambda_team_kernel<true, false>), nb, TEAM_THREADS, TEAM_SMEM, st, x, soa, xl, rho, cell_range, first, n, pair_js, pair_cnt, hp, hs, g, c);
    else
        PBF_LAUNCH((lambda_team_kernel<true, true>), nb, TEAM_THREADS, TEAM_SMEM, st, x, soa, xl, rho, cell_range, first, n, pair_js, pair_cnt, hp, hs, g, c);
}

void launch_delta_p_replay_team(const float4* xl, float4* x_out, const CullOut co, int64_t first, int64_t n, const uint2* pair_js,
                                const uint32_t* pair_cnt, const uint2* cell_range, const HaloPush& hp, HaloSync hs,
                                const VelTail& vt, const GridConsts& g, const SolverConsts& c, int pow_mode, cudaStream_t st) {
    const unsigned nb = team_blocks(n);
    halo_sync_blocks(hs, n, TEAM_PARTICLES);
    if (pow_mode == 3) PBF_LAUNCH((delta_p_replay_team_kernel<3>), nb, TEAM_THREADS, 0, st, xl, x_out, co, first, n, pair_js, pair_cnt, cell_range, hp, hs, vt, g, c);
    else if (pow_mode == 2) PBF_LAUNCH((delta_p_replay_team_kernel<2>), nb, TEAM_THREADS, 0, st, xl, x_out, co, first, n, pair_js, pair_cnt, cell_range, hp, hs, vt, g, c);
    else if (pow_mode == 1) PBF_LAUNCH((delta_p_replay_team_kernel<1>), nb, TEAM_THREADS, 0, st, xl, x_out, co, first, n, pair_js, pair_cnt, cell_range, hp, hs, vt, g, c);
    else PBF_LAUNCH((delta_p_replay_team_kernel<0>), nb, TEAM_THREADS, 0, st, xl, x_out, co, first, n, pair_js, pair_cnt, cell_range, hp, hs, vt, g, c);
}

void launch_xsph_team(const float4* x, const CullSoA soa, const float4* v4, const uint2* cell_range, float* nvel_out,
                      const uint32_t* iid_sorted, uint32_t* iid_out, int64_t first, int64_t n, HaloSync hs, const StatePush& sp,
                      const GridConsts& g, const SolverConsts& c, cudaStream_t st) {
    halo_sync_blocks(hs, n, TEAM_PARTICLES);
    PBF_LAUNCH((xsph_team_kernel), team_blocks(n), TEAM_THREADS, TEAM_SMEM, st, x, soa, v4, cell_range, nvel_out, iid_sorted, iid_out, first, n, hs, sp, g, c);
}

}  // namespace pbf
