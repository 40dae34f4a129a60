// solver_common.cuh — what the neighbour-sweep kernels of solver.cu (one thread per particle) and
// solver_team.cu (four lanes per particle, small scenes) share: sizes, the cull's coordinate arrays and its
// two-lane FP32 test, the neighbour-list layout, the tails of the delta-p pass.
#pragma once
#include "pbf_math.cuh"

namespace pbf {

#ifndef PBF_GATHER_MINBLOCKS
#define PBF_GATHER_MINBLOCKS 8
#endif
#ifndef PBF_WORD_CAP
#define PBF_WORD_CAP 15
#endif
#ifndef PBF_PAIR_CAP
#define PBF_PAIR_CAP 96
#endif
#ifndef PBF_GATHER_THREADS
#define PBF_GATHER_THREADS 128
#endif
constexpr int GATHER_THREADS = PBF_GATHER_THREADS;
constexpr int WORD_CAP = PBF_WORD_CAP;  // hit words (32 candidates each) buffered per thread before a flush
constexpr int PAIR_CAP = PBF_PAIR_CAP;  // neighbours per particle the lambda pass can hand to the delta-p pass
constexpr size_t LIST_SMEM = (size_t)WORD_CAP * GATHER_THREADS * sizeof(uint2);  // 15 KB per CTA

constexpr uint32_t PAIR_OVERFLOW = 1u << 31;   // pair_cnt: more than PAIR_CAP neighbours, the delta-p pass gathers
// pair_cnt word of list column c of a block: bits 0-7 the number of records, bits 8-14 WHICH of the block's
// GATHER_THREADS particles the column belongs to (the re-binned sweeps deal a block's particles to its threads in the
// order of their current home cell, see rebin_block in solver.cu; otherwise it is c itself), bit 31 PAIR_OVERFLOW.
static_assert(PAIR_CAP < 256 && GATHER_THREADS <= 128, "pair_cnt word: 8 bits of count, 7 bits of particle");
__device__ __forceinline__ uint32_t pair_word(int n_pairs, uint32_t local) {
    return (n_pairs <= PAIR_CAP ? (uint32_t)n_pairs : PAIR_OVERFLOW) | (local << 8);
}
// The list is written once by the lambda pass and read once by the delta-p replay, 8 bytes x ~40 per particle: it
// streams through the caches and must not push out what the sweeps gather from (positions, cull coordinates, cell
// table — a few tens of MB that otherwise live in L2). PBF_LIST_STREAM: records stored / loaded with the streaming
// (evict-first) cache operator. Measured: no effect (dam_1m 1.7694 vs 1.7650 ms per step, 16 M 24.327 vs 24.334) — the
// gathers' L2 misses are not what the sweeps wait for; off.
#ifndef PBF_LIST_STREAM
#define PBF_LIST_STREAM 0
#endif
__device__ __forceinline__ void list_store(uint2* p, uint32_t j, uint32_t s) {
#if PBF_LIST_STREAM
    asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(j), "r"(s) : "memory");
#else
    *p = make_uint2(j, s);
#endif
}
__device__ __forceinline__ uint2 list_load(const uint2* p) {
#if PBF_LIST_STREAM
    uint2 v;
    asm volatile("ld.global.cs.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint32_t pair_count(uint32_t w) { return w & 0xffu; }
__device__ __forceinline__ uint32_t pair_local(uint32_t w) { return (w >> 8) & 0x7fu; }

// Cull-side copy of the positions: three float arrays (structure of arrays), written along with the float4
// iterate by the kernels that produce it (CullOut below; by pack_kernel in slab mode, where the neighbours
// fill the ghost slots). Four consecutive candidates are then three 16-byte loads
// (instead of four), and their coordinates sit in adjacent registers, which is what the packed FP32
// instructions of sm_100 want.
struct CullSoA {
    const float* xs;
    const float* ys;
    const float* zs;
};
// the same arrays for the kernels that PRODUCE an iterate (reorder, the delta-p kernels): they write the
// coordinates along, so that the next sweep needs no pack_kernel (pbf_internal.h CullScratch::holds)
struct CullOut {
    float* xs;
    float* ys;
    float* zs;
    __device__ __forceinline__ void store(int64_t i, const float4 q) const { xs[i] = q.x; ys[i] = q.y; zs[i] = q.z; }
};

// two candidates: RN(r2 - limit) of each, sign bits pushed into `hits` (first candidate first)
__device__ __forceinline__ uint32_t push_hits2(uint32_t hits, f32x2 px, f32x2 py, f32x2 pz, f32x2 lim,
                                               float x0, float x1, float y0, float y1, float z0, float z1) {
    const f32x2 dx = sub2(px, pack2(x0, x1)), dy = sub2(py, pack2(y0, y1)), dz = sub2(pz, pack2(z0, z1));
    const f32x2 t = sub2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), lim);
    uint32_t t0, t1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(t));
    hits = __funnelshift_l(t0, hits, 1);
    return __funnelshift_l(t1, hits, 1);
}

// w^n_corr of the delta-p pass's s_corr (Simulator_kernel.cuh:166 powf(..., n_corr)). POW = 1: powf with
// the run-time exponent, exactly the call inside the reference; POW = 2: the same libdevice powf with the
// exponent known to be 4.0f (the default n_corr) — the compiler folds the exponent-dependent parts of the
// routine (~16 of ~84 instructions), the arithmetic and hence the bits are the same; POW = 3: that routine's
// arithmetic without its special-case tests (pbf_math.cuh pow4_trim), used only after it matched powf(w, 4.0f)
// for every w the pass can produce on this device (SolverConsts::trim_pow); POW = 0: (w*w)^2,
// opt-in (pbf_set_option_exact_pow(0)), within 1e-5 but not bit-identical.
template <int POW>
__device__ __forceinline__ float pow_ncorr(float w, const SolverConsts& c) {
    if (POW == 1) return powf(w, c.n_corr);
    if (POW == 2) return powf(w, 4.0f);
    if (POW == 3) return pow4_trim(w);
    const float w2 = __fmul_rn(w, w);
    return __fmul_rn(w2, w2);
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

// ---- HaloSync (pbf_internal.h): the in-kernel handshake of the fused halo -------------------------------------
// logical block of this CTA: the edges first (see HaloSync)
__device__ __forceinline__ uint32_t halo_block(const HaloSync& hs) {
    const uint32_t b = blockIdx.x;
    if (hs.nb == 0 || b < hs.nb_left) return b;
    if (b < hs.nb_left + hs.nb_right) return hs.nb - hs.nb_right + (b - hs.nb_left);
    return b - hs.nb_right;
}
__device__ __forceinline__ bool halo_is_left(const HaloSync& hs, uint32_t lb) { return lb < hs.nb_left; }
__device__ __forceinline__ bool halo_is_right(const HaloSync& hs, uint32_t lb) { return hs.nb_right && lb >= hs.nb - hs.nb_right; }
__device__ __forceinline__ void halo_spin(const uint32_t* w, uint32_t seq, uint64_t timeout_ns, uint32_t* flags) {
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w) : "memory");
        if ((int32_t)(v - seq) >= 0) return;   // (int32 difference: the sequence number may wrap)
        uint64_t t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {            // a dead neighbour must not hang the device
            if (flags) atomicOr(flags, (uint32_t)PBF_SLAB_FLAG_TIMEOUT);
            return;
        }
        __nanosleep(64);
    }
}
// consumer side: called by ALL threads of the block before anything reads a ghost slot
__device__ __forceinline__ void halo_enter(const HaloSync& hs, uint32_t lb) {
    const bool wl = hs.wait_left && halo_is_left(hs, lb), wr = hs.wait_right && halo_is_right(hs, lb);
    if (!(wl || wr)) return;   // (uniform in the block)
    if (threadIdx.x == 0) {
        if (wl) halo_spin(hs.wait_left, hs.wait_seq, hs.timeout_ns, hs.flags);
        if (wr) halo_spin(hs.wait_right, hs.wait_seq, hs.timeout_ns, hs.flags);
    }
    __syncthreads();
}
// producer side: called by every thread of the block that is still alive, behind its last push (thread 0 always is)
__device__ __forceinline__ void halo_exit(const HaloSync& hs, uint32_t lb) {
    const bool sl = hs.peer_left && halo_is_left(hs, lb), sr = hs.peer_right && halo_is_right(hs, lb);
    if (!(sl || sr)) return;   // (uniform in the block)
    __threadfence_system();    // my pushes are performed before ...
    __syncthreads();           // ... thread 0 counts the block as done
    if (threadIdx.x == 0) {
        if (sl && atomicAdd(hs.done + 0, 1u) == hs.nb_left - 1) {
            hs.done[0] = 0;
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(hs.peer_left), "r"(hs.signal_seq) : "memory");
        }
        if (sr && atomicAdd(hs.done + 1, 1u) == hs.nb_right - 1) {
            hs.done[1] = 0;
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(hs.peer_right), "r"(hs.signal_seq) : "memory");
        }
    }
}

// StatePush (pbf_internal.h): a boundary particle's final position / velocity + iid into the neighbours' next input
__device__ __forceinline__ void push_state_pos(const StatePush& sp, int64_t t, float x, float y, float z) {
    if (sp.pos_l && t < sp.left_count && sp.left_dst + t < sp.cap_l) store_f3(sp.pos_l, sp.left_dst + t, x, y, z);
    if (sp.pos_r && t >= sp.right_first && sp.right_dst + (t - sp.right_first) < sp.cap_r) store_f3(sp.pos_r, sp.right_dst + (t - sp.right_first), x, y, z);
}
__device__ __forceinline__ void push_state_vel(const StatePush& sp, int64_t t, float x, float y, float z, uint32_t id) {
    if (sp.vel_l && t < sp.left_count && sp.left_dst + t < sp.cap_l) {
        store_f3(sp.vel_l, sp.left_dst + t, x, y, z);
        sp.iid_l[sp.left_dst + t] = id;
    }
    if (sp.vel_r && t >= sp.right_first && sp.right_dst + (t - sp.right_first) < sp.cap_r) {
        store_f3(sp.vel_r, sp.right_dst + (t - sp.right_first), x, y, z);
        sp.iid_r[sp.right_dst + (t - sp.right_first)] = id;
    }
}

// VelTail (pbf_internal.h): the velocity update of particle t (slot i) right behind its final position `q`
__device__ __forceinline__ void velocity_tail(const VelTail& vt, int64_t t, int64_t i, const float4 q) {
    const float3 p0 = load_f3(vt.npos_io, t);
    const float vx = __fmul_rn(__fsub_rn(q.x, p0.x), vt.inv_dt);
    const float vy = __fmul_rn(__fsub_rn(q.y, p0.y), vt.inv_dt);
    const float vz = __fmul_rn(__fsub_rn(q.z, p0.z), vt.inv_dt);
    vt.v4[i] = make_float4(vx, vy, vz, vt.rho[i]);
    store_f3(vt.vel_out, t, vx, vy, vz);
    store_f3(vt.pos_out, t, p0.x, p0.y, p0.z);
    store_f3(vt.npos_io, t, q.x, q.y, q.z);
}

// The delta-p pass of ONE particle the plain way — every candidate of the 27 cells in the reference's visiting
// order, the exact pair arithmetic for those in range — for the rare particle whose neighbour list did not fit
// PAIR_CAP records (a collapsing cluster): the replay kernels call it instead of a second, list-less kernel
// that every Jacobi iteration would have to launch just to find nothing to do. Candidates outside h contribute
// exact zeros in the reference (computetpos, Simulator_kernel.cuh:150-170), so skipping them changes no bit.
// Not inlined: the replay kernels keep their 32 registers.
template <int POW>
__device__ __noinline__ float4 delta_p_one(const float4* __restrict__ xl, uint32_t i, const uint2* __restrict__ cell_range,
                                           const GridConsts& g, const SolverConsts& c) {
    const float4 p = xl[i];
    const int3 cc = cell_of(p.x, p.y, p.z, g);
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int dx = -1; dx <= 1; dx++) {
        const int cx = cc.x + dx, lx = cx - g.xoff;
        if (cx < 0 || cx >= g.dim[0]) continue;
        if (lx < 0 || lx >= g.nxl) {   // the search leaves the stored planes (slab mode): say so, see gather()
            if (g.flags) atomicOr(g.flags, (uint32_t)PBF_SLAB_FLAG_GHOST);
            continue;
        }
        for (int dy = -1; dy <= 1; dy++) {
            const int cy = cc.y + dy;
            if (cy < 0 || cy >= g.dim[1]) continue;
            for (int dz = -1; dz <= 1; dz++) {
                const int cz = cc.z + dz;
                if (cz < 0 || cz >= g.dim[2]) continue;
                const uint2 r = __ldg(&cell_range[cell_id(cx, cy, cz, g)]);   // (x-major: lx * dyz + cy * dim z + cz)
                for (uint32_t j = r.x; j < r.y; j++) {
                    if (j == i) continue;
                    const float4 q = __ldg(&xl[j]);
                    const float ddx = __fsub_rn(p.x, q.x), ddy = __fsub_rn(p.y, q.y), ddz = __fsub_rn(p.z, q.z);
                    const float r2 = sumsq(ddx, ddy, ddz);
                    if (!(r2 < c.h2_cull)) continue;
                    const float pw = pow_ncorr<POW>(poly6(r2, c), c);
                    const float sc = __fmaf_rn(c.coef_corr, pw, __fadd_rn(p.w, q.w));
                    const float s = spiky_scale(r2, c);
                    ax = __fmaf_rn(sc, __fmul_rn(ddx, s), ax);
                    ay = __fmaf_rn(sc, __fmul_rn(ddy, s), ay);
                    az = __fmaf_rn(sc, __fmul_rn(ddz, s), az);
                }
            }
        }
    }
    return delta_p_finish(p, ax, ay, az, c);
}

// solver_team.cu: the same sweeps with four lanes per particle, for scenes too small to fill the machine
// Measured (B200, ms per step, thread-per-particle vs team): 32 K 0.380 vs 0.282; 131 K 0.599 vs 0.871; 262 K 0.893 vs
// 1.490 — the team kernels stop being latency bound near 45 K particles and cost ~1.6x the instruction slots from
// there on, the thread kernels stay near their latency floor up to ~130 K: the crossover is near 60 K.
constexpr int64_t TEAM_MAX_PARTICLES = 48 * 1024;
cudaError_t preload_solver_team();
void launch_lambda_team(const float4* x, const CullSoA soa, float4* xl, float* rho, const uint2* cell_range, int64_t first,
                        int64_t n, uint2* pair_js, uint32_t* pair_cnt, const HaloPush& hp, HaloSync hs,
                        const GridConsts& g, const SolverConsts& c, cudaStream_t st);
void launch_delta_p_replay_team(const float4* xl, float4* x_out, const CullOut co, int64_t first, int64_t n, const uint2* pair_js,
                                const uint32_t* pair_cnt, const uint2* cell_range, const HaloPush& hp, HaloSync hs,
                                const VelTail& vt, const GridConsts& g, const SolverConsts& c, int pow_mode, cudaStream_t st);
void launch_xsph_team(const float4* x, const CullSoA soa, const float4* v4, const uint2* cell_range, float* nvel_out,
                      const uint32_t* iid_sorted, uint32_t* iid_out, int64_t first, int64_t n, HaloSync hs, const StatePush& sp,
                      const GridConsts& g, const SolverConsts& c, cudaStream_t st);

}  // namespace pbf
