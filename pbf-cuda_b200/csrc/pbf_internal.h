// pbf_internal.h — types shared by the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pbf.h"

namespace pbf {

// Everything a kernel needs to know about the box and the grid for one step.
// Recomputed on the host whenever the box or h changes (reference: Simulator.cu:187-188
// recomputes m_gridHashDim every step from the current m_ulim/m_llim).
struct GridConsts {
    float llim[3];
    float ulim[3];
    float h;
    int32_t dim[3];
    int32_t ncell;   // cells of the LOCAL table: nxl * dyz (== dim[0]*dyz on a single GPU)
    int32_t dyz;     // dim[1]*dim[2]
    // x-slab of a multi-GPU decomposition (slab.cu): this handle stores cell planes
    // [xoff, xoff + nxl) of the global grid — the owned planes plus the ghost planes either
    // side. Cell ids / sort keys are local: (x - xoff)*dyz + y*dim[2] + z. Single GPU: 0, dim[0].
    int32_t xoff;
    int32_t nxl;
    uint32_t* flags;  // sticky PBF_SLAB_FLAG_* word (mapped host memory) in slab mode, else null
    // (p - llim) / h of the cell coordinate (Simulator.cu:30-35) as a reciprocal sequence VERIFIED exhaustively
    // against div.rn for |a| in [hdiv_lo, hdiv_hi] (pbf_math.cuh cell_coord, stats.cu verify_const_div); an empty
    // interval (lo > hi) keeps the plain division
    float h_rcp;      // RN(1 / h)
    float hdiv_lo, hdiv_hi;
    // Morton-ordered keys (PBF_OPT_MORTON, the A/B of DESIGN.md 3.1; single GPU): null = the reference's x-major key.
    // Else 3 x 1024 words: the bits of a cell coordinate spread to their places in the interleaved key, per axis
    // ([0..1023] x, [1024..2047] y, [2048..3071] z): key = tx[x] | ty[y] | tz[z]; ncell = the size of that key space.
    const uint32_t* morton;
};

// How the caller's particle arrays map to the sort's input order in slab mode (slab.cu).
// Physical layout [own (n_own) | from the left rank (m_left) | from the right rank]; logical
// (= tie-break) order of the stable sort [from left | own | from right], which reproduces the
// single-GPU within-cell order because every particle of the left rank preceded every own
// particle in the previous global order. Single GPU: n_own = n, m_left = 0 (identity).
struct SlabInput {
    int64_t n_own;
    int64_t m_left;
    int64_t send_left_end;     // own slots [0, send_left_end) were sent to the left rank
    int64_t send_right_begin;  // own slots [send_right_begin, n_own) were sent to the right rank
    int32_t need_left_below;   // an unsent own particle landing in a plane < this was needed left
    int32_t need_right_from;   // an unsent own particle landing in a plane >= this was needed right
    uint32_t* flags;           // sticky PBF_SLAB_FLAG_* word (mapped host memory), or null
};
__host__ __device__ inline int64_t slab_logical(const SlabInput& si, int64_t phys) {
    if (phys < si.n_own) return phys + si.m_left;
    if (phys < si.n_own + si.m_left) return phys - si.n_own;
    return phys;
}
__host__ __device__ inline int64_t slab_physical(const SlabInput& si, int64_t logical) {
    if (logical < si.m_left) return logical + si.n_own;
    if (logical < si.m_left + si.n_own) return logical - si.m_left;
    return logical;
}

// Per-launch constants of the solver kernels. Host-side values are computed exactly the way
// the reference's functor constructors compute them (Simulator.cu:77-83, 94-98, 129, 235).
struct SolverConsts {
    float h;
    float h2;          // h*h in float                      (getPoly6::h2)
    float h2_cull;     // slightly above h2: pairs below go to the exact range tests
    float poly6_coef;  // 315/(64 pi h^9)                   (getPoly6::coef)
    float spiky_coef;  // -45/(pi h^6)                      (getSpikyGrad::coef)
    float pho0;
    float lambda_eps;
    float k_boundary;
    float coef_corr;   // -k_corr / powf(poly6(dq^2), n_corr)   (Simulator.cu:235)
    float n_corr;
    float c_xsph;
    float dt;
    float inv_dt;      // 1.f/dt                            (h_updateVelocity)
    float gravity;     // g, applied as (0,0,-g)
    double lim_hi[3];  // (double)ulim - LIM_EPS            (Simulator_kernel.cuh:190-192)
    double lim_lo[3];  // (double)llim + LIM_EPS
    int32_t exact_pow; // 1: powf(w, n_corr) like the reference; 0: (w*w)^2 when n_corr == 4
    // division by the constant pho0 without the divide sequence (pbf_math.cuh div_pho0): valid for
    // |a| in [div_lo, div_hi], an interval VERIFIED exhaustively on the device against div.rn
    float pho0_rcp;    // RN(1 / pho0)
    float div_lo, div_hi;
    // 1: spiky_scale_fast (pbf_math.cuh) matched spiky_scale for EVERY float r2 in [0, h2_cull] on this device
    int32_t fast_spiky;
    // 1: pow4_trim (pbf_math.cuh) matched powf(w, 4.0f) for EVERY float w in [0, poly6(0)] on this device
    int32_t trim_pow;
};

// (key, source index) pair the radix sort moves; one 8-byte transaction per element.
struct __align__(8) KeyIdx {
    uint32_t key;
    uint32_t idx;
};

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 4;
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
constexpr int SORT_ITEMS_SMALL = 4;              // tiles of 1024 keys ...
constexpr int64_t SORT_SMALL_N = 128 * 1024;     // ... for inputs below this many keys

// Fused halo push (slab mode with attached peers): the pass that produces a float4 per owned
// particle also stores it — for the particles of the first / last ghost-width planes — straight
// into the neighbour rank's ghost slots over NVLink (peer-mapped memory), so that a halo refresh
// needs no copy kernel and no NCCL call, only a flag handshake (slab.cu). `t` below is the
// particle's index among the owned slots.
struct HaloPush {
    float4* left = nullptr;    // left peer's array (its slot 0)
    float4* right = nullptr;   // right peer's array: its left-ghost slots start at slot 0
    const int64_t* left_tail = nullptr;  // local word the left peer publishes after its sort: the slot
                                         // where its right-ghost particles begin (own_first + own_count)
    int64_t left_count = 0;    // owned t in [0, left_count) are mirrored by the left peer
    int64_t right_first = 0;   // owned t in [right_first, n) are mirrored by the right peer
};
__device__ __forceinline__ void halo_push(const HaloPush& hp, int64_t t, const float4 v) {
    if (hp.left && t < hp.left_count) hp.left[*hp.left_tail + t] = v;
    if (hp.right && t >= hp.right_first) hp.right[t - hp.right_first] = v;
}

// Fused raw-state hand-over (slab mode with attached neighbours and registered state arrays): the kernels that write
// a particle's FINAL position (update_velocity), velocity and iid (the XSPH sweep) also store them — for the owned
// particles of the planes a neighbour will need for its next step — straight into that neighbour's input arrays of
// the next step, behind its own particles, where pbf_slab_begin would otherwise copy them to with six peer-memory
// copies at the start of the next step (pbf_slab_push_state). `t` = index among the owned slots.
struct StatePush {
    float* pos_l = nullptr;  float* vel_l = nullptr;  uint32_t* iid_l = nullptr;   // the left neighbour's next input arrays
    float* pos_r = nullptr;  float* vel_r = nullptr;  uint32_t* iid_r = nullptr;   // the right neighbour's
    int64_t left_count = 0, left_dst = 0;     // owned t in [0, left_count)  -> the left neighbour's slot left_dst + t
    int64_t right_first = 0, right_dst = 0;   // owned t in [right_first, n) -> the right neighbour's slot right_dst + t - right_first
    int64_t cap_l = 0, cap_r = 0;             // slots the neighbours' arrays hold (a store beyond is dropped)
};

// In-kernel handshake of the fused halo (slab mode with attached peers; everything zero / null otherwise).
// Only the particles of a slab's first and last `ghost` planes — its two EDGES — exchange anything with a
// neighbour: they are the ones whose results are pushed into the neighbour's ghost slots, and the only ones whose
// neighbour search can reach this rank's own ghost slots. So
//   * a pass kernel runs its edge blocks FIRST (left edge, right edge, then the interior): the values the
//     neighbours are waiting for leave at the start of the kernel, not at its end;
//   * when the last edge block of a side has pushed, that block tells the neighbour of that side "refresh
//     signal_seq complete" (a release store over NVLink) — no signalling kernel;
//   * before an edge block reads ghost slots it waits until that side's neighbour has reported wait_seq (an
//     acquire spin with a time-out) — no waiting kernel; interior blocks never wait, so a late neighbour is hidden
//     behind interior work.
// Correctness of the ghost slots' reuse (a neighbour overwrites them in the next pass): a side is signalled only
// after this rank's edge blocks of that side — the only readers of those ghost slots — have finished.
struct HaloSync {
    int64_t edge_left = 0;                  // owned t <  edge_left        : left edge  (0: none)
    int64_t edge_right_first = INT64_MAX;   // owned t >= edge_right_first : right edge
    const uint32_t* wait_left = nullptr;    // local words the neighbours raise (consumer side), or null
    const uint32_t* wait_right = nullptr;
    uint32_t wait_seq = 0;
    uint32_t* peer_left = nullptr;          // the neighbours' words this rank raises (producer side), or null
    uint32_t* peer_right = nullptr;
    uint32_t signal_seq = 0;
    uint32_t* done = nullptr;               // two device counters (left, right edge blocks finished), zero between kernels
    uint64_t timeout_ns = 0;
    uint32_t* flags = nullptr;              // PBF_SLAB_FLAG_TIMEOUT goes here
    // filled by the launcher for its block size (particles per block)
    uint32_t nb = 0, nb_left = 0, nb_right = 0;
    bool on() const { return wait_left || wait_right || peer_left || peer_right; }
};
// blocks of `per_block` particles over n owned particles: which are edges. A slab so thin that the edges meet
// makes every block both.
inline void halo_sync_blocks(HaloSync& hs, int64_t n, int per_block) {
    hs.nb = hs.nb_left = hs.nb_right = 0;
    if (!hs.on() || n <= 0) return;
    const int64_t nb = (n + per_block - 1) / per_block;
    int64_t l = hs.edge_left > 0 ? (hs.edge_left + per_block - 1) / per_block : 0;
    int64_t r = hs.edge_right_first < n ? nb - (hs.edge_right_first < 0 ? 0 : hs.edge_right_first / per_block) : 0;
    if (l > nb) l = nb;
    if (l + r > nb) l = r = nb;
    hs.nb = (uint32_t)nb; hs.nb_left = (uint32_t)l; hs.nb_right = (uint32_t)r;
}

// ---- launchers (each returns the CUDA error of its launches) ---------------------------

// advect + cell key + per-pass digit histograms, in input order (advect_key.cu)
// cell_clear: the cell table to empty along the way (g.ncell entries, allocation padded to an even count), or null
cudaError_t launch_advect_key(const float* pos, const float* vel, uint32_t* keys, uint32_t* hist, uint2* cell_clear,
                              int64_t n, int npass, const SlabInput& si, const GridConsts& g,
                              const SolverConsts& c, cudaStream_t st, int64_t* launches);

// onesweep LSD radix sort of (key, idx) (radix_sort.cu). `keys` is consumed by pass 0 with the
// implicit index; result ends in bufs[result_buf].
struct SortScratch {
    uint32_t* hist;          // [MAX_PASSES][RADIX] digit counts (raw: every tile scans its pass's 256 counts)
    uint32_t* tile_counter;  // [MAX_PASSES]
    uint32_t* tile_desc;     // [npass][ntiles][RADIX] decoupled look-back state
    KeyIdx* bufs[2];
    int64_t tile_desc_words; // capacity per pass
};
cudaError_t launch_sort(const uint32_t* keys, SortScratch& s, int64_t n, int npass, const SlabInput& si,
                        int* result_buf, cudaStream_t st, int64_t* launches);
size_t sort_scratch_zero_bytes(int64_t n, int npass);
size_t sort_scratch_capacity_bytes(int64_t max_n);

// Coordinate arrays the cull of the neighbour sweeps reads (solver.cu CullSoA): max_particles + 8 floats
// each, 16-byte aligned.
struct CullScratch {
    // two sets (the delta-p kernels write the next iterate's coordinates while their overflow kernel still
    // culls on this iterate's); `cur` = the set the sweeps read
    float* xs[2] = {nullptr, nullptr};
    float* ys[2] = {nullptr, nullptr};
    float* zs[2] = {nullptr, nullptr};
    int cur = 0;
    // the float4 array whose positions (ALL stored slots) the arrays of set `cur` mirror right now, or null.
    // The kernels that produce an iterate write the coordinates along (reorder: every slot; the delta-p
    // kernels: the owned slots, which is every slot on a single GPU), so the sweeps launch pack_kernel only
    // when they find something else here — in slab mode, where ghost slots are refreshed by the neighbours.
    const float4* holds = nullptr;
};

// gather the payload into sorted SoA + cell ranges (reorder.cu)
// `n` slots are gathered (slab mode: ghosts included); pos0_out is written for the owned slots
// [own_first, own_first + own_count) only, at slot - own_first. plane_start != null (slab mode): n is only an upper
// bound (the grid), the three numbers are read on the device from the plane table: n = plane_start[g.nxl],
// own_first = plane_start[gl], own_count = plane_start[gl + nx] - own_first.
cudaError_t launch_reorder(const KeyIdx* sorted, const float* pos, const float* vel, const uint32_t* iid,
                           float4* x0, CullScratch& cs, float* pos0_out, uint32_t* iid_sorted, uint2* cell_range,
                           uint32_t* sort_zero, size_t sort_zero_bytes, int64_t n, int64_t own_first, int64_t own_count,
                           const int64_t* plane_start, int gl, int nx, const GridConsts& g,
                           const SolverConsts& c, cudaStream_t st, int64_t* launches);

// The velocity update (h_updateVelocity, Simulator.cu:127-137, 267-274) as the tail of the LAST delta-p pass of a
// step: the thread that has just produced a particle's final position also forms vel = (npos - pos) * inv_dt and
// writes what update_velocity_kernel writes — one launch and one read of the iterate fewer. v4 == null: not the
// last pass (or the caller wants the stages apart). Single-GPU steps only (slab mode pushes two halos here).
struct VelTail {
    const float* rho = nullptr;   // density of the last lambda pass, by slot
    float* pos_out = nullptr;     // caller's pos  <- step-input position (parked in npos by the reorder pass)
    float* npos_io = nullptr;     // caller's npos: step-input position in, final position out
    float* vel_out = nullptr;     // caller's vel  <- new velocity (pre-XSPH, what the reference leaves there)
    float4* v4 = nullptr;         // (vx, vy, vz, rho) by slot, what the XSPH sweep gathers
    float inv_dt = 0.f;
};

// Run-time options of the neighbour sweeps, owned by the handle (pbf_set_option): nothing on the launch path
// reads the environment.
struct SweepMode {
    int team = -1;   // -1: by particle count (solver_common.cuh TEAM_MAX_PARTICLES); 0 / 1: thread / four-lane kernels
    int rebin = 0;   // thread kernels: re-deal a block's particles by CURRENT home cell once the iterate has moved
                     // (measured: no gain, DESIGN.md 3.7 — off by default, kept selectable for the A/B)
    int staged = 0;  // first-iteration lambda pass with its candidates staged in shared memory by TMA bulk copies
                     // (the A/B of DESIGN.md 3.3: measured, not faster — off)
    int paired = 0;  // thread kernels: two consecutive slots per thread, one walk over the union of their candidate runs
                     // (solver.cu gather2): half the cull's loads per test
    int coop = 0;    // small scenes: the solver passes of pbf_step as one persistent cooperative kernel (solver_team.cu)
    int morton = 0;  // Morton-ordered keys (GridConsts::morton): thread kernels only, 27 one-cell runs instead of 9 runs
    int pdl = 1;     // programmatic dependent launch between the step's kernels (launch.cuh)
    int halo_inkernel = 1;   // fused halo: handshakes inside the pass kernels (HaloSync) instead of two one-thread kernels per refresh
    int graph = -1;  // pbf_step replayed from a CUDA graph: -1 below 256 K particles, 0 never, 1 always (pbf_capi.cu)
    bool moved = false;   // set per launch by the stage functions: the iterate is not the one the sort keyed on
};

// solver passes (solver.cu)
// Neighbour list the lambda pass saves for the delta-p pass of the same iteration (null = off).
struct PairList {
    uint2* js = nullptr;      // (slot of the k-th in-range neighbour, spiky scale of that pair as bits)
    uint32_t* cnt = nullptr;  // per list column: records | owner particle << 8 | PAIR_OVERFLOW (solver_common.cuh pair_word)
};
size_t pair_list_bytes(int64_t max_particles, size_t* js_bytes, size_t* cnt_bytes);
// The passes compute slots [first, first + n) (slab mode: the owned slots; single GPU: 0, n) and
// read neighbours from every slot. Internal arrays (x, xl, rho, v4, iid_sorted) are indexed by
// slot; caller-facing arrays (pos/npos/vel/nvel/iid) and the pair list by slot - first.
// `n_slots` = every slot the handle stores (ghosts included): the sweeps' cull reads them all.
cudaError_t launch_lambda(const float4* x, CullScratch& cs, int64_t n_slots, float4* xl, float* rho,
                          const uint2* cell_range, int64_t first, int64_t n, const PairList& pl, const HaloPush& hp, const HaloSync& hs,
                          const GridConsts& g, const SolverConsts& c, const SweepMode& mode, cudaStream_t st, int64_t* launches);
cudaError_t launch_delta_p(const float4* xl, CullScratch& cs, int64_t n_slots, float4* x_out, const uint2* cell_range,
                           int64_t first, int64_t n, const PairList& pl, const HaloPush& hp, const HaloSync& hs, const VelTail& vt,
                           const GridConsts& g, const SolverConsts& c, const SweepMode& mode, cudaStream_t st, int64_t* launches,
                           uint32_t block0 = 0, uint32_t nblk = 0, bool last_slice = true);
bool delta_p_sliceable(const PairList& pl, const SweepMode& mode, int64_t n);
bool sweeps_use_team(const SweepMode& mode, int64_t n);   // whether the sweeps of n particles take the four-lane kernels
// solver_team.cu: niter x (lambda, delta-p with the velocity update on the last) + XSPH of a single-GPU small-scene step
// as ONE persistent cooperative kernel (PBF_OPT_COOP). cudaErrorNotSupported: not this configuration, launch as usual.
cudaError_t launch_solve_team_coop(float4* const x[2], CullScratch& cs, float4* xl, float* rho, const uint2* cell_range,
                                   const PairList& pl, float* pos_out, float* npos_io, float* vel_out, float* nvel_out,
                                   const uint32_t* iid_sorted, uint32_t* iid_out, int64_t n, int niter,
                                   const GridConsts& g, const SolverConsts& c, cudaStream_t st, int64_t* launches);
cudaError_t launch_update_velocity(const float4* x, const float* rho, float* pos_out, float* npos_io,
                                   float* vel_out, float4* v4, int64_t first, int64_t n, const HaloPush& hp,
                                   const HaloSync& hs, const StatePush& sp, const SolverConsts& c, cudaStream_t st, int64_t* launches);
// slab.cu: flag handshake of a fused halo refresh. signal: store `seq` (release, system scope) to
// a word in a peer's memory; wait: spin until the local word reaches `seq`, give up after
// `timeout_ns` and raise PBF_SLAB_FLAG_TIMEOUT instead of hanging the device.
cudaError_t launch_halo_signal(uint32_t* peer_word_left, uint32_t* peer_word_right, uint32_t seq, cudaStream_t st,
                               int64_t* launches);
// the same signal, after first publishing *tail_src (a device word: where this rank's right-ghost
// slots begin) into the right neighbour's memory
cudaError_t launch_halo_publish(const int64_t* tail_src, int64_t* peer_right_tail, uint32_t* peer_word_left,
                                uint32_t* peer_word_right, uint32_t seq, cudaStream_t st, int64_t* launches);
cudaError_t launch_halo_wait(const uint32_t* word_left, const uint32_t* word_right, uint32_t seq,
                             uint64_t timeout_ns, uint32_t* flags, cudaStream_t st, int64_t* launches);
cudaError_t launch_xsph(const float4* x, CullScratch& cs, int64_t n_slots, const float4* v4,
                        const uint2* cell_range, float* nvel_out, const uint32_t* iid_sorted, uint32_t* iid_out,
                        int64_t first, int64_t n, const HaloSync& hs, const StatePush& sp, const GridConsts& g, const SolverConsts& c,
                        const SweepMode& mode, cudaStream_t st, int64_t* launches);
// slab mode with in-kernel halo handshakes: the cull's coordinates of the GHOST slots [0, own_first) and
// [own_first + own_count, n_slots) of `x` (the owned slots were written by the pass that produced x); its blocks
// wait for the neighbours' wait_seq first
cudaError_t launch_pack_ghosts(const float4* x, CullScratch& cs, int64_t n_slots, int64_t own_first, int64_t own_count,
                               const HaloSync& hs, cudaStream_t st, int64_t* launches);
// slab.cu: first slot of every local plane in the sorted pairs (nxl + 1 entries, the last one =
// number of particles inside the local plane range; the rest carry the discard key)
cudaError_t launch_plane_table(const KeyIdx* sorted, int64_t n, int64_t* plane_start, const GridConsts& g,
                               cudaStream_t st, int64_t* launches);
// slab.cu: npos/nvel/iid_out[s] = pos/vel/iid[sorted[s].idx] (state sort without a step)
cudaError_t launch_gather_state(const KeyIdx* sorted, const float* pos, const float* vel, const uint32_t* iid,
                                float* npos, float* nvel, uint32_t* iid_out, int64_t n, cudaStream_t st,
                                int64_t* launches);
cudaError_t launch_neighbor_count(const float4* x, CullScratch& cs, const uint2* cell_range, uint32_t* count,
                                  int64_t n, const GridConsts& g, const SolverConsts& c, cudaStream_t st);

// stats.cu: exhaustive check of the reciprocal division sequence for divisor d over all 2^32 bit
// patterns of the dividend; returns the verified interval of |a| around 1 (lo > hi: none)
cudaError_t verify_const_div(float d, float rcp, float* lo, float* hi, void* scratch8, cudaStream_t st);
// stats.cu: exhaustive comparison of spiky_scale_fast with spiky_scale over every float r2 in [0, top]
cudaError_t verify_spiky(const SolverConsts& c, float top, unsigned long long* mismatches, void* scratch8, cudaStream_t st);
cudaError_t verify_pow4(float top, unsigned long long* mismatches, void* scratch8, cudaStream_t st);

// force-load every kernel of a translation unit (see the comment at preload_solver in solver.cu)
cudaError_t preload_advect_key();
cudaError_t preload_sort();
cudaError_t preload_reorder();
cudaError_t preload_scene();
cudaError_t preload_slab();
cudaError_t preload_solver();
cudaError_t preload_stats();

// scene + stats (scene.cu, stats.cu)
cudaError_t launch_scene_block(const float origin[3], const int32_t n3[3], float spacing, uint32_t seed,
                               uint32_t first_iid, int32_t ix_begin, int32_t ix_end, float* pos, float* vel,
                               uint32_t* iid, cudaStream_t st);
void scene_block_host(const float origin[3], const int32_t n3[3], float spacing, uint32_t seed,
                      uint32_t first_iid, int32_t ix_begin, int32_t ix_end, float* pos, float* vel,
                      uint32_t* iid);
// stats.cu: order-independent 128-bit digest of (iid, pos, vel) (include/pbf.h pbf_state_digest_device / _host)
cudaError_t launch_digest(const float* pos, const float* vel, const uint32_t* iid, int64_t n, unsigned long long* out,
                          cudaStream_t st);
void digest_host(const float* pos, const float* vel, const uint32_t* iid, int64_t n, uint64_t out[2]);
cudaError_t launch_stats(const float* rho, const float* npos, const float* nvel, int64_t n, float pho0,
                         double* partial, int nblocks, cudaStream_t st);

}  // namespace pbf
