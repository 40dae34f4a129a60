// pbf_capi.cu — the C-ABI of include/pbf.h: handle lifetime, parameters, the step and its stages.
//
// Host-side mirror of the reference's Simulator (fluids/Simulator.h, Simulator.cpp): same five
// stages in the same order; GL interop replaced by raw device pointers; the process-wide
// GUIParams singleton replaced by a parameter block in the handle; checkCudaErrors' print+exit
// replaced by status codes.
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <mutex>
#include <new>

#include "launch.cuh"
#include "pbf_internal.h"

using namespace pbf;

namespace pbf {
thread_local bool tl_pdl = true;   // launch.cuh: set from the handle's PBF_OPT_PDL by every stage function
}

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail(PBF_ERR_CUDA, "CUDA error at %s:%d code=%d(%s) \"%s\"", __FILE__, __LINE__, \
                        (int)e__, cudaGetErrorName(e__), #expr);                                 \
    } while (0)

// dst[i*width + k] = src[i*stride + offset + k]: de-interleaves one field of an internal
// SoA-of-struct array into a tight array for the parity read-backs.
__global__ void extract_words_kernel(const uint32_t* __restrict__ src, int stride, int offset, int width,
                                     uint32_t* __restrict__ dst, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * width) return;
    const int64_t i = t / width;
    const int k = (int)(t % width);
    dst[t] = src[i * stride + offset + k];
}

enum Stage { ST_IDLE = 0, ST_BOUND, ST_ADVECTED, ST_GRID, ST_LAMBDA, ST_DENSITY, ST_VELOCITY, ST_XSPH };

constexpr int GRAPH_SLOTS = 4;
constexpr int64_t GRAPH_AUTO_MAX = 256 * 1024;   // PBF_OPT_GRAPH = -1: graphs below this many particles

// what a step's launches depend on (key) and what the stage functions leave in the handle (post)
struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    uint64_t used = 0;
    const void* ptr[5] = {};
    int64_t n = 0;
    cudaStream_t stream = nullptr;
    uint64_t consts_hash = 0;
    int niter = 0, team = 0, rebin = 0, pdl = 0;
    int64_t launches = 0;
    // post-step handle state
    int sorted_buf = 0, cur = 0, iters_done = 0, cull_cur = 0;
    const float4* cull_holds = nullptr;
    float4* v4 = nullptr;
};

}  // namespace

struct pbf_sim {
    int device = 0;
    pbf_params p{};
    float ulim[3]{}, llim[3]{};
    int64_t max_particles = 0;
    int64_t cell_capacity = 0;
    int exact_pow = 1;
    SweepMode mode;   // run-time options of the neighbour sweeps (pbf_set_option)

    // scratch (device)
    uint32_t* keys = nullptr;
    // [hist | tile counters | tile descriptors] of the sort. Invariant: all zero whenever no sort is in flight —
    // zeroed at create, and whoever sorts cleans up behind itself (the step: inside the reorder kernel)
    uint32_t* sort_zero = nullptr;
    size_t sort_zero_capacity = 0;
    bool sort_dirty = false;        // advect_key ran, the cleaning reorder has not yet (a step abandoned half-way)
    KeyIdx* pairs[2] = {nullptr, nullptr};
    float4* x[2] = {nullptr, nullptr};
    float4* xl = nullptr;   // (x, y, z, lambda); reused as (vx, vy, vz, rho) by a separate velocity update
    float4* v4 = nullptr;   // where (vx, vy, vz, rho) of the step in flight is: xl, or — when the velocity update rode
                            // along with the last delta-p pass — the iterate buffer that pass no longer read
    bool fuse_velocity = false;   // set by pbf_step for its last Jacobi iteration
    float* rho = nullptr;
    uint32_t* iid_sorted = nullptr;
    uint2* cell_range = nullptr;
    PairList pairs_list;            // lambda -> delta-p neighbour list (null when disabled / too large)
    CullScratch cull;               // coordinate arrays of the sweeps' cull (solver.cu pack_kernel)
    uint32_t* count_scratch = nullptr;
    uint32_t* read_scratch = nullptr;
    double* stats_partial = nullptr;
    double* stats_host = nullptr;
    // staging for pbf_step_host
    float* h_pos = nullptr; float* h_npos = nullptr; float* h_vel = nullptr; float* h_nvel = nullptr;
    uint32_t* h_iid = nullptr;
    // pbf_step_host: its own two non-blocking streams, so that the download of the final positions and of
    // iid runs on the copy engine while the XSPH sweep still computes the velocities
    cudaStream_t host_main = nullptr, host_copy = nullptr;
    cudaEvent_t host_ev = nullptr, host_iid_ev = nullptr;
    cudaEvent_t reorder_wait = nullptr;   // step_host: the iid upload, still in flight while the keys are sorted
    cudaEvent_t layout_ev[2] = {nullptr, nullptr};   // slab mode: plane table ready / downloaded (side stream)

    // bound state of the step in flight
    Stage stage = ST_IDLE;
    float *pos = nullptr, *npos = nullptr, *vel = nullptr, *nvel = nullptr;
    uint32_t* iid = nullptr;
    int64_t n = 0;
    cudaStream_t stream = nullptr;
    GridConsts g{};
    SolverConsts c{};
    int npass = 1;
    int sorted_buf = 0;
    int cur = 0;         // x[cur] holds the current iterate
    int iters_done = 0;
    bool pos0_in_npos = false;

    // x-slab decomposition (multi-GPU); off for the plain single-GPU handle
    bool slab_on = false;
    pbf_slab_step slab{};
    SlabInput si{};
    int64_t n_local = 0;      // slots the solver arrays hold this step (single GPU: n)
    int64_t own_first = 0;    // first owned slot                        (single GPU: 0)
    int64_t own_count = 0;    // owned particles                         (single GPU: n)
    int32_t ghost_left = 0;   // ghost planes actually stored left of x_begin
    int64_t* plane_dev = nullptr;
    int64_t* plane_host = nullptr;   // pinned copy: first slot of every local plane, + total
    int64_t plane_capacity = 0;
    uint32_t* flags_host = nullptr;  // mapped pinned word the kernels raise PBF_SLAB_FLAG_* in
    uint32_t* flags_dev = nullptr;
    pbf_slab_layout layout{};
    bool layout_valid = false;
    // fused halo over peer memory
    struct Peer {
        bool on = false;
        float4* x[2] = {nullptr, nullptr};
        float4* xl = nullptr;
        uint32_t* sync = nullptr;
        float* state[4] = {nullptr, nullptr, nullptr, nullptr};    // pos A, pos B, vel A, vel B (if registered)
        uint32_t* iid = nullptr;
        int64_t state_capacity = 0;                                // particles its state arrays hold
        void* ipc_base[9] = {};                                    // opened IPC mappings to close
    } peer[2];
    // Morton-ordered keys (PBF_OPT_MORTON): spread tables of the current grid dimensions on the device
    uint32_t* morton_dev = nullptr;
    int32_t morton_dims[3] = {0, 0, 0};
    int morton_bits = 0;
    float* tail_npos_host = nullptr;   // pbf_step_host: where the final positions go (the last delta-p pass sends them in slices)
    bool tail_npos_sent = false;
    StatePush state_push;            // armed by pbf_slab_push_state for this step's velocity / XSPH kernels
    float* state[4] = {nullptr, nullptr, nullptr, nullptr};        // my registered pos A, pos B, vel A, vel B
    uint32_t* state_iid = nullptr;
    uint32_t state_seq = 0;          // handshake number of my neighbours' "state complete" signal to wait for
    uint32_t* sync_words = nullptr;   // device: [0] raised by the left neighbour, [1] by the right one,
                                      // [2..3] int64 published by the left neighbour: its first right-ghost slot
    uint32_t halo_seq = 0;
    // in-kernel handshakes of the fused halo (HaloSync, pbf_internal.h)
    uint32_t* halo_done = nullptr;   // device: edge blocks finished, left / right (zero between kernels)
    int64_t edge_left = 0, edge_right_first = 0;   // owned particles t < / >= these form the slab's two edges (HaloSync)
    uint32_t wait_seq = 0;           // the handshake number of the last pushing pass: what the next reader waits for
    bool wait_valid = false;
    uint64_t halo_timeout_ns = 10ull * 1000 * 1000 * 1000;

    // the exhaustive verifications below run on a stream of their own with result words allocated once (a
    // parameter change costs no device-wide synchronisation), and what they found is remembered per value:
    // toggling a GUI slider back is free (the reference reloads its parameters every frame, Simulator.cpp:101-115)
    cudaStream_t verify_stream = nullptr;
    void* verify_scratch = nullptr;
    struct DivEntry { float d, lo, hi; } div_cache[4] = {};
    struct SpikyEntry { float h; int ok; unsigned long long bad; } spiky_cache[4] = {};
    struct PowEntry { float top; int ok; unsigned long long bad; } pow_cache[4] = {};
    int div_n = 0, spiky_n = 0, pow_n = 0;   // entries filled (round robin beyond 4)
    // verified constant division (refresh_consts)
    bool div_verified = false;
    float div_d = 0.f, div_rcp = 0.f, div_lo = 1.f, div_hi = 0.f;
    bool hdiv_verified = false;      // the same for the division by h of the cell coordinate
    float hdiv_d = 0.f, hdiv_rcp = 0.f, hdiv_lo = 1.f, hdiv_hi = 0.f;
    // branch-free spiky scale (pbf_math.cuh spiky_scale_fast): verified exhaustively for this h, or off
    bool spiky_checked = false;
    float spiky_h = 0.f;
    int spiky_ok = 0;
    unsigned long long spiky_mismatches = 0;
    // trimmed powf(w, 4.0f) (pbf_math.cuh pow4_trim): verified exhaustively up to this poly6(0), or off
    bool pow4_checked = false;
    float pow4_top = 0.f;
    int pow4_ok = 0;
    unsigned long long pow4_mismatches = 0;

    // pbf_step as a CUDA graph (PBF_OPT_GRAPH): a few instantiated graphs keyed by everything a step's launches
    // depend on — the caller's pointers (two keys under the caller's ping-pong), n, the stream, the constants
    StepGraph graphs[GRAPH_SLOTS];
    uint64_t graph_clock = 0;
    int graph_misses = 0;      // consecutive steps that had to capture
    int graph_holdoff = 0;     // steps to run without graphs after the key kept changing (a moving wall)
    bool graph_broken = false; // capture failed once on this handle: plain launches from then on

    int64_t launches = 0;
    bool timing = false;
    cudaEvent_t ev[6] = {};
    cudaEvent_t kev[2 * PBF_KERNEL_SLOTS] = {};  // [2k] before / [2k+1] after kernel slot k
    bool ev_valid = false;
};

namespace {

int key_bits(int64_t ncell) {
    int b = 0;
    while (((int64_t)1 << b) < ncell) b++;
    return b < 1 ? 1 : b;
}

// Grid dims as the reference recomputes them every step (Simulator.cu:187-188).
void compute_dim(const float ulim[3], const float llim[3], float h, int32_t dim[3]) {
    for (int a = 0; a < 3; a++) {
        float diff = ulim[a] - llim[a];
        dim[a] = (int32_t)ceilf(diff / h);
    }
}

// The interval of |a| in which a / d as the reciprocal sequence q = a*y, q' = fma(fma(-q, d, a), y, q) equals div.rn
// for EVERY float a on this device (stats.cu verify_const_div, remembered per divisor), or the empty interval
// (lo > hi: always the plain division) when it is switched off, cannot be checked, or is not comfortably wide.
void verified_div_interval(pbf_sim* s, float d, float rcp, float* out_lo, float* out_hi) {
    *out_lo = 1.f; *out_hi = 0.f;
    const char* off = getenv("PBF_NO_CONST_DIV");
    if ((off && off[0] == '1') || !(d > 0.f) || !(d < 3.0e38f) || cudaSetDevice(s->device) != cudaSuccess) return;
    float lo = 1.f, hi = 0.f;
    bool known = false;
    for (int k = 0; k < (s->div_n < 4 ? s->div_n : 4); k++)
        if (s->div_cache[k].d == d) { lo = s->div_cache[k].lo; hi = s->div_cache[k].hi; known = true; }
    if (!known) {
        if (verify_const_div(d, rcp, &lo, &hi, s->verify_scratch, s->verify_stream) == cudaSuccess) {
            s->div_cache[s->div_n++ % 4] = {d, lo, hi};
        } else {
            cudaGetLastError();
            lo = 1.f; hi = 0.f;
        }
    }
    // use it only if the verified interval is comfortably wide around the values that occur
    if (lo <= 1e-30f && hi >= 1e30f) { *out_lo = lo; *out_hi = hi; }
}

int refresh_consts(pbf_sim* s) {
    GridConsts& g = s->g;
    SolverConsts& c = s->c;
    const pbf_params& p = s->p;
    if (!(p.h > 0.f) || !(p.dt > 0.f) || p.niter < 0) return fail(PBF_ERR_INVALID, "bad parameters (h, dt, niter)");
    for (int a = 0; a < 3; a++) { g.llim[a] = s->llim[a]; g.ulim[a] = s->ulim[a]; }
    g.h = p.h;
    compute_dim(s->ulim, s->llim, p.h, g.dim);
    if (g.dim[0] < 1 || g.dim[1] < 1 || g.dim[2] < 1) return fail(PBF_ERR_INVALID, "empty box");
    // planes this handle stores: the whole grid, or in slab mode ghost | owned | ghost
    g.xoff = 0;
    g.nxl = g.dim[0];
    g.flags = nullptr;
    s->ghost_left = 0;
    if (s->slab_on) {
        const pbf_slab_step& sl = s->slab;
        int lo = sl.has_left ? sl.x_begin - sl.ghost : 0;
        int hi = sl.has_right ? sl.x_end + sl.ghost : g.dim[0];
        if (lo < 0) lo = 0;
        if (hi > g.dim[0]) hi = g.dim[0];
        if (sl.x_begin < lo || sl.x_end > hi || sl.x_begin >= sl.x_end)
            return fail(PBF_ERR_INVALID, "slab [%d, %d) does not fit the grid (%d planes)", sl.x_begin, sl.x_end, g.dim[0]);
        g.xoff = lo;
        g.nxl = hi - lo;
        g.flags = s->flags_dev;
        s->ghost_left = sl.x_begin - lo;
    }
    int64_t ncell = (int64_t)g.nxl * g.dim[1] * g.dim[2];
    g.morton = nullptr;
    if (s->mode.morton) {
        // bit-interleaved keys: axis a contributes ceil(log2 dim[a]) bits, dealt round-robin from the lowest bit in the
        // order z, y, x (z fastest, like the reference's key); the key space — and the cell table — is the power of two
        if (s->slab_on) return fail(PBF_ERR_INVALID, "PBF_OPT_MORTON: single-GPU steps only (x-planes are not contiguous slot ranges in Morton order)");
        int bits[3], total = 0;
        for (int a = 0; a < 3; a++) { bits[a] = key_bits(g.dim[a]); if (g.dim[a] <= 1) bits[a] = 0; total += bits[a]; }
        if (bits[0] > 10 || bits[1] > 10 || bits[2] > 10 || total > 29) return fail(PBF_ERR_CAPACITY, "PBF_OPT_MORTON: grid %d x %d x %d needs more than 10 bits per axis", g.dim[0], g.dim[1], g.dim[2]);
        ncell = (int64_t)1 << total;
        if (memcmp(s->morton_dims, g.dim, sizeof(g.dim)) != 0 || !s->morton_dev) {
            static uint32_t table[3 * 1024];
            int place[3][10];
            int pos = 0;
            for (int b = 0; b < 10; b++)
                for (int a = 2; a >= 0; a--)
                    if (b < bits[a]) place[a][b] = pos++;
            for (int a = 0; a < 3; a++)
                for (int v = 0; v < 1024; v++) {
                    uint32_t k = 0;
                    for (int b = 0; b < bits[a]; b++) k |= (uint32_t)((v >> b) & 1) << place[a][b];
                    table[a * 1024 + v] = k;
                }
            if (cudaSetDevice(s->device) != cudaSuccess) return fail(PBF_ERR_CUDA, "cudaSetDevice");
            if (!s->morton_dev && cudaMalloc((void**)&s->morton_dev, sizeof(table)) != cudaSuccess) return fail(PBF_ERR_CUDA, "PBF_OPT_MORTON: no memory for the spread tables");
            if (cudaMemcpy(s->morton_dev, table, sizeof(table), cudaMemcpyHostToDevice) != cudaSuccess) return fail(PBF_ERR_CUDA, "PBF_OPT_MORTON: table upload failed");
            memcpy(s->morton_dims, g.dim, sizeof(g.dim));
        }
        s->morton_bits = total;
        g.morton = s->morton_dev;
    }
    if (ncell > s->cell_capacity || ncell >= ((int64_t)1 << 30) - 1)
        return fail(PBF_ERR_CAPACITY, "box has %lld cells, handle holds %lld", (long long)ncell, (long long)s->cell_capacity);
    g.ncell = (int32_t)ncell;
    g.dyz = g.dim[1] * g.dim[2];
    // slab mode sorts one more key value: the discard key == ncell
    s->npass = (key_bits(ncell + (s->slab_on ? 1 : 0)) + RADIX_BITS - 1) / RADIX_BITS;

    // getPoly6 / getSpikyGrad constructors and h_updateVelocity, host arithmetic as in the
    // reference (Simulator.cu:77-83, 94-98, 129); M_PI there is the double literal 3.14159265359.
    const double ref_pi = 3.14159265359;
    const float h = p.h;
    c.h = h;
    c.h2 = h * h;
    c.h2_cull = c.h2 * 1.000001f;
    const float ih = 1.f / h;
    const float ih3 = ih * ih * ih;
    const float ih9 = ih3 * ih3 * ih3;
    c.poly6_coef = (float)((double)(315.f * ih9) / ((double)64.f * ref_pi));
    float h6 = h * h;
    h6 = h6 * h6 * h6;
    c.spiky_coef = (float)((double)-45.f / (ref_pi * (double)h6));
    c.pho0 = p.pho0;
    c.lambda_eps = p.lambda_eps;
    c.k_boundary = p.k_boundaryDensity;
    // m_coef_corr = -k_corr / powf(poly6(dq*dq), n_corr), host powf (Simulator.cu:235)
    {
        const float r2 = p.delta_q * p.delta_q;
        float w = 0.f;
        if (!(r2 >= c.h2)) {
            const float d = c.h2 - r2;
            w = c.poly6_coef * d * d * d;
        }
        c.coef_corr = -p.k_corr / powf(w, p.n_corr);
    }
    c.n_corr = p.n_corr;
    c.c_xsph = p.c_XSPH;
    c.dt = p.dt;
    c.inv_dt = 1.f / p.dt;
    c.gravity = p.g;
    for (int a = 0; a < 3; a++) {
        c.lim_hi[a] = (double)s->ulim[a] - 1e-3;  // LIM_EPS, helper.h:5
        c.lim_lo[a] = (double)s->llim[a] + 1e-3;
    }
    c.exact_pow = s->exact_pow;
    // division by pho0 as a verified reciprocal sequence (pbf_math.cuh div_pho0); re-verified on the
    // device whenever pho0 changes, plain division if anything is off
    if (!(s->div_verified && s->div_d == p.pho0)) {
        s->div_rcp = (float)(1.0 / (double)p.pho0);
        verified_div_interval(s, p.pho0, s->div_rcp, &s->div_lo, &s->div_hi);
        s->div_d = p.pho0;
        s->div_verified = true;
    }
    // ... and the division by h of the cell coordinate (pbf_math.cuh cell_coord), whenever h changes
    if (!(s->hdiv_verified && s->hdiv_d == p.h)) {
        s->hdiv_rcp = (float)(1.0 / (double)p.h);
        verified_div_interval(s, p.h, s->hdiv_rcp, &s->hdiv_lo, &s->hdiv_hi);
        s->hdiv_d = p.h;
        s->hdiv_verified = true;
    }
    g.h_rcp = s->hdiv_rcp;
    g.hdiv_lo = s->hdiv_lo;
    g.hdiv_hi = s->hdiv_hi;
    c.pho0_rcp = s->div_rcp;
    c.div_lo = s->div_lo;
    c.div_hi = s->div_hi;
    // the same idea for the spiky scale: every float r2 the sweeps can hand to it is checked on the device
    // whenever h changes; PBF_NO_FAST_SPIKY=1 keeps the sqrt.rn / div.rn sequence
    if (!(s->spiky_checked && s->spiky_h == p.h)) {
        s->spiky_ok = 0;
        s->spiky_mismatches = 0;
        const char* off = getenv("PBF_NO_FAST_SPIKY");
        if (!(off && off[0] == '1') && cudaSetDevice(s->device) == cudaSuccess) {
            unsigned long long bad = ~0ull;
            bool known = false;
            for (int k = 0; k < (s->spiky_n < 4 ? s->spiky_n : 4); k++)
                if (s->spiky_cache[k].h == p.h) { bad = s->spiky_cache[k].bad; known = true; }
            if (known || verify_spiky(c, c.h2_cull, &bad, s->verify_scratch, s->verify_stream) == cudaSuccess) {
                if (!known) s->spiky_cache[s->spiky_n++ % 4] = {p.h, bad == 0 ? 1 : 0, bad};
                s->spiky_mismatches = bad;
                s->spiky_ok = bad == 0 ? 1 : 0;
            } else {
                cudaGetLastError();
            }
        }
        s->spiky_h = p.h;
        s->spiky_checked = true;
    }
    c.fast_spiky = s->spiky_ok;
    // and for the delta-p pass's powf(w, 4.0f): every float w in [0, poly6(0)]; PBF_NO_TRIM_POW=1 keeps powf
    {
        const float w_top = ((c.poly6_coef * c.h2) * c.h2) * c.h2;   // poly6_in(0), the largest w (pbf_math.cuh)
        if (!(s->pow4_checked && s->pow4_top == w_top)) {
            s->pow4_ok = 0;
            s->pow4_mismatches = 0;
            const char* off = getenv("PBF_NO_TRIM_POW");
            if (!(off && off[0] == '1') && w_top >= 0.f && w_top < 3.0e38f && cudaSetDevice(s->device) == cudaSuccess) {
                unsigned long long bad = ~0ull;
                bool known = false;
                for (int k = 0; k < (s->pow_n < 4 ? s->pow_n : 4); k++)
                    if (s->pow_cache[k].top == w_top) { bad = s->pow_cache[k].bad; known = true; }
                if (known || verify_pow4(w_top, &bad, s->verify_scratch, s->verify_stream) == cudaSuccess) {
                    if (!known) s->pow_cache[s->pow_n++ % 4] = {w_top, bad == 0 ? 1 : 0, bad};
                    s->pow4_mismatches = bad;
                    s->pow4_ok = bad == 0 ? 1 : 0;
                } else {
                    cudaGetLastError();
                }
            }
            s->pow4_top = w_top;
            s->pow4_checked = true;
        }
        c.trim_pow = s->pow4_ok;
    }
    return PBF_OK;
}

void free_all(pbf_sim* s) {
    for (auto& gq : s->graphs)
        if (gq.exec) { cudaGraphExecDestroy(gq.exec); gq.exec = nullptr; }
    cudaFree(s->keys); cudaFree(s->sort_zero); cudaFree(s->pairs[0]); cudaFree(s->pairs[1]);
    cudaFree(s->x[0]); cudaFree(s->x[1]); cudaFree(s->xl); cudaFree(s->rho); cudaFree(s->iid_sorted);
    cudaFree(s->pairs_list.js); cudaFree(s->pairs_list.cnt);
    for (int k = 0; k < 2; k++) { cudaFree(s->cull.xs[k]); cudaFree(s->cull.ys[k]); cudaFree(s->cull.zs[k]); }
    cudaFree(s->cell_range); cudaFree(s->count_scratch); cudaFree(s->read_scratch); cudaFree(s->stats_partial);
    cudaFree(s->h_pos); cudaFree(s->h_npos); cudaFree(s->h_vel); cudaFree(s->h_nvel); cudaFree(s->h_iid);
    if (s->host_main) cudaStreamDestroy(s->host_main);
    if (s->host_copy) cudaStreamDestroy(s->host_copy);
    if (s->host_ev) cudaEventDestroy(s->host_ev);
    if (s->host_iid_ev) cudaEventDestroy(s->host_iid_ev);
    for (auto& e : s->layout_ev) if (e) cudaEventDestroy(e);
    if (s->stats_host) cudaFreeHost(s->stats_host);
    cudaFree(s->plane_dev);
    if (s->plane_host) cudaFreeHost(s->plane_host);
    if (s->flags_host) cudaFreeHost(s->flags_host);
    for (auto& pr : s->peer)
        for (void* b : pr.ipc_base)
            if (b) cudaIpcCloseMemHandle(b);
    cudaFree(s->sync_words);
    cudaFree(s->halo_done);
    cudaFree(s->morton_dev);
    cudaFree(s->verify_scratch);
    if (s->verify_stream) cudaStreamDestroy(s->verify_stream);
    if (s->ev_valid) {
        for (auto& e : s->ev) cudaEventDestroy(e);
        for (auto& e : s->kev) cudaEventDestroy(e);
    }
}

}  // namespace

// ---- state files (checkpoint / resume) -------------------------------------------------------------

namespace {

struct StateHeader {            // 128 bytes on disk, little endian
    char magic[8];              // "PBFSTAT1"
    uint32_t version;           // 1
    uint32_t header_bytes;      // 128
    int64_t n;
    int64_t frame;
    pbf_params params;          // 44 bytes
    float ulim[3];
    float llim[3];
    int32_t exact_pow;
    uint64_t checksum;          // FNV-1a-64 over the payload, 8 bytes at a time (+ byte-wise tail)
    uint8_t pad[128 - 8 - 4 - 4 - 8 - 8 - 44 - 12 - 12 - 4 - 8];
} __attribute__((packed));
static_assert(sizeof(StateHeader) == 128, "state header is 128 bytes");

uint64_t fnv1a64(uint64_t h, const void* data, size_t bytes) {
    const uint8_t* p = (const uint8_t*)data;
    const uint64_t prime = 1099511628211ull;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = (h ^ w) * prime;
    }
    for (; i < bytes; i++) h = (h ^ p[i]) * prime;
    return h;
}
uint64_t payload_checksum(const float* pos, const float* vel, const uint32_t* iid, int64_t n) {
    uint64_t h = 14695981039346656037ull;
    h = fnv1a64(h, pos, (size_t)n * 12);
    h = fnv1a64(h, vel, (size_t)n * 12);
    return fnv1a64(h, iid, (size_t)n * 4);
}
int read_header(FILE* f, const char* path, StateHeader* hd) {
    if (fread(hd, 1, sizeof(*hd), f) != sizeof(*hd)) return fail(PBF_ERR_INVALID, "%s: shorter than a state header", path);
    if (memcmp(hd->magic, "PBFSTAT1", 8) != 0) return fail(PBF_ERR_INVALID, "%s: not a pbf state file (bad magic)", path);
    if (hd->version != 1 || hd->header_bytes != sizeof(*hd)) return fail(PBF_ERR_INVALID, "%s: unsupported state version %u", path, hd->version);
    if (hd->n < 0 || hd->n >= ((int64_t)1 << 30)) return fail(PBF_ERR_INVALID, "%s: implausible particle count %lld", path, (long long)hd->n);
    return PBF_OK;
}
void info_of(const StateHeader& hd, pbf_state_info* out) {
    memset(out, 0, sizeof(*out));
    out->n = hd.n;
    out->frame = hd.frame;
    out->params = hd.params;
    memcpy(out->ulim, hd.ulim, sizeof(out->ulim));
    memcpy(out->llim, hd.llim, sizeof(out->llim));
    out->exact_pow = hd.exact_pow;
    out->checksum = hd.checksum;
}

}  // namespace


extern "C" {

static int peers_signal(pbf_sim* s);

const char* pbf_last_error(void) { return g_err; }
const char* pbf_version(void) { return "pbf-cuda_b200 0.1 (sm_100a)"; }

int pbf_default_params(pbf_params* p) {
    if (!p) return fail(PBF_ERR_INVALID, "null params");
    // FluidSystem.cpp:15-25
    p->g = 9.8f;
    p->h = .1f;
    p->dt = 0.0083f;
    p->pho0 = 8000.f;
    p->lambda_eps = 1000.f;
    p->delta_q = (float)(0.3 * (double)p->h);
    p->k_corr = 0.001f;
    p->n_corr = 4;
    p->k_boundaryDensity = 0.f;
    p->c_XSPH = 0.5f;
    p->niter = 4;
    return PBF_OK;
}

int pbf_create(const pbf_params* params, const float ulim[3], const float llim[3], int64_t max_particles,
               int device, pbf_sim** out) {
    if (!params || !ulim || !llim || !out) return fail(PBF_ERR_INVALID, "null argument");
    if (max_particles <= 0 || max_particles >= ((int64_t)1 << 30)) return fail(PBF_ERR_INVALID, "max_particles out of range");
    *out = nullptr;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PBF_ERR_CUDA, "no CUDA device %d (have %d)", device, ndev);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(PBF_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(device));

    {   // load all kernels now, not at their first launch (preload_solver, solver.cu): once per device, and safe
        // when several host threads create their handles at the same time (one thread per rank); a device beyond
        // the table is preloaded on every create (the query is cheap once the module is resident)
        auto preload = []() -> cudaError_t {
            cudaFuncAttributes fa;
            cudaError_t e = cudaFuncGetAttributes(&fa, extract_words_kernel);
            if (e == cudaSuccess) e = preload_advect_key();
            if (e == cudaSuccess) e = preload_sort();
            if (e == cudaSuccess) e = preload_reorder();
            if (e == cudaSuccess) e = preload_scene();
            if (e == cudaSuccess) e = preload_slab();
            if (e == cudaSuccess) e = preload_solver();
            if (e == cudaSuccess) e = preload_stats();
            return e;
        };
        static std::once_flag once[64];
        static cudaError_t result[64];
        cudaError_t pe;
        if (device < 64) {
            std::call_once(once[device], [&]() { result[device] = preload(); });
            pe = result[device];
        } else {
            pe = preload();
        }
        if (pe != cudaSuccess) return fail(PBF_ERR_CUDA, "kernel preload failed: %s", cudaGetErrorString(pe));
    }
    pbf_sim* s = new (std::nothrow) pbf_sim;
    if (!s) return fail(PBF_ERR_INVALID, "out of host memory");
    s->device = device;
    s->p = *params;
    memcpy(s->ulim, ulim, sizeof(float) * 3);
    memcpy(s->llim, llim, sizeof(float) * 3);
    s->max_particles = max_particles;
    // cell capacity: the reference's 4*(int)(dx*dy*dz) with d = (ulim-llim)/0.1 (Simulator.h:13-14),
    // but never less than twice the cells of this box at the actual h (the sweep widens it 1.5x).
    int32_t dim[3];
    compute_dim(ulim, llim, params->h, dim);
    const int64_t cells = (int64_t)dim[0] * dim[1] * dim[2];
    const double dx = (ulim[0] - llim[0]) / 0.1, dy = (ulim[1] - llim[1]) / 0.1, dz = (ulim[2] - llim[2]) / 0.1;
    int64_t cap = 4 * (int64_t)(dx * dy * dz);
    if (cap < 2 * cells) cap = 2 * cells;
    if (cap < 1) cap = 1;
    if (cap >= ((int64_t)1 << 30)) cap = ((int64_t)1 << 30) - 1;
    s->cell_capacity = cap;
    // default: powf like the reference (bit-identical results); PBF_FAST_POW=1 opts into (w*w)^2
    const char* ep = getenv("PBF_FAST_POW");
    s->exact_pow = (ep && ep[0] == '1') ? 0 : 1;
    // (environment variables only set the defaults of the handle's options, here; nothing reads them per launch)
    if (const char* tm = getenv("PBF_TEAM")) s->mode.team = tm[0] == '1' ? 1 : tm[0] == '0' ? 0 : -1;
    if (const char* rb = getenv("PBF_REBIN")) s->mode.rebin = rb[0] == '1' ? 1 : 0;
    if (const char* sg = getenv("PBF_STAGED")) s->mode.staged = sg[0] == '1' ? 1 : 0;
    if (const char* pr = getenv("PBF_PAIRED")) s->mode.paired = pr[0] == '1' ? 1 : 0;
    if (const char* mo = getenv("PBF_MORTON")) s->mode.morton = mo[0] == '1' ? 1 : 0;
    if (const char* co = getenv("PBF_COOP")) s->mode.coop = co[0] == '1' ? 1 : 0;
    if (const char* pd = getenv("PBF_PDL")) s->mode.pdl = pd[0] == '0' ? 0 : 1;
    if (const char* hk = getenv("PBF_HALO_INKERNEL")) s->mode.halo_inkernel = hk[0] == '0' ? 0 : 1;
    if (const char* gr = getenv("PBF_GRAPH")) s->mode.graph = gr[0] == '1' ? 1 : gr[0] == '0' ? 0 : -1;

    const size_t n = (size_t)max_particles;
    s->sort_zero_capacity = sort_scratch_capacity_bytes(max_particles);
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void**)&s->keys, n * 4);
    A((void**)&s->sort_zero, s->sort_zero_capacity + 16);
    if (e == cudaSuccess) e = cudaMemset(s->sort_zero, 0, s->sort_zero_capacity + 16);
    A((void**)&s->pairs[0], n * sizeof(KeyIdx));
    A((void**)&s->pairs[1], n * sizeof(KeyIdx));
    // (+8: the cull of the neighbour sweeps reads runs in groups of four, up to 3 slots past their end)
    A((void**)&s->x[0], (n + 8) * sizeof(float4));
    A((void**)&s->x[1], (n + 8) * sizeof(float4));
    A((void**)&s->xl, (n + 8) * sizeof(float4));
    // (the cull reads whole groups of four slots, the hits of the slots outside a run are masked off: the
    //  slots past the last particle are never written, so give them a defined value once)
    for (int k = 0; k < 2; k++) {
        A((void**)&s->cull.xs[k], (n + 8) * 4);
        A((void**)&s->cull.ys[k], (n + 8) * 4);
        A((void**)&s->cull.zs[k], (n + 8) * 4);
        if (e == cudaSuccess) e = cudaMemset(s->cull.xs[k], 0, (n + 8) * 4);
        if (e == cudaSuccess) e = cudaMemset(s->cull.ys[k], 0, (n + 8) * 4);
        if (e == cudaSuccess) e = cudaMemset(s->cull.zs[k], 0, (n + 8) * 4);
    }
    A((void**)&s->rho, n * 4);
    A((void**)&s->iid_sorted, n * 4);
    A((void**)&s->cell_range, (size_t)(cap + 1) * sizeof(uint2));   // (+1: emptied two cells per store)
    A((void**)&s->stats_partial, 1024 * 5 * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&s->stats_host, 1024 * 5 * sizeof(double));
    // slab mode: plane table (every plane the box can have, x2 for a moving wall) + flag word
    s->plane_capacity = 2 * (int64_t)dim[0] + 8;
    A((void**)&s->plane_dev, (size_t)s->plane_capacity * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&s->plane_host, (size_t)s->plane_capacity * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->flags_host, sizeof(uint32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) { *s->flags_host = 0; e = cudaHostGetDevicePointer((void**)&s->flags_dev, s->flags_host, 0); }
    A((void**)&s->sync_words, 8 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(s->sync_words, 0, 8 * sizeof(uint32_t));
    A((void**)&s->verify_scratch, 16);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->verify_stream, cudaStreamNonBlocking);
    A((void**)&s->halo_done, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(s->halo_done, 0, 2 * sizeof(uint32_t));
    // (the memset runs on the legacy stream; a neighbour's first flag store comes from a non-blocking
    //  stream and must not be overtaken by it)
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (const char* to = getenv("PBF_HALO_TIMEOUT_MS")) s->halo_timeout_ns = (uint64_t)atoll(to) * 1000000ull;
    if (e != cudaSuccess) {
        free_all(s);
        delete s;
        return fail(PBF_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(e));
    }
    // Neighbour-list reuse between the lambda and delta-p passes: ~0.77 KB per particle of scratch.
    // Taken only if it fits comfortably (<= 40 % of the free memory); PBF_NO_PAIR_REUSE=1 disables it.
    {
        const char* np = getenv("PBF_NO_PAIR_REUSE");
        size_t jb, cb, free_b = 0, total_b = 0;
        const size_t need = pair_list_bytes(max_particles, &jb, &cb);
        cudaMemGetInfo(&free_b, &total_b);
        if (!(np && np[0] == '1') && need <= free_b / 10 * 4) {
            cudaError_t pe = cudaMalloc((void**)&s->pairs_list.js, jb);
            if (pe == cudaSuccess) pe = cudaMalloc((void**)&s->pairs_list.cnt, cb);
            if (pe != cudaSuccess) {
                cudaFree(s->pairs_list.js); cudaFree(s->pairs_list.cnt);
                s->pairs_list = PairList();
                cudaGetLastError();
            }
        }
    }
    int rc = refresh_consts(s);
    if (rc != PBF_OK) { free_all(s); delete s; return rc; }
    *out = s;
    return PBF_OK;
}

int pbf_destroy(pbf_sim* s) {
    if (!s) return PBF_OK;
    cudaSetDevice(s->device);
    free_all(s);
    delete s;
    return PBF_OK;
}

int pbf_set_params(pbf_sim* s, const pbf_params* p) {
    if (!s || !p) return fail(PBF_ERR_INVALID, "null argument");
    pbf_params old = s->p;
    s->p = *p;
    int rc = refresh_consts(s);
    if (rc != PBF_OK) { s->p = old; refresh_consts(s); }
    return rc;
}
int pbf_get_params(const pbf_sim* s, pbf_params* out) {
    if (!s || !out) return fail(PBF_ERR_INVALID, "null argument");
    *out = s->p;
    return PBF_OK;
}
int pbf_set_option(pbf_sim* s, int option, int value) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    switch (option) {
        case PBF_OPT_TEAM:
            if (value < -1 || value > 1) return fail(PBF_ERR_INVALID, "PBF_OPT_TEAM takes -1, 0 or 1");
            s->mode.team = value;
            return PBF_OK;
        case PBF_OPT_REBIN:
            s->mode.rebin = value ? 1 : 0;
            return PBF_OK;
        case PBF_OPT_PDL:
            s->mode.pdl = value ? 1 : 0;
            return PBF_OK;
        case PBF_OPT_STAGED:
            s->mode.staged = value ? 1 : 0;
            return PBF_OK;
        case PBF_OPT_PAIRED:
            s->mode.paired = value ? 1 : 0;
            return PBF_OK;
        case PBF_OPT_COOP:
            s->mode.coop = value ? 1 : 0;
            return PBF_OK;
        case PBF_OPT_MORTON: {
            const int old = s->mode.morton;
            s->mode.morton = value ? 1 : 0;
            int rc = refresh_consts(s);
            if (rc != PBF_OK) { s->mode.morton = old; refresh_consts(s); }
            return rc;
        }
        case PBF_OPT_HALO_INKERNEL:
            s->mode.halo_inkernel = value ? 1 : 0;
            return PBF_OK;
        case PBF_OPT_GRAPH:
            if (value < -1 || value > 1) return fail(PBF_ERR_INVALID, "PBF_OPT_GRAPH takes -1, 0 or 1");
            s->mode.graph = value;
            s->graph_holdoff = s->graph_misses = 0;
            return PBF_OK;
        default: return fail(PBF_ERR_INVALID, "unknown option %d", option);
    }
}
int pbf_get_option(const pbf_sim* s, int option, int* value) {
    if (!s || !value) return fail(PBF_ERR_INVALID, "null argument");
    switch (option) {
        case PBF_OPT_TEAM: *value = s->mode.team; return PBF_OK;
        case PBF_OPT_REBIN: *value = s->mode.rebin; return PBF_OK;
        case PBF_OPT_PDL: *value = s->mode.pdl; return PBF_OK;
        case PBF_OPT_STAGED: *value = s->mode.staged; return PBF_OK;
        case PBF_OPT_PAIRED: *value = s->mode.paired; return PBF_OK;
        case PBF_OPT_MORTON: *value = s->mode.morton; return PBF_OK;
        case PBF_OPT_COOP: *value = s->mode.coop; return PBF_OK;
        case PBF_OPT_GRAPH: *value = s->mode.graph; return PBF_OK;
        case PBF_OPT_HALO_INKERNEL: *value = s->mode.halo_inkernel; return PBF_OK;
        default: return fail(PBF_ERR_INVALID, "unknown option %d", option);
    }
}
int pbf_set_option_exact_pow(pbf_sim* s, int on) {
    if (!s) return fail(PBF_ERR_INVALID, "null argument");
    s->exact_pow = on ? 1 : 0;
    return refresh_consts(s);
}
int pbf_set_lim(pbf_sim* s, const float ulim[3], const float llim[3]) {
    if (!s || !ulim || !llim) return fail(PBF_ERR_INVALID, "null argument");
    float ou[3], ol[3];
    memcpy(ou, s->ulim, sizeof(ou)); memcpy(ol, s->llim, sizeof(ol));
    memcpy(s->ulim, ulim, sizeof(float) * 3);
    memcpy(s->llim, llim, sizeof(float) * 3);
    int rc = refresh_consts(s);
    if (rc != PBF_OK) { memcpy(s->ulim, ou, sizeof(ou)); memcpy(s->llim, ol, sizeof(ol)); refresh_consts(s); }
    return rc;
}
int pbf_get_lim(const pbf_sim* s, float ulim[3], float llim[3]) {
    if (!s || !ulim || !llim) return fail(PBF_ERR_INVALID, "null argument");
    memcpy(ulim, s->ulim, sizeof(float) * 3);
    memcpy(llim, s->llim, sizeof(float) * 3);
    return PBF_OK;
}
int pbf_get_const_div_interval(const pbf_sim* s, float* lo, float* hi) {
    if (!s || !lo || !hi) return fail(PBF_ERR_INVALID, "null argument");
    *lo = s->div_lo; *hi = s->div_hi;
    return PBF_OK;
}
int pbf_get_trim_pow(const pbf_sim* s, int32_t* on, uint64_t* mismatches) {
    if (!s || !on || !mismatches) return fail(PBF_ERR_INVALID, "null argument");
    *on = s->pow4_ok;
    *mismatches = (uint64_t)s->pow4_mismatches;
    return PBF_OK;
}
int pbf_get_pair_list(const pbf_sim* s, int32_t* on, uint64_t* bytes) {
    if (!s || !on || !bytes) return fail(PBF_ERR_INVALID, "null argument");
    size_t jb = 0, cb = 0;
    const size_t need = pair_list_bytes(s->max_particles, &jb, &cb);
    *on = s->pairs_list.js ? 1 : 0;
    *bytes = (uint64_t)need;
    return PBF_OK;
}
int pbf_get_fast_spiky(const pbf_sim* s, int32_t* on, uint64_t* mismatches) {
    if (!s || !on || !mismatches) return fail(PBF_ERR_INVALID, "null argument");
    *on = s->spiky_ok;
    *mismatches = (uint64_t)s->spiky_mismatches;
    return PBF_OK;
}
int pbf_get_grid_dim(const pbf_sim* s, int32_t dim[3]) {
    if (!s || !dim) return fail(PBF_ERR_INVALID, "null argument");
    dim[0] = s->g.dim[0]; dim[1] = s->g.dim[1]; dim[2] = s->g.dim[2];
    return PBF_OK;
}
int64_t pbf_launch_count(const pbf_sim* s) { return s ? s->launches : 0; }

/* ---- stages ------------------------------------------------------------------------------- */

// (every stage function: the launchers read the thread-local PDL switch, and stages may be called from any thread)
#define STAGE_ENTER(s) do { if (s) tl_pdl = (s)->mode.pdl != 0; } while (0)

static int stage_event(pbf_sim* s, int k) {
    if (s->timing) CUDA_TRY(cudaEventRecord(s->ev[k], s->stream));
    return PBF_OK;
}
// brackets one kernel slot (PBF_KERNEL_*) with events when timing is on
static int kernel_event(pbf_sim* s, int slot, int after) {
    if (s->timing) CUDA_TRY(cudaEventRecord(s->kev[2 * slot + after], s->stream));
    return PBF_OK;
}
#define KTIMED(slot, call)                                   \
    do {                                                     \
        int rc__ = kernel_event(s, slot, 0);                 \
        if (rc__) return rc__;                               \
        CUDA_TRY(call);                                      \
        if ((rc__ = kernel_event(s, slot, 1))) return rc__;  \
    } while (0)

// Slab mode, after the sort: where every stored plane starts in the sorted order. One small
// kernel + a 8*(nxl+1)-byte download + the step's only host synchronisation; from the table
// follow the owned range, the ghost counts and the sizes of every halo message.
// ... in two halves, so that the reorder pass (which reads the table on the device) can run while the host waits for
// the download: begin = plane table, download on the handle's side stream, publish; end = handshake wait on the
// compute stream, host waits for the DOWNLOAD only (an event on the side stream), layout.
static int slab_layout_begin(pbf_sim* s) {
    const GridConsts& g = s->g;
    if (g.nxl + 1 > s->plane_capacity) return fail(PBF_ERR_CAPACITY, "slab stores %d planes, handle holds %lld", g.nxl, (long long)s->plane_capacity - 1);
    if (!s->layout_ev[0]) {
        CUDA_TRY(cudaEventCreateWithFlags(&s->layout_ev[0], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s->layout_ev[1], cudaEventDisableTiming));
    }
    CUDA_TRY(launch_plane_table(s->pairs[s->sorted_buf], s->n, s->plane_dev, g, s->stream, &s->launches));
    CUDA_TRY(cudaEventRecord(s->layout_ev[0], s->stream));
    CUDA_TRY(cudaStreamWaitEvent(s->verify_stream, s->layout_ev[0], 0));
    CUDA_TRY(cudaMemcpyAsync(s->plane_host, s->plane_dev, (size_t)(g.nxl + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s->verify_stream));
    CUDA_TRY(cudaEventRecord(s->layout_ev[1], s->verify_stream));
    const pbf_slab_step& sl = s->slab;
    if (s->peer[0].on || s->peer[1].on) {
        // fused halo: device to device, tell the right neighbour where my right-ghost slots begin (the
        // first slot behind my owned planes) — its lambda pass pushes there — and exchange a handshake so
        // that nobody pushes before the word has arrived. No host is involved.
        s->halo_seq++;
        if (getenv("PBF_HALO_TRACE")) fprintf(stderr, "[halo %p] publish+wait %u\n", (void*)s, s->halo_seq);
        CUDA_TRY(launch_halo_publish(s->plane_dev + s->ghost_left + (sl.x_end - sl.x_begin),
                                     s->peer[1].on ? (int64_t*)(s->peer[1].sync + 2) : nullptr,
                                     s->peer[0].on ? s->peer[0].sync + 1 : nullptr, s->peer[1].on ? s->peer[1].sync + 0 : nullptr,
                                     s->halo_seq, s->stream, &s->launches));
    }
    return PBF_OK;
}
static int slab_layout_end(pbf_sim* s) {
    const GridConsts& g = s->g;
    const pbf_slab_step& sl = s->slab;
    if (s->peer[0].on || s->peer[1].on)
        CUDA_TRY(launch_halo_wait(s->peer[0].on ? s->sync_words + 0 : nullptr, s->peer[1].on ? s->sync_words + 1 : nullptr,
                                  s->halo_seq, s->halo_timeout_ns, s->flags_dev, s->stream, &s->launches));
    CUDA_TRY(cudaEventSynchronize(s->layout_ev[1]));   // the step's one host wait: for 8 * (planes + 1) bytes
    const int64_t* ps = s->plane_host;
    const int gl = s->ghost_left, nx = sl.x_end - sl.x_begin;
    const int gw_l = sl.has_left ? (sl.ghost < nx ? sl.ghost : nx) : 0;   // owned planes a neighbour mirrors
    const int gw_r = sl.has_right ? (sl.ghost < nx ? sl.ghost : nx) : 0;
    pbf_slab_layout& L = s->layout;
    L.n_local = ps[g.nxl];
    L.own_first = ps[gl];
    L.own_count = ps[gl + nx] - ps[gl];
    L.recv_left_count = ps[gl];
    L.recv_right_count = ps[g.nxl] - ps[gl + nx];
    L.send_left_count = ps[gl + gw_l] - ps[gl];
    L.send_right_count = ps[gl + nx] - ps[gl + nx - gw_r];
    L.flags = *s->flags_host;
    // the edges of the in-kernel handshake: every owned particle whose neighbour search can reach a ghost plane.
    // A particle drifts at most one cell per Jacobi iteration from the cell it was sorted into (MAX_DP, helper.h:9)
    // and searches one cell further: niter + 1 planes — and at least the planes the neighbours mirror.
    {
        int w = s->p.niter + 1 > sl.ghost ? s->p.niter + 1 : sl.ghost;
        if (w > nx) w = nx;
        s->edge_left = sl.has_left ? ps[gl + w] - ps[gl] : 0;
        s->edge_right_first = sl.has_right ? ps[gl + nx - w] - ps[gl] : L.own_count;
    }
    s->n_local = L.n_local;
    s->own_first = L.own_first;
    s->own_count = L.own_count;
    s->layout_valid = true;
    return PBF_OK;
}
static int slab_learn_layout(pbf_sim* s) {
    int rc = slab_layout_begin(s);
    return rc ? rc : slab_layout_end(s);
}

// Where the boundary values of the pass writing array `a` (x[0], x[1] or xl of this handle) go on
// the attached neighbours. Empty when no neighbour is attached: the caller's transport does it.
static int make_push(pbf_sim* s, const float4* a, HaloPush* hp) {
    *hp = HaloPush();
    if (!s->slab_on || (!s->peer[0].on && !s->peer[1].on)) return PBF_OK;
    if (!s->layout_valid) return fail(PBF_ERR_STATE, "fused halo: no slab layout");
    const pbf_slab_layout& L = s->layout;
    for (int side = 0; side < 2; side++) {
        const pbf_sim::Peer& pr = s->peer[side];
        if (!pr.on) continue;
        float4* dst = a == s->x[0] ? pr.x[0] : a == s->x[1] ? pr.x[1] : pr.xl;
        if (side == 0 && L.send_left_count > 0) {
            hp->left = dst;
            hp->left_tail = (const int64_t*)(s->sync_words + 2);
            hp->left_count = L.send_left_count;
        }
        if (side == 1 && L.send_right_count > 0) { hp->right = dst; hp->right_first = L.own_count - L.send_right_count; }
    }
    return PBF_OK;
}

// The in-kernel handshake of one pass (HaloSync): `consumer` = its edge blocks read ghost slots the neighbours
// filled in the last pushing pass; `producer` = it pushes boundary values itself, under a fresh handshake number.
// Empty unless the neighbours are attached and PBF_OPT_HALO_INKERNEL is on.
static bool inkernel_halo(const pbf_sim* s) {
    return s->slab_on && s->mode.halo_inkernel && (s->peer[0].on || s->peer[1].on);
}
static void make_sync(pbf_sim* s, bool consumer, bool producer, HaloSync* hs) {
    *hs = HaloSync();
    if (!inkernel_halo(s) || !s->layout_valid) return;
    const pbf_slab_layout& L = s->layout;
    hs->edge_left = s->peer[0].on ? s->edge_left : 0;
    hs->edge_right_first = s->peer[1].on ? s->edge_right_first : INT64_MAX;
    (void)L;
    hs->done = s->halo_done;
    hs->timeout_ns = s->halo_timeout_ns;
    hs->flags = s->flags_dev;
    if (consumer && s->wait_valid) {
        hs->wait_left = s->peer[0].on ? s->sync_words + 0 : nullptr;
        hs->wait_right = s->peer[1].on ? s->sync_words + 1 : nullptr;
        hs->wait_seq = s->wait_seq;
    }
    if (producer) {
        s->halo_seq++;
        // I am my left neighbour's RIGHT neighbour: its word [1]; and my right neighbour's word [0]
        hs->peer_left = s->peer[0].on ? s->peer[0].sync + 1 : nullptr;
        hs->peer_right = s->peer[1].on ? s->peer[1].sync + 0 : nullptr;
        hs->signal_seq = s->halo_seq;
        s->wait_seq = s->halo_seq;
        s->wait_valid = true;
    }
}
// A side whose edge is empty has no block that could raise the neighbour's word: a one-thread kernel does it.
static int signal_empty_edges(pbf_sim* s, const HaloSync& hs) {
    if (!hs.peer_left && !hs.peer_right) return PBF_OK;
    const bool none = s->own_count <= 0;
    uint32_t* l = hs.peer_left && (none || hs.edge_left <= 0) ? hs.peer_left : nullptr;
    uint32_t* r = hs.peer_right && (none || hs.edge_right_first >= s->own_count) ? hs.peer_right : nullptr;
    if (l || r) CUDA_TRY(launch_halo_signal(l, r, hs.signal_seq, s->stream, &s->launches));
    return PBF_OK;
}
// the ghost slots' coordinates for the cull (the owned slots' came with the pass that produced `x`)
static int pack_ghosts_if_needed(pbf_sim* s, const float4* x) {
    if (!inkernel_halo(s) || s->cull.holds == x || !s->wait_valid) return PBF_OK;
    HaloSync hs;
    make_sync(s, true, false, &hs);
    CUDA_TRY(launch_pack_ghosts(x, s->cull, s->n_local, s->own_first, s->own_count, hs, s->stream, &s->launches));
    return PBF_OK;
}

int pbf_stage_begin(pbf_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n,
                    void* stream) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    if (n < 0 || n > s->max_particles) return fail(PBF_ERR_CAPACITY, "n=%lld exceeds max_particles=%lld", (long long)n, (long long)s->max_particles);
    if (n > 0 && (!pos || !npos || !vel || !nvel || !iid)) return fail(PBF_ERR_INVALID, "null particle buffer");
    CUDA_TRY(cudaSetDevice(s->device));
    tl_pdl = s->mode.pdl != 0;
    s->pos = pos; s->npos = npos; s->vel = vel; s->nvel = nvel; s->iid = iid; s->n = n;
    s->stream = (cudaStream_t)stream;
    s->reorder_wait = nullptr;
    if (s->slab_on) {  // a plain step on a handle that ran slab steps before: back to the whole grid
        s->slab_on = false;
        int rc = refresh_consts(s);
        if (rc) return rc;
    }
    s->si = SlabInput{};
    s->si.n_own = n;
    s->n_local = n; s->own_first = 0; s->own_count = n;
    s->layout_valid = false;
    s->wait_valid = false;
    s->state_push = StatePush();
    s->cur = 0;
    s->iters_done = 0;
    s->pos0_in_npos = false;
    s->stage = ST_BOUND;
    return PBF_OK;
}

int pbf_stage_advect(pbf_sim* s) {
    STAGE_ENTER(s);
    if (!s || s->stage != ST_BOUND) return fail(PBF_ERR_STATE, "advect: call pbf_stage_begin first");
    int rc = stage_event(s, 0);
    if (rc) return rc;
    if (s->sort_dirty) {   // the previous step never reached its reorder pass: restore the invariant
        CUDA_TRY(cudaMemsetAsync(s->sort_zero, 0, s->sort_zero_capacity, s->stream));
        s->launches++;
    }
    s->sort_dirty = s->n > 0;
    KTIMED(PBF_KERNEL_ADVECT_KEY, launch_advect_key(s->pos, s->vel, s->keys, s->sort_zero, s->cell_range, s->n, s->npass, s->si, s->g, s->c, s->stream, &s->launches));
    s->stage = ST_ADVECTED;
    return stage_event(s, 1);
}

int pbf_stage_build_grid(pbf_sim* s) {
    STAGE_ENTER(s);
    if (!s || s->stage != ST_ADVECTED) return fail(PBF_ERR_STATE, "build_grid: advect first");
    SortScratch sc;
    sc.hist = s->sort_zero;
    sc.tile_counter = s->sort_zero + MAX_PASSES * RADIX;
    sc.tile_desc = s->sort_zero + MAX_PASSES * RADIX + MAX_PASSES;
    sc.bufs[0] = s->pairs[0];
    sc.bufs[1] = s->pairs[1];
    sc.tile_desc_words = 0;
    KTIMED(PBF_KERNEL_SORT, launch_sort(s->keys, sc, s->n, s->npass, s->si, &s->sorted_buf, s->stream, &s->launches));
    if (s->slab_on) {
        int rc = slab_layout_begin(s);
        if (rc) return rc;
    }
    if (s->reorder_wait) {   // reorder is the first kernel that reads iid
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->reorder_wait, 0));
        s->reorder_wait = nullptr;
    }
    // slab mode: launched over every sorted entry with the layout read from the device table, so that it runs while
    // the host is still waiting for that table (slab_layout_end below)
    const pbf_slab_step& sl = s->slab;
    KTIMED(PBF_KERNEL_REORDER, launch_reorder(s->pairs[s->sorted_buf], s->pos, s->vel, s->iid, s->x[0], s->cull, s->npos, s->iid_sorted,
                                              s->cell_range, s->sort_zero, sort_scratch_zero_bytes(s->n, s->npass),
                                              s->slab_on ? s->n : s->n_local, s->own_first, s->own_count,
                                              s->slab_on ? s->plane_dev : nullptr, s->ghost_left, sl.x_end - sl.x_begin,
                                              s->g, s->c, s->stream, &s->launches));
    if (s->slab_on) {
        int rc = slab_layout_end(s);
        if (rc) return rc;
    }
    s->sort_dirty = false;   // (the reorder kernel left the sort's scratch zeroed)
    s->cur = 0;
    s->pos0_in_npos = true;
    s->stage = ST_GRID;
    return stage_event(s, 2);
}

int pbf_stage_lambda(pbf_sim* s) {
    STAGE_ENTER(s);
    if (!s || (s->stage != ST_GRID && s->stage != ST_DENSITY)) return fail(PBF_ERR_STATE, "lambda: build_grid first");
    // (each iteration overwrites the slot: the timers report the LAST iteration of the step)
    HaloPush hp;
    int prc = make_push(s, s->xl, &hp);
    if (prc) return prc;
    s->mode.moved = s->iters_done > 0;   // the first iteration runs on the positions the sort keyed on
    if (s->iters_done > 0 && (prc = pack_ghosts_if_needed(s, s->x[s->cur]))) return prc;
    HaloSync hs;
    make_sync(s, false, true, &hs);
    KTIMED(PBF_KERNEL_LAMBDA, launch_lambda(s->x[s->cur], s->cull, s->n_local, s->xl, s->rho, s->cell_range, s->own_first, s->own_count, s->pairs_list, hp, hs, s->g, s->c, s->mode, s->stream, &s->launches));
    if ((prc = signal_empty_edges(s, hs))) return prc;
    s->stage = ST_LAMBDA;
    return PBF_OK;
}

int pbf_stage_delta_p(pbf_sim* s) {
    STAGE_ENTER(s);
    if (!s || s->stage != ST_LAMBDA) return fail(PBF_ERR_STATE, "delta_p: lambda first");
    HaloPush hp;
    int prc = make_push(s, s->x[s->cur ^ 1], &hp);
    if (prc) return prc;
    // pbf_step's last iteration on a single GPU: the velocity update rides along (VelTail, pbf_internal.h); the
    // velocities go to the iterate buffer this pass does not read any more (it works from xl)
    VelTail vt;
    const bool fused = s->fuse_velocity && !s->slab_on && !s->timing;
    if (fused) {
        vt.rho = s->rho; vt.pos_out = s->pos; vt.npos_io = s->npos; vt.vel_out = s->vel; vt.v4 = s->x[s->cur];
        vt.inv_dt = s->c.inv_dt;
    }
    HaloSync hs;
    make_sync(s, true, true, &hs);
    if (fused && s->tail_npos_host && s->own_count >= 4 * 131072 && delta_p_sliceable(s->pairs_list, s->mode, s->own_count)) {
        // pbf_step_host: the pass that produces the final positions runs in four slices of the sorted order, and the
        // positions of a finished slice go home on the copy stream while the next slice is computed
        const int64_t nb = (s->own_count + 127) / 128;
        for (int k = 0; k < 4; k++) {
            const int64_t b0 = nb * k / 4, b1 = nb * (k + 1) / 4;
            CUDA_TRY(launch_delta_p(s->xl, s->cull, s->n_local, s->x[s->cur ^ 1], s->cell_range, s->own_first, s->own_count, s->pairs_list, hp, hs, vt, s->g, s->c, s->mode, s->stream, &s->launches,
                                    (uint32_t)b0, (uint32_t)(b1 - b0), k == 3));
            const int64_t a = b0 * 128, b = b1 * 128 < s->own_count ? b1 * 128 : s->own_count;
            CUDA_TRY(cudaEventRecord(s->host_ev, s->stream));
            CUDA_TRY(cudaStreamWaitEvent(s->host_copy, s->host_ev, 0));
            CUDA_TRY(cudaMemcpyAsync(s->tail_npos_host + 3 * a, s->npos + 3 * a, (size_t)(b - a) * 12, cudaMemcpyDeviceToHost, s->host_copy));
        }
        s->tail_npos_sent = true;
    } else
    KTIMED(PBF_KERNEL_DELTA_P, launch_delta_p(s->xl, s->cull, s->n_local, s->x[s->cur ^ 1], s->cell_range, s->own_first, s->own_count, s->pairs_list, hp, hs, vt, s->g, s->c, s->mode, s->stream, &s->launches));
    if ((prc = signal_empty_edges(s, hs))) return prc;
    s->cur ^= 1;
    s->iters_done++;
    s->stage = ST_DENSITY;
    if (fused) {   // what pbf_stage_update_velocity leaves behind
        s->v4 = s->x[s->cur ^ 1];
        s->pos0_in_npos = false;
        s->stage = ST_VELOCITY;
    }
    return PBF_OK;
}

int pbf_stage_correct_density(pbf_sim* s) {
    if (!s || (s->stage != ST_GRID && s->stage != ST_DENSITY)) return fail(PBF_ERR_STATE, "correct_density: build_grid first");
    int rc = pbf_stage_lambda(s);
    if (rc) return rc;
    return pbf_stage_delta_p(s);
}

int pbf_stage_update_velocity(pbf_sim* s) {
    STAGE_ENTER(s);
    if (!s || (s->stage != ST_GRID && s->stage != ST_DENSITY)) return fail(PBF_ERR_STATE, "update_velocity: build_grid first");
    int rc = stage_event(s, 3);
    if (rc) return rc;
    // xl is dead after the last delta-p pass: reuse it for (velocity, rho)
    HaloPush hp;
    int prc = make_push(s, s->xl, &hp);
    if (prc) return prc;
    // (a reader too: its pushes overwrite the (x, y, z, lambda) the neighbours' last delta-p pass gathered from
    //  their ghost slots, so its edge blocks wait for that pass's completion word first)
    HaloSync hs;
    make_sync(s, true, true, &hs);
    KTIMED(PBF_KERNEL_UPDATE_VELOCITY, launch_update_velocity(s->x[s->cur], s->rho, s->pos, s->npos, s->vel, s->xl, s->own_first, s->own_count, hp, hs, s->state_push, s->c, s->stream, &s->launches));
    if ((prc = signal_empty_edges(s, hs))) return prc;
    s->v4 = s->xl;
    s->pos0_in_npos = false;
    s->stage = ST_VELOCITY;
    return stage_event(s, 4);
}

int pbf_stage_correct_velocity(pbf_sim* s) {
    STAGE_ENTER(s);
    if (!s || s->stage != ST_VELOCITY) return fail(PBF_ERR_STATE, "correct_velocity: update_velocity first");
    s->mode.moved = s->iters_done > 0;
    int prc = s->iters_done > 0 ? pack_ghosts_if_needed(s, s->x[s->cur]) : PBF_OK;
    if (prc) return prc;
    HaloSync hs;
    make_sync(s, true, false, &hs);
    KTIMED(PBF_KERNEL_XSPH, launch_xsph(s->x[s->cur], s->cull, s->n_local, s->v4, s->cell_range, s->nvel, s->iid_sorted, s->iid, s->own_first, s->own_count, hs, s->state_push, s->g, s->c, s->mode, s->stream, &s->launches));
    s->stage = ST_XSPH;
    return stage_event(s, 5);
}

int pbf_stage_end(pbf_sim* s) {
    STAGE_ENTER(s);
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    s->stage = ST_IDLE;
    // fused mode: tell the neighbours that this step's state is complete — they pull from it next step
    if (s->slab_on && s->state_iid && (s->peer[0].on || s->peer[1].on)) {
        int rc = peers_signal(s);
        if (rc) return rc;
        s->state_seq = s->halo_seq;
    }
    // A handshake that timed out, or a neighbour search that left the stored planes, lets the kernels run on with
    // stale / missing ghost values: a caller of the bare C ABI that never polls pbf_slab_flags must not get
    // that silently. The word is mapped host memory: no synchronisation, so what is seen here is what the device
    // has raised SO FAR (typically the previous step's) — the sticky flags stay up until pbf_slab_flags clears them.
    if (s->slab_on && s->flags_host) {
        const uint32_t f = *(volatile uint32_t*)s->flags_host;
        if (f & PBF_SLAB_FLAG_TIMEOUT) return fail(PBF_ERR_STATE, "slab step: a neighbour's halo completion word did not arrive in time (PBF_SLAB_FLAG_TIMEOUT); results are not valid");
        if (f & PBF_SLAB_FLAG_GHOST) return fail(PBF_ERR_STATE, "slab step: a neighbour search left the stored ghost planes (PBF_SLAB_FLAG_GHOST); results are not exact — raise `ghost`");
    }
    return PBF_OK;
}

static int step_direct(pbf_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n, void* stream) {
    int rc = pbf_stage_begin(s, pos, npos, vel, nvel, iid, n, stream);
    if (rc) return rc;
    if ((rc = pbf_stage_advect(s))) return rc;
    if ((rc = pbf_stage_build_grid(s))) return rc;
    if (s->mode.coop && !s->timing && s->p.niter >= 1 && n > 0 && sweeps_use_team(s->mode, n)) {
        // PBF_OPT_COOP: every solver pass of the step in one persistent cooperative kernel (north-star item 3)
        float4* const xx[2] = {s->x[0], s->x[1]};
        const cudaError_t ce = launch_solve_team_coop(xx, s->cull, s->xl, s->rho, s->cell_range, s->pairs_list, s->pos, s->npos, s->vel, s->nvel,
                                                      s->iid_sorted, s->iid, n, s->p.niter, s->g, s->c, s->stream, &s->launches);
        if (ce == cudaSuccess) {
            s->cur = s->p.niter & 1;
            s->iters_done = s->p.niter;
            s->v4 = s->x[s->cur ^ 1];
            s->pos0_in_npos = false;
            s->stage = ST_XSPH;
            return pbf_stage_end(s);
        }
        cudaGetLastError();
        if (ce != cudaErrorNotSupported) return fail(PBF_ERR_CUDA, "cooperative solver kernel: %s", cudaGetErrorString(ce));
    }
    for (int i = 0; i < s->p.niter; i++) {
        s->fuse_velocity = i == s->p.niter - 1;   // (pbf_stage_delta_p decides; off again right away: the stage
        rc = pbf_stage_correct_density(s);        //  entry points called one by one keep the stages apart)
        s->fuse_velocity = false;
        if (rc) return rc;
    }
    if (s->stage != ST_VELOCITY && (rc = pbf_stage_update_velocity(s))) return rc;
    if ((rc = pbf_stage_correct_velocity(s))) return rc;
    return pbf_stage_end(s);
}

static uint64_t hash_bytes(uint64_t h, const void* p, size_t bytes) { return fnv1a64(h, p, bytes); }

// pbf_step replayed from an instantiated CUDA graph: the 20-odd launches of a step become ONE submission, and the
// programmatic-dependent-launch edges between its kernels (launch.cuh) are kept. Worth it where a step is short
// (the reference's own 32 000-particle scene: 22 launches in 0.2 ms). A step is captured the first time its key
// is seen; the caller's ping-pong gives two keys. A key that keeps changing (a wall that moves every step
// changes the constants) would re-capture every step: after three misses in a row graphs pause for 32 steps.
static int step_graph(pbf_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n, void* stream) {
    uint64_t hc = 14695981039346656037ull;
    hc = hash_bytes(hc, &s->g, sizeof(s->g));
    hc = hash_bytes(hc, &s->c, sizeof(s->c));
    hc = hash_bytes(hc, &s->npass, sizeof(s->npass));
    const void* ptr[5] = {pos, npos, vel, nvel, iid};
    StepGraph* hit = nullptr;
    StepGraph* lru = nullptr;   // where a new graph goes: a free slot, else the least recently used one
    for (auto& gq : s->graphs) {
        if (gq.exec && gq.n == n && gq.stream == (cudaStream_t)stream && gq.consts_hash == hc &&
            gq.niter == s->p.niter && gq.team == s->mode.team && gq.rebin == s->mode.rebin + 2 * s->mode.staged + 4 * s->mode.paired + 8 * s->mode.morton + 16 * s->mode.coop && gq.pdl == s->mode.pdl &&
            memcmp(gq.ptr, ptr, sizeof(ptr)) == 0) {
            hit = &gq;
            break;
        }
        if (!lru || (lru->exec && (!gq.exec || gq.used < lru->used))) lru = &gq;
    }
    s->graph_clock++;
    if (!hit) {
        if (++s->graph_misses > 3 && s->mode.graph < 1) {   // the key will not settle: stop paying for captures
            s->graph_misses = 0;
            s->graph_holdoff = 32;
            return step_direct(s, pos, npos, vel, nvel, iid, n, stream);
        }
        CUDA_TRY(cudaSetDevice(s->device));
        const int64_t l0 = s->launches;
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) {
            cudaGetLastError();
            s->graph_broken = true;
            return step_direct(s, pos, npos, vel, nvel, iid, n, stream);
        }
        const int rc = step_direct(s, pos, npos, vel, nvel, iid, n, stream);
        e = cudaStreamEndCapture((cudaStream_t)stream, &graph);
        if (rc != PBF_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            s->graph_broken = true;
            s->launches = l0;
            if (rc != PBF_OK) return rc;   // (an argument error: the same with or without a graph)
            return step_direct(s, pos, npos, vel, nvel, iid, n, stream);
        }
        if (lru->exec) { cudaGraphExecDestroy(lru->exec); lru->exec = nullptr; }
        e = cudaGraphInstantiate(&lru->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            cudaGetLastError();
            lru->exec = nullptr;
            s->graph_broken = true;
            s->launches = l0;
            return step_direct(s, pos, npos, vel, nvel, iid, n, stream);
        }
        memcpy(lru->ptr, ptr, sizeof(ptr));
        lru->n = n; lru->stream = (cudaStream_t)stream; lru->consts_hash = hc;
        lru->niter = s->p.niter; lru->team = s->mode.team; lru->rebin = s->mode.rebin + 2 * s->mode.staged + 4 * s->mode.paired + 8 * s->mode.morton + 16 * s->mode.coop; lru->pdl = s->mode.pdl;
        lru->launches = s->launches - l0;
        lru->sorted_buf = s->sorted_buf; lru->cur = s->cur; lru->iters_done = s->iters_done;
        lru->cull_cur = s->cull.cur; lru->cull_holds = s->cull.holds; lru->v4 = s->v4;
        s->launches = l0;   // (counted below, when the graph runs)
        hit = lru;
    } else {
        s->graph_misses = 0;
        // the handle's view of the step in flight, as the stage functions would have left it
        int rc = pbf_stage_begin(s, pos, npos, vel, nvel, iid, n, stream);
        if (rc) return rc;
        s->sorted_buf = hit->sorted_buf; s->cur = hit->cur; s->iters_done = hit->iters_done;
        s->cull.cur = hit->cull_cur; s->cull.holds = hit->cull_holds; s->v4 = hit->v4;
        s->pos0_in_npos = false;
        s->stage = ST_IDLE;
    }
    hit->used = s->graph_clock;
    CUDA_TRY(cudaGraphLaunch(hit->exec, (cudaStream_t)stream));
    s->launches += hit->launches;
    return PBF_OK;
}

int pbf_step(pbf_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n, void* stream) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    // a graph needs a capturable stream (not the legacy default stream) and a step without event timers
    const bool capturable = stream != nullptr && (cudaStream_t)stream != cudaStreamLegacy && (cudaStream_t)stream != cudaStreamPerThread;
    const bool wanted = s->mode.graph > 0 || (s->mode.graph < 0 && n < GRAPH_AUTO_MAX);
    if (wanted && capturable && !s->timing && !s->graph_broken && n > 0) {
        if (s->graph_holdoff > 0) s->graph_holdoff--;
        else return step_graph(s, pos, npos, vel, nvel, iid, n, stream);
    }
    return step_direct(s, pos, npos, vel, nvel, iid, n, stream);
}

int pbf_step_host(pbf_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    if (n < 0 || n > s->max_particles) return fail(PBF_ERR_CAPACITY, "n exceeds max_particles");
    if (!pos || !npos || !vel || !nvel || !iid) return fail(PBF_ERR_INVALID, "null particle buffer");
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t m = (size_t)s->max_particles;
    if (!s->h_pos) {
        CUDA_TRY(cudaMalloc((void**)&s->h_pos, m * 12));
        CUDA_TRY(cudaMalloc((void**)&s->h_npos, m * 12));
        CUDA_TRY(cudaMalloc((void**)&s->h_vel, m * 12));
        CUDA_TRY(cudaMalloc((void**)&s->h_nvel, m * 12));
        CUDA_TRY(cudaMalloc((void**)&s->h_iid, m * 4));
    }
    if (!s->host_main) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s->host_main, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&s->host_copy, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s->host_ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s->host_iid_ev, cudaEventDisableTiming));
    }
    cudaStream_t st = s->host_main;
    CUDA_TRY(cudaMemcpyAsync(s->h_pos, pos, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->h_vel, vel, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    // iid is not needed before the reorder: it goes up on the copy stream, behind pos and vel on the wire,
    // while the keys are computed and sorted
    if (n > 0) {
        CUDA_TRY(cudaEventRecord(s->host_ev, st));
        CUDA_TRY(cudaStreamWaitEvent(s->host_copy, s->host_ev, 0));
        CUDA_TRY(cudaMemcpyAsync(s->h_iid, iid, (size_t)n * 4, cudaMemcpyHostToDevice, s->host_copy));
        CUDA_TRY(cudaEventRecord(s->host_iid_ev, s->host_copy));
    }
    // the stage sequence of pbf_step, with the downloads of what is final after update_velocity — the
    // positions and the sorted iid — issued on the copy stream before the XSPH sweep starts
    int rc = pbf_stage_begin(s, s->h_pos, s->h_npos, s->h_vel, s->h_nvel, s->h_iid, n, st);
    if (rc) return rc;
    if (n > 0) s->reorder_wait = s->host_iid_ev;
    if ((rc = pbf_stage_advect(s))) return rc;
    if ((rc = pbf_stage_build_grid(s))) return rc;
    // the sorted iid is final as soon as the reorder pass has run: it goes home under the Jacobi iterations (the copy
    // stream is idle until the positions are final, and everything that leaves after the last delta-p pass — 24 bytes
    // per particle — is more than the XSPH sweep can hide on PCIe)
    if (n > 0) {
        CUDA_TRY(cudaEventRecord(s->host_ev, st));
        CUDA_TRY(cudaStreamWaitEvent(s->host_copy, s->host_ev, 0));
        CUDA_TRY(cudaMemcpyAsync(iid, s->iid_sorted, (size_t)n * 4, cudaMemcpyDeviceToHost, s->host_copy));
    }
    s->tail_npos_sent = false;
    for (int i = 0; i < s->p.niter; i++) {
        s->fuse_velocity = i == s->p.niter - 1;
        s->tail_npos_host = s->fuse_velocity ? npos : nullptr;   // (pbf_stage_delta_p: the last pass in slices, positions home slice by slice)
        rc = pbf_stage_correct_density(s);
        s->fuse_velocity = false;
        s->tail_npos_host = nullptr;
        if (rc) return rc;
    }
    if (s->stage != ST_VELOCITY && (rc = pbf_stage_update_velocity(s))) return rc;
    if (n > 0 && !s->tail_npos_sent) {
        CUDA_TRY(cudaEventRecord(s->host_ev, st));
        CUDA_TRY(cudaStreamWaitEvent(s->host_copy, s->host_ev, 0));
        CUDA_TRY(cudaMemcpyAsync(npos, s->h_npos, (size_t)n * 12, cudaMemcpyDeviceToHost, s->host_copy));
    }
    // the XSPH sweep in up to four slices of the sorted order: the velocities of a finished slice go home
    // while the next slice is computed (a slice is at least 128 K particles, so small scenes take one launch)
    const int64_t slices = n >= 4 * 131072 ? 4 : n >= 2 * 131072 ? 2 : 1;
    if (s->stage != ST_VELOCITY) return fail(PBF_ERR_STATE, "step_host: velocity update missing");
    if ((rc = kernel_event(s, PBF_KERNEL_XSPH, 0))) return rc;
    s->mode.moved = s->iters_done > 0;
    for (int64_t k = 0; k < slices && n > 0; k++) {
        const int64_t a = n * k / slices, b = n * (k + 1) / slices;
        CUDA_TRY(launch_xsph(s->x[s->cur], s->cull, k == 0 ? s->n_local : 0, s->v4, s->cell_range, s->nvel + 3 * a,
                             s->iid_sorted, s->iid + a, a, b - a, HaloSync(), StatePush(), s->g, s->c, s->mode, st, &s->launches));
        CUDA_TRY(cudaEventRecord(s->host_ev, st));
        CUDA_TRY(cudaStreamWaitEvent(s->host_copy, s->host_ev, 0));
        CUDA_TRY(cudaMemcpyAsync(nvel + 3 * a, s->h_nvel + 3 * a, (size_t)(b - a) * 12, cudaMemcpyDeviceToHost, s->host_copy));
    }
    if ((rc = kernel_event(s, PBF_KERNEL_XSPH, 1))) return rc;
    s->stage = ST_XSPH;
    if ((rc = stage_event(s, 5))) return rc;
    if ((rc = pbf_stage_end(s))) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->host_copy));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PBF_OK;
}

/* ---- multi-GPU: x-slab decomposition ------------------------------------------------------ */

int pbf_slab_begin(pbf_sim* s, const pbf_slab_step* st, float* pos, float* npos, float* vel, float* nvel,
                   uint32_t* iid, void* stream) {
    if (!s || !st) return fail(PBF_ERR_INVALID, "null argument");
    const int64_t n_in = st->n_own + st->m_left + st->m_right;
    if (st->n_own < 0 || st->m_left < 0 || st->m_right < 0) return fail(PBF_ERR_INVALID, "negative particle count");
    if (st->ghost < 1) return fail(PBF_ERR_INVALID, "slab needs at least one ghost plane");
    if ((st->has_left || st->has_right) && st->x_end - st->x_begin < st->ghost)
        return fail(PBF_ERR_INVALID, "slab [%d, %d) is narrower than its %d ghost planes", st->x_begin, st->x_end, st->ghost);
    if (st->send_left_end < 0 || st->send_right_begin > st->n_own || st->send_left_end > st->n_own || st->send_right_begin < 0)
        return fail(PBF_ERR_INVALID, "send ranges outside the own particles");
    int rc = pbf_stage_begin(s, pos, npos, vel, nvel, iid, n_in, stream);
    if (rc) return rc;
    s->slab_on = true;
    s->slab = *st;
    if ((rc = refresh_consts(s))) { s->slab_on = false; refresh_consts(s); s->stage = ST_IDLE; return rc; }
    SlabInput& si = s->si;
    si.n_own = st->n_own;
    si.m_left = st->m_left;
    si.send_left_end = st->has_left ? st->send_left_end : 0;
    si.send_right_begin = st->has_right ? st->send_right_begin : st->n_own;
    // what the neighbours keep of my side: their ghost planes reach `ghost` planes into my slab
    si.need_left_below = st->has_left ? st->x_begin + st->ghost : INT32_MIN;
    si.need_right_from = st->has_right ? st->x_end - st->ghost : INT32_MAX;
    si.flags = s->flags_dev;
    // fused mode: pull the neighbours' raw particles out of their state arrays (peer memory)
    if (s->state_iid && (s->peer[0].on || s->peer[1].on) && (st->m_left > 0 || st->m_right > 0)) {
        const int which = pos == s->state[0] ? 0 : pos == s->state[1] ? 1 : -1;
        if (which < 0 || vel != s->state[2 + which] || iid != s->state_iid)
            return fail(PBF_ERR_INVALID, "fused slab step: pos / vel / iid are not the registered state arrays");
        if (getenv("PBF_HALO_TRACE")) fprintf(stderr, "[halo %p] pull wait %u (m_left %lld m_right %lld)\n", (void*)s, s->state_seq, (long long)st->m_left, (long long)st->m_right);
        CUDA_TRY(launch_halo_wait(s->peer[0].on && st->m_left > 0 ? s->sync_words + 0 : nullptr,
                                  s->peer[1].on && st->m_right > 0 ? s->sync_words + 1 : nullptr, s->state_seq,
                                  s->halo_timeout_ns, s->flags_dev, s->stream, &s->launches));
        struct { int side; int64_t src, dst, cnt; } pull[2] = {{0, st->pull_left_first, st->n_own, st->m_left},
                                                                {1, 0, st->n_own + st->m_left, st->m_right}};
        for (auto& q : pull) {
            if (q.cnt <= 0) continue;
            // the neighbours' velocity / XSPH kernels of the last step stored the particles here themselves
            // (pbf_slab_push_state): nothing to copy, the wait above was for their end-of-step signal
            if (st->pull_left_first == PBF_SLAB_STATE_PUSHED) continue;
            const pbf_sim::Peer& pr = s->peer[q.side];
            if (!pr.on || !pr.iid) return fail(PBF_ERR_STATE, "fused slab step: neighbour %d has no registered state attached", q.side);
            CUDA_TRY(cudaMemcpyAsync(pos + 3 * q.dst, pr.state[which] + 3 * q.src, (size_t)q.cnt * 12, cudaMemcpyDefault, s->stream));
            CUDA_TRY(cudaMemcpyAsync(vel + 3 * q.dst, pr.state[2 + which] + 3 * q.src, (size_t)q.cnt * 12, cudaMemcpyDefault, s->stream));
            CUDA_TRY(cudaMemcpyAsync(iid + q.dst, pr.iid + q.src, (size_t)q.cnt * 4, cudaMemcpyDefault, s->stream));
            s->launches += 3;
        }
    }
    return PBF_OK;
}

int pbf_slab_get_layout(pbf_sim* s, pbf_slab_layout* out) {
    if (!s || !out) return fail(PBF_ERR_INVALID, "null argument");
    if (!s->slab_on || !s->layout_valid) return fail(PBF_ERR_STATE, "no slab layout: pbf_slab_begin .. pbf_stage_build_grid first");
    s->layout.flags = *s->flags_host;
    *out = s->layout;
    return PBF_OK;
}

int pbf_slab_plane_counts(pbf_sim* s, int32_t x_first, int32_t count, int64_t* out) {
    if (!s || !out || count < 0) return fail(PBF_ERR_INVALID, "bad argument");
    if (!s->slab_on || !s->layout_valid) return fail(PBF_ERR_STATE, "no slab layout: pbf_slab_begin .. pbf_stage_build_grid first");
    for (int32_t k = 0; k < count; k++) {
        const int x = x_first + k;
        const int lp = x - s->g.xoff;
        out[k] = (x >= s->slab.x_begin && x < s->slab.x_end) ? s->plane_host[lp + 1] - s->plane_host[lp] : 0;
    }
    return PBF_OK;
}

int pbf_slab_halo(pbf_sim* s, int what, void** send_left, void** recv_left, void** send_right, void** recv_right) {
    if (!s || !send_left || !recv_left || !send_right || !recv_right) return fail(PBF_ERR_INVALID, "null argument");
    if (!s->slab_on || !s->layout_valid) return fail(PBF_ERR_STATE, "no slab layout: pbf_slab_begin .. pbf_stage_build_grid first");
    float4* a = nullptr;
    switch (what) {
        case PBF_HALO_LAMBDA:
            if (s->stage != ST_LAMBDA) return fail(PBF_ERR_STATE, "lambda halo: pbf_stage_lambda first");
            a = s->xl; break;
        case PBF_HALO_POSITION:
            if (s->stage != ST_DENSITY && s->stage != ST_GRID) return fail(PBF_ERR_STATE, "position halo: pbf_stage_delta_p first");
            a = s->x[s->cur]; break;
        case PBF_HALO_VELOCITY:
            if (s->stage != ST_VELOCITY) return fail(PBF_ERR_STATE, "velocity halo: pbf_stage_update_velocity first");
            a = s->xl; break;
        default: return fail(PBF_ERR_INVALID, "unknown halo selector %d", what);
    }
    const pbf_slab_layout& L = s->layout;
    *recv_left = a;
    *send_left = a + L.own_first;
    *send_right = a + L.own_first + L.own_count - L.send_right_count;
    *recv_right = a + L.own_first + L.own_count;
    return PBF_OK;
}

int pbf_slab_register_state(pbf_sim* s, float* pos_a, float* pos_b, float* vel_a, float* vel_b, uint32_t* iid) {
    if (!s || !pos_a || !pos_b || !vel_a || !vel_b || !iid) return fail(PBF_ERR_INVALID, "null argument");
    s->state[0] = pos_a; s->state[1] = pos_b; s->state[2] = vel_a; s->state[3] = vel_b;
    s->state_iid = iid;
    return PBF_OK;
}

int pbf_slab_peer_export(pbf_sim* s, pbf_slab_peer_info* out) {
    if (!s || !out) return fail(PBF_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    memset(out, 0, sizeof(*out));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "pbf_slab_peer_info::ipc holds CUDA IPC handles");
    void* arr[9] = {s->x[0], s->x[1], s->xl, s->sync_words, s->state[0], s->state[1], s->state[2], s->state[3], s->state_iid};
    out->has_state = s->state_iid != nullptr;
    for (int k = 0; k < (out->has_state ? 9 : 4); k++) {
        out->ptr[k] = (uint64_t)(uintptr_t)arr[k];
        cudaIpcMemHandle_t h;
        if (cudaIpcGetMemHandle(&h, arr[k]) == cudaSuccess) memcpy(out->ipc[k], &h, 64);
        else cudaGetLastError();   // no IPC on this platform: same-process neighbours still work
    }
    out->pid = (int64_t)getpid();
    out->device = s->device;
    out->state_capacity = s->max_particles;
    return PBF_OK;
}

int pbf_slab_peer_attach(pbf_sim* s, int side, const pbf_slab_peer_info* peer) {
    if (!s || side < 0 || side > 1) return fail(PBF_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    pbf_sim::Peer& pr = s->peer[side];
    for (void*& b : pr.ipc_base)
        if (b) { cudaIpcCloseMemHandle(b); b = nullptr; }
    pr = pbf_sim::Peer();
    if (!peer) return PBF_OK;   // detach
    const int na = peer->has_state ? 9 : 4;
    void* arr[9] = {};
    if (peer->pid == (int64_t)getpid()) {
        if (peer->device != s->device) {   // another device of this process: plain peer access
            cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(PBF_ERR_CUDA, "no peer access from device %d to %d: %s", s->device, peer->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        for (int k = 0; k < na; k++) arr[k] = (void*)(uintptr_t)peer->ptr[k];
    } else {
        for (int k = 0; k < na; k++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, peer->ipc[k], 64);
            cudaError_t e = cudaIpcOpenMemHandle(&arr[k], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                for (int j = 0; j < k; j++) { cudaIpcCloseMemHandle(arr[j]); pr.ipc_base[j] = nullptr; }
                return fail(PBF_ERR_CUDA, "cudaIpcOpenMemHandle failed for the neighbour's array %d: %s", k, cudaGetErrorString(e));
            }
            pr.ipc_base[k] = arr[k];
        }
    }
    pr.x[0] = (float4*)arr[0]; pr.x[1] = (float4*)arr[1]; pr.xl = (float4*)arr[2]; pr.sync = (uint32_t*)arr[3];
    for (int k = 0; k < 4; k++) pr.state[k] = (float*)arr[4 + k];
    pr.iid = (uint32_t*)arr[8];
    pr.state_capacity = peer->has_state ? peer->state_capacity : 0;
    pr.on = true;
    return PBF_OK;
}

// "everything I enqueued so far is complete" to both neighbours, under a fresh handshake number
static int peers_signal(pbf_sim* s) {
    s->halo_seq++;
    if (getenv("PBF_HALO_TRACE")) fprintf(stderr, "[halo %p] signal %u\n", (void*)s, s->halo_seq);
    // I am my left neighbour's RIGHT neighbour: raise its word [1]; and my right neighbour's word [0]
    CUDA_TRY(launch_halo_signal(s->peer[0].on ? s->peer[0].sync + 1 : nullptr, s->peer[1].on ? s->peer[1].sync + 0 : nullptr,
                                s->halo_seq, s->stream, &s->launches));
    return PBF_OK;
}

int pbf_slab_halo_sync(pbf_sim* s) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    if (!s->peer[0].on && !s->peer[1].on) return PBF_OK;
    if (inkernel_halo(s)) return PBF_OK;   // the pass kernels signalled, the next reader's edge blocks wait (HaloSync)
    int rc = peers_signal(s);
    if (rc) return rc;
    CUDA_TRY(launch_halo_wait(s->peer[0].on ? s->sync_words + 0 : nullptr, s->peer[1].on ? s->sync_words + 1 : nullptr,
                              s->halo_seq, s->halo_timeout_ns, s->flags_dev, s->stream, &s->launches));
    return PBF_OK;
}

int pbf_slab_push_state(pbf_sim* s, int64_t left_count, int64_t left_dst, int64_t right_first, int64_t right_dst) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    if (!s->slab_on || !s->layout_valid) return fail(PBF_ERR_STATE, "push_state: no slab layout (pbf_slab_begin .. pbf_stage_build_grid first)");
    if (s->stage == ST_VELOCITY || s->stage == ST_XSPH) return fail(PBF_ERR_STATE, "push_state: arm it before pbf_stage_update_velocity");
    if (!s->state_iid) return fail(PBF_ERR_STATE, "push_state: no registered state arrays");
    const int which = s->pos == s->state[0] ? 0 : s->pos == s->state[1] ? 1 : -1;
    if (which < 0 || s->npos != s->state[which ^ 1] || s->nvel != s->state[2 + (which ^ 1)] || s->iid != s->state_iid)
        return fail(PBF_ERR_INVALID, "push_state: the step does not run on the registered state arrays");
    const int64_t n = s->own_count;
    if (left_count < 0 || left_count > n || right_first < 0 || right_first > n || left_dst < 0 || right_dst < 0)
        return fail(PBF_ERR_INVALID, "push_state: ranges outside the %lld owned particles", (long long)n);
    StatePush sp;
    // every rank uses buffers A and B in the same rhythm (pbf_slab_register_state): my npos / nvel of this step and
    // the neighbours' are the same letter, and they are the next step's pos / vel
    const pbf_sim::Peer& L = s->peer[0];
    const pbf_sim::Peer& R = s->peer[1];
    if (L.on && L.iid && left_count > 0) {
        if (left_dst + left_count > L.state_capacity) return fail(PBF_ERR_CAPACITY, "push_state: the left neighbour's arrays hold %lld particles, %lld needed", (long long)L.state_capacity, (long long)(left_dst + left_count));
        sp.pos_l = L.state[which ^ 1]; sp.vel_l = L.state[2 + (which ^ 1)]; sp.iid_l = L.iid;
        sp.left_count = left_count; sp.left_dst = left_dst; sp.cap_l = L.state_capacity;
    }
    if (R.on && R.iid && right_first < n) {
        if (right_dst + (n - right_first) > R.state_capacity) return fail(PBF_ERR_CAPACITY, "push_state: the right neighbour's arrays hold %lld particles, %lld needed", (long long)R.state_capacity, (long long)(right_dst + n - right_first));
        sp.pos_r = R.state[which ^ 1]; sp.vel_r = R.state[2 + (which ^ 1)]; sp.iid_r = R.iid;
        sp.right_first = right_first; sp.right_dst = right_dst; sp.cap_r = R.state_capacity;
    }
    s->state_push = sp;
    return PBF_OK;
}

int pbf_slab_flags(pbf_sim* s, uint32_t* out) {
    if (!s || !out) return fail(PBF_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    *out = *s->flags_host;
    *s->flags_host = 0;
    return PBF_OK;
}

static int slab_sort_state_impl(pbf_sim* s, int32_t x_begin, int32_t x_end, int32_t has_left, int32_t has_right,
                                float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n,
                                int64_t* n_kept, void* stream) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    pbf_slab_step st{};
    st.x_begin = x_begin; st.x_end = x_end; st.ghost = 1;
    st.has_left = has_left; st.has_right = has_right;
    st.n_own = n; st.send_right_begin = n;
    int rc = pbf_slab_begin(s, &st, pos, npos, vel, nvel, iid, stream);
    if (rc) return rc;
    s->si.flags = nullptr;          // nothing was sent: no migration check
    s->si.send_left_end = 0; s->si.send_right_begin = n;
    // keys from the positions as they are: advect with dt = 0 is the identity (fma(0, v, p) == p)
    SolverConsts c0 = s->c;
    c0.dt = 0.f; c0.gravity = 0.f;
    if (s->sort_dirty) CUDA_TRY(cudaMemsetAsync(s->sort_zero, 0, s->sort_zero_capacity, s->stream));
    s->sort_dirty = false;
    CUDA_TRY(launch_advect_key(s->pos, s->vel, s->keys, s->sort_zero, nullptr, s->n, s->npass, s->si, s->g, c0, s->stream, &s->launches));
    SortScratch sc;
    sc.hist = s->sort_zero;
    sc.tile_counter = s->sort_zero + MAX_PASSES * RADIX;
    sc.tile_desc = s->sort_zero + MAX_PASSES * RADIX + MAX_PASSES;
    sc.bufs[0] = s->pairs[0];
    sc.bufs[1] = s->pairs[1];
    sc.tile_desc_words = 0;
    CUDA_TRY(launch_sort(s->keys, sc, s->n, s->npass, s->si, &s->sorted_buf, s->stream, &s->launches));
    // (no reorder pass follows here: clean the sort's scratch by hand, see pbf_sim::sort_zero)
    CUDA_TRY(cudaMemsetAsync(s->sort_zero, 0, sort_scratch_zero_bytes(s->n, s->npass), s->stream));
    if ((rc = slab_learn_layout(s))) return rc;
    s->stage = ST_IDLE;
    if (!n_kept && (s->layout.own_count != n || s->layout.own_first != 0))
        return fail(PBF_ERR_INVALID, "sort_state: %lld of %lld particles lie outside planes [%d, %d)",
                    (long long)(n - s->layout.own_count), (long long)n, x_begin, x_end);
    const int64_t kept = s->layout.own_count;
    if (n_kept) *n_kept = kept;
    // (adopt: the owned particles are the slots [own_first, own_first + kept) of the sort; what follows
    //  describes a state that holds exactly them, as a step's result would)
    CUDA_TRY(launch_gather_state(s->pairs[s->sorted_buf] + s->layout.own_first, pos, vel, iid, npos, nvel, s->iid_sorted, kept, s->stream, &s->launches));
    CUDA_TRY(cudaMemcpyAsync(iid, s->iid_sorted, (size_t)kept * 4, cudaMemcpyDeviceToDevice, s->stream));
    n = kept;
    if (s->state_iid && (s->peer[0].on || s->peer[1].on)) {   // fused mode: the sorted state is what neighbours pull from
        if ((rc = peers_signal(s))) return rc;
        s->state_seq = s->halo_seq;
    }
    return PBF_OK;
}

int pbf_slab_sort_state(pbf_sim* s, int32_t x_begin, int32_t x_end, int32_t has_left, int32_t has_right,
                        float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n, void* stream) {
    return slab_sort_state_impl(s, x_begin, x_end, has_left, has_right, pos, npos, vel, nvel, iid, n, nullptr, stream);
}

int pbf_slab_adopt_state(pbf_sim* s, int32_t x_begin, int32_t x_end, int32_t has_left, int32_t has_right,
                         float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n, int64_t* n_kept,
                         void* stream) {
    if (!n_kept) return fail(PBF_ERR_INVALID, "null argument");
    return slab_sort_state_impl(s, x_begin, x_end, has_left, has_right, pos, npos, vel, nvel, iid, n, n_kept, stream);
}

int pbf_scene_block_slice_device(const float origin[3], const int32_t n[3], float spacing, uint32_t seed,
                                 uint32_t first_iid, int32_t ix_begin, int32_t ix_end, float* d_pos,
                                 float* d_vel, uint32_t* d_iid, void* stream) {
    if (!origin || !n) return fail(PBF_ERR_INVALID, "null argument");
    if (ix_begin < 0 || ix_end > n[0] || ix_begin > ix_end) return fail(PBF_ERR_INVALID, "bad layer range");
    if (ix_end > ix_begin && (!d_pos || !d_vel || !d_iid)) return fail(PBF_ERR_INVALID, "null argument");
    CUDA_TRY(launch_scene_block(origin, n, spacing, seed, first_iid, ix_begin, ix_end, d_pos, d_vel, d_iid, (cudaStream_t)stream));
    return PBF_OK;
}

/* ---- read-backs --------------------------------------------------------------------------- */

int pbf_read(pbf_sim* s, int what, void* dst, int64_t count) {
    if (!s || !dst) return fail(PBF_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (count < 0) return fail(PBF_ERR_INVALID, "negative count");
    const bool per_particle = what != PBF_READ_CELL_START && what != PBF_READ_CELL_END;
    if (per_particle && count > s->n_local) return fail(PBF_ERR_INVALID, "count exceeds the bound particle count");
    if (!per_particle && count > s->g.ncell) return fail(PBF_ERR_INVALID, "count exceeds the cell count");
    if (count == 0) return PBF_OK;
    const KeyIdx* sorted = s->pairs[s->sorted_buf];
    if (!s->read_scratch) {
        size_t bytes = (size_t)s->max_particles * 12;
        if (bytes < (size_t)s->cell_capacity * 4) bytes = (size_t)s->cell_capacity * 4;
        CUDA_TRY(cudaMalloc((void**)&s->read_scratch, bytes));
    }
    // extract one field (stride/offset/width in 32-bit words) into tight host memory
    auto extract = [&](const void* src, int stride, int offset, int width) -> int {
        const int64_t total = count * width;
        extract_words_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s->stream>>>((const uint32_t*)src, stride, offset, width, s->read_scratch, count);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(dst, s->read_scratch, (size_t)total * 4, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        return PBF_OK;
    };
    switch (what) {
        case PBF_READ_KEY: return extract(sorted, 2, 0, 1);
        case PBF_READ_SRC_INDEX: return extract(sorted, 2, 1, 1);
        case PBF_READ_IID: return extract(s->iid_sorted, 1, 0, 1);
        case PBF_READ_CELL_START: return extract(s->cell_range, 2, 0, 1);
        case PBF_READ_CELL_END: return extract(s->cell_range, 2, 1, 1);
        case PBF_READ_NPOS: return extract(s->x[s->cur], 4, 0, 3);
        case PBF_READ_LAMBDA:
            if (s->stage != ST_DENSITY) return fail(PBF_ERR_STATE, "lambda is only valid between correct_density and update_velocity");
            return extract(s->xl, 4, 3, 1);
        case PBF_READ_RHO: return extract(s->rho, 1, 0, 1);
        case PBF_READ_POS0:  // (slab mode: the owned particles only, from slot own_first on)
            if (count > s->own_count) return fail(PBF_ERR_INVALID, "count exceeds the owned particle count");
            return extract(s->pos0_in_npos ? s->npos : s->pos, 3, 0, 3);
        case PBF_READ_VEL:
            if (s->stage != ST_VELOCITY && s->stage != ST_XSPH) return fail(PBF_ERR_STATE, "velocity not computed yet");
            return extract(s->v4, 4, 0, 3);
        case PBF_READ_NEIGHBOR_COUNT: {
            if (!s->count_scratch) CUDA_TRY(cudaMalloc((void**)&s->count_scratch, (size_t)s->max_particles * 4));
            CUDA_TRY(launch_neighbor_count(s->x[s->cur], s->cull, s->cell_range, s->count_scratch, s->n_local, s->g, s->c, s->stream));
            return extract(s->count_scratch, 1, 0, 1);
        }
        default:
            return fail(PBF_ERR_INVALID, "unknown read selector %d", what);
    }
}

int pbf_get_stats(pbf_sim* s, const float* npos, const float* nvel, int64_t n, pbf_stats* out) {
    if (!s || !npos || !nvel || !out) return fail(PBF_ERR_INVALID, "null argument");
    if (n <= 0 || n > s->max_particles) return fail(PBF_ERR_INVALID, "bad n");
    if (n > s->own_count) return fail(PBF_ERR_INVALID, "n = %lld exceeds the %lld particles of the last step", (long long)n, (long long)s->own_count);
    CUDA_TRY(cudaSetDevice(s->device));
    int nb = (int)((n + 255) / 256);
    if (nb > 1024) nb = 1024;
    CUDA_TRY(launch_stats(s->rho + s->own_first, npos, nvel, n, s->p.pho0, s->stats_partial, nb, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->stats_host, s->stats_partial, (size_t)nb * 5 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    double e_sum = 0, e_max = -1e300, ke = 0, v_max = 0, z_sum = 0;
    for (int b = 0; b < nb; b++) {
        const double* q = s->stats_host + b * 5;
        e_sum += q[0];
        if (q[1] > e_max) e_max = q[1];
        ke += q[2];
        if (q[3] > v_max) v_max = q[3];
        z_sum += q[4];
    }
    out->density_err_mean = e_sum / (double)n;
    out->density_err_max = e_max;
    out->kinetic_energy = ke;
    out->max_speed = sqrt(v_max);
    out->mean_z = z_sum / (double)n;
    return PBF_OK;
}

int pbf_state_digest_device(int device, const float* pos, const float* vel, const uint32_t* iid, int64_t n, void* stream,
                            uint64_t digest[2]) {
    if (!digest || n < 0) return fail(PBF_ERR_INVALID, "bad argument");
    if (n > 0 && (!pos || !vel || !iid)) return fail(PBF_ERR_INVALID, "null particle buffer");
    CUDA_TRY(cudaSetDevice(device));
    unsigned long long* dev = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dev, 16));
    cudaError_t e = launch_digest(pos, vel, iid, n, dev, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(digest, dev, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(PBF_ERR_CUDA, "state digest: %s", cudaGetErrorString(e));
    return PBF_OK;
}
int pbf_state_digest_host(const float* pos, const float* vel, const uint32_t* iid, int64_t n, uint64_t digest[2]) {
    if (!digest || n < 0) return fail(PBF_ERR_INVALID, "bad argument");
    if (n > 0 && (!pos || !vel || !iid)) return fail(PBF_ERR_INVALID, "null particle buffer");
    digest_host(pos, vel, iid, n, digest);
    return PBF_OK;
}

int pbf_enable_stage_timing(pbf_sim* s, int enable) {
    if (!s) return fail(PBF_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(s->device));
    if (enable && !s->ev_valid) {
        for (auto& e : s->ev) CUDA_TRY(cudaEventCreate(&e));
        for (auto& e : s->kev) CUDA_TRY(cudaEventCreate(&e));
        s->ev_valid = true;
    }
    s->timing = enable != 0;
    return PBF_OK;
}

int pbf_get_stage_ms(pbf_sim* s, float ms[5]) {
    if (!s || !ms) return fail(PBF_ERR_INVALID, "null argument");
    if (!s->timing || !s->ev_valid) return fail(PBF_ERR_STATE, "stage timing is not enabled");
    CUDA_TRY(cudaEventSynchronize(s->ev[5]));
    for (int k = 0; k < 5; k++) CUDA_TRY(cudaEventElapsedTime(&ms[k], s->ev[k], s->ev[k + 1]));
    return PBF_OK;
}

int pbf_get_kernel_ms(pbf_sim* s, float ms[PBF_KERNEL_SLOTS]) {
    if (!s || !ms) return fail(PBF_ERR_INVALID, "null argument");
    if (!s->timing || !s->ev_valid) return fail(PBF_ERR_STATE, "stage timing is not enabled");
    CUDA_TRY(cudaEventSynchronize(s->ev[5]));
    for (int k = 0; k < PBF_KERNEL_SLOTS; k++) CUDA_TRY(cudaEventElapsedTime(&ms[k], s->kev[2 * k], s->kev[2 * k + 1]));
    return PBF_OK;
}

/* ---- device memory helpers ---------------------------------------------------------------- */

int pbf_device_alloc(int device, int64_t bytes, void** out) {
    if (!out || bytes < 0) return fail(PBF_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMalloc(out, (size_t)(bytes ? bytes : 1)));
    return PBF_OK;
}
int pbf_device_free(int device, void* ptr) {
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(ptr));
    return PBF_OK;
}
int pbf_copy_h2d(void* dst, const void* src, int64_t bytes) {
    CUDA_TRY(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice));
    return PBF_OK;
}
int pbf_copy_d2h(void* dst, const void* src, int64_t bytes) {
    CUDA_TRY(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return PBF_OK;
}
int pbf_device_sync(int device) {
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaDeviceSynchronize());
    return PBF_OK;
}

int pbf_stream_create(int device, void** stream_out) {
    if (!stream_out) return fail(PBF_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st;
    CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream_out = (void*)st;
    return PBF_OK;
}
int pbf_stream_destroy(int device, void* stream) {
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaStreamDestroy((cudaStream_t)stream));
    return PBF_OK;
}
int pbf_stream_sync(int device, void* stream) {
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return PBF_OK;
}
int pbf_copy_d2h_async(void* dst, const void* src, int64_t bytes, void* stream) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return PBF_OK;
}
int pbf_device_count(int* count) {
    if (!count) return fail(PBF_ERR_INVALID, "null argument");
    CUDA_TRY(cudaGetDeviceCount(count));
    return PBF_OK;
}

int pbf_scene_block_device(const float origin[3], const int32_t n[3], float spacing, uint32_t seed,
                           uint32_t first_iid, float* d_pos, float* d_vel, uint32_t* d_iid, void* stream) {
    if (!origin || !n || !d_pos || !d_vel || !d_iid) return fail(PBF_ERR_INVALID, "null argument");
    CUDA_TRY(launch_scene_block(origin, n, spacing, seed, first_iid, 0, n[0], d_pos, d_vel, d_iid, (cudaStream_t)stream));
    return PBF_OK;
}

// ---- state files (checkpoint / resume), see include/pbf.h ----------------------------------------

// host staging that cannot throw through the C boundary
struct HostBuf {
    void* p;
    explicit HostBuf(size_t bytes) : p(malloc(bytes ? bytes : 1)) {}
    ~HostBuf() { free(p); }
    HostBuf(const HostBuf&) = delete;
    HostBuf& operator=(const HostBuf&) = delete;
};

int pbf_state_write(const char* path, const pbf_state_info* info, const float* pos, const float* vel, const uint32_t* iid) {
    if (!path || !info) return fail(PBF_ERR_INVALID, "null argument");
    if (info->n < 0 || info->n >= ((int64_t)1 << 30)) return fail(PBF_ERR_INVALID, "bad particle count");
    if (info->n > 0 && (!pos || !vel || !iid)) return fail(PBF_ERR_INVALID, "null particle buffer");
    StateHeader hd;
    memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, "PBFSTAT1", 8);
    hd.version = 1;
    hd.header_bytes = sizeof(hd);
    hd.n = info->n;
    hd.frame = info->frame;
    hd.params = info->params;
    memcpy(hd.ulim, info->ulim, sizeof(hd.ulim));
    memcpy(hd.llim, info->llim, sizeof(hd.llim));
    hd.exact_pow = info->exact_pow;
    hd.checksum = payload_checksum(pos, vel, iid, info->n);
    HostBuf tmp_buf(strlen(path) + 5);
    if (!tmp_buf.p) return fail(PBF_ERR_INVALID, "out of host memory");
    char* const tmp = (char*)tmp_buf.p;
    snprintf(tmp, strlen(path) + 5, "%s.tmp", path);
    FILE* f = fopen(tmp, "wb");
    if (!f) return fail(PBF_ERR_INVALID, "%s: cannot open for writing", tmp);
    const size_t n = (size_t)info->n;
    bool ok = fwrite(&hd, 1, sizeof(hd), f) == sizeof(hd);
    ok = ok && fwrite(pos, 12, n, f) == n && fwrite(vel, 12, n, f) == n && fwrite(iid, 4, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { remove(tmp); return fail(PBF_ERR_INVALID, "%s: write failed", tmp); }
    if (rename(tmp, path) != 0) { remove(tmp); return fail(PBF_ERR_INVALID, "%s: rename failed", path); }
    return PBF_OK;
}

int pbf_state_read_info(const char* path, pbf_state_info* out) {
    if (!path || !out) return fail(PBF_ERR_INVALID, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(PBF_ERR_INVALID, "%s: cannot open", path);
    StateHeader hd;
    int rc = read_header(f, path, &hd);
    fclose(f);
    if (rc) return rc;
    info_of(hd, out);
    return PBF_OK;
}

int pbf_state_read(const char* path, pbf_state_info* out, float* pos, float* vel, uint32_t* iid, int64_t capacity) {
    if (!path || !out) return fail(PBF_ERR_INVALID, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(PBF_ERR_INVALID, "%s: cannot open", path);
    StateHeader hd;
    int rc = read_header(f, path, &hd);
    if (rc) { fclose(f); return rc; }
    if (hd.n > capacity) { fclose(f); return fail(PBF_ERR_CAPACITY, "%s holds %lld particles, the buffers %lld", path, (long long)hd.n, (long long)capacity); }
    if (hd.n > 0 && (!pos || !vel || !iid)) { fclose(f); return fail(PBF_ERR_INVALID, "null particle buffer"); }
    const size_t n = (size_t)hd.n;
    const bool ok = fread(pos, 12, n, f) == n && fread(vel, 12, n, f) == n && fread(iid, 4, n, f) == n;
    char extra;
    const bool trailing = ok && fread(&extra, 1, 1, f) == 1;
    fclose(f);
    if (!ok) return fail(PBF_ERR_INVALID, "%s: truncated (header says %lld particles)", path, (long long)hd.n);
    if (trailing) return fail(PBF_ERR_INVALID, "%s: longer than its header says", path);
    const uint64_t sum = payload_checksum(pos, vel, iid, hd.n);
    if (sum != hd.checksum) return fail(PBF_ERR_INVALID, "%s: checksum mismatch (file %016llx, data %016llx)", path, (unsigned long long)hd.checksum, (unsigned long long)sum);
    info_of(hd, out);
    return PBF_OK;
}

int pbf_checkpoint_save(pbf_sim* s, const char* path, const float* pos, const float* vel, const uint32_t* iid, int64_t n, int64_t frame) {
    if (!s || !path) return fail(PBF_ERR_INVALID, "null argument");
    if (n < 0 || n > s->max_particles) return fail(PBF_ERR_INVALID, "bad n");
    if (n > 0 && (!pos || !vel || !iid)) return fail(PBF_ERR_INVALID, "null particle buffer");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    HostBuf h_pos((size_t)n * 12), h_vel((size_t)n * 12), h_iid((size_t)n * 4);
    if (!h_pos.p || !h_vel.p || !h_iid.p) return fail(PBF_ERR_INVALID, "out of host memory (%lld particles)", (long long)n);
    if (n > 0) {
        CUDA_TRY(cudaMemcpy(h_pos.p, pos, (size_t)n * 12, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_vel.p, vel, (size_t)n * 12, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_iid.p, iid, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
    pbf_state_info info;
    memset(&info, 0, sizeof(info));
    info.n = n;
    info.frame = frame;
    info.params = s->p;
    memcpy(info.ulim, s->ulim, sizeof(info.ulim));
    memcpy(info.llim, s->llim, sizeof(info.llim));
    info.exact_pow = s->exact_pow;
    return pbf_state_write(path, &info, (const float*)h_pos.p, (const float*)h_vel.p, (const uint32_t*)h_iid.p);
}

int pbf_checkpoint_load(pbf_sim* s, const char* path, float* pos, float* vel, uint32_t* iid, int64_t capacity, int64_t* n_out, int64_t* frame_out) {
    if (!s || !path) return fail(PBF_ERR_INVALID, "null argument");
    pbf_state_info info;
    int rc = pbf_state_read_info(path, &info);
    if (rc) return rc;
    if (info.n > capacity || info.n > s->max_particles)
        return fail(PBF_ERR_CAPACITY, "%s holds %lld particles, the handle %lld, the buffers %lld", path, (long long)info.n, (long long)s->max_particles, (long long)capacity);
    if (info.n > 0 && (!pos || !vel || !iid)) return fail(PBF_ERR_INVALID, "null particle buffer");
    HostBuf h_pos((size_t)info.n * 12), h_vel((size_t)info.n * 12), h_iid((size_t)info.n * 4);
    if (!h_pos.p || !h_vel.p || !h_iid.p) return fail(PBF_ERR_INVALID, "out of host memory (%lld particles)", (long long)info.n);
    rc = pbf_state_read(path, &info, (float*)h_pos.p, (float*)h_vel.p, (uint32_t*)h_iid.p, info.n);
    if (rc) return rc;
    // parameters, box and pow option TOGETHER, validated once (a file whose (h, box) pair fits the handle must not
    // be rejected because of an intermediate (new h, old box) state); a file that does not fit leaves the handle
    // as it was
    const pbf_params old_p = s->p;
    float old_u[3], old_l[3];
    memcpy(old_u, s->ulim, sizeof(old_u));
    memcpy(old_l, s->llim, sizeof(old_l));
    const int old_exact = s->exact_pow;
    s->p = info.params;
    memcpy(s->ulim, info.ulim, sizeof(s->ulim));
    memcpy(s->llim, info.llim, sizeof(s->llim));
    s->exact_pow = info.exact_pow ? 1 : 0;
    rc = refresh_consts(s);
    if (rc != PBF_OK) {
        char msg[sizeof(g_err)];
        snprintf(msg, sizeof(msg), "%s", g_err);
        s->p = old_p;
        memcpy(s->ulim, old_u, sizeof(old_u));
        memcpy(s->llim, old_l, sizeof(old_l));
        s->exact_pow = old_exact;
        refresh_consts(s);
        return fail(rc, "%s does not fit this handle: %s", path, msg);
    }
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (info.n > 0) {
        CUDA_TRY(cudaMemcpy(pos, h_pos.p, (size_t)info.n * 12, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(vel, h_vel.p, (size_t)info.n * 12, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(iid, h_iid.p, (size_t)info.n * 4, cudaMemcpyHostToDevice));
    }
    if (n_out) *n_out = info.n;
    if (frame_out) *frame_out = info.frame;
    return PBF_OK;
}

}  // extern "C"
