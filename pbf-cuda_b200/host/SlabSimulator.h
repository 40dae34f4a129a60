// SlabSimulator.h — one scene on several GPUs from C++: the x-slab decomposition of the PBF step
//  driven through the C-ABI (include/pbf.h "multi-GPU"), one host thread per rank.
//
// The reference's Simulator (fluids/Simulator.h:7-61) is single-GPU; this class is the multi-GPU host
// around G library handles, the C++ twin of pbf-cuda_b200/slab.py (same planner, same step protocol,
// same results bit for bit). Ranks live in ONE process, so the neighbours' arrays are attached as plain
// device pointers (peer access) and the whole exchange is the library's fused transport: raw boundary
// state pulled with peer-memory copies, ghost values stored by the pass kernels straight into the
// neighbour's slots, flag handshakes — no NCCL, no MPI. Several ranks may share a device (each on its own
// stream), which is how the one-GPU test box runs it.
//
// Plain C++11 + pthreads; needs no CUDA headers.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pbf.h"

namespace pbfslab {

struct Block { float origin[3]; int32_t n[3]; };
struct Scene {
    float ulim[3], llim[3];
    std::vector<Block> blocks;   // jittered lattice blocks (pbf_scene_block_*), spacing 0.05, seed 27
};

class Barrier {
public:
    explicit Barrier(int n) : m_n(n), m_count(0), m_gen(0), m_failed(false) {}
    // returns false if a rank failed (everybody leaves)
    bool wait() {
        std::unique_lock<std::mutex> lk(m_mu);
        if (m_failed) return false;
        const int gen = m_gen;
        if (++m_count == m_n) { m_count = 0; m_gen++; m_cv.notify_all(); return true; }
        m_cv.wait(lk, [&] { return gen != m_gen || m_failed; });
        return !m_failed;
    }
    void fail() { std::lock_guard<std::mutex> lk(m_mu); m_failed = true; m_cv.notify_all(); }
private:
    std::mutex m_mu; std::condition_variable m_cv; int m_n, m_count, m_gen; bool m_failed;
};

// ---- planning: the same arithmetic as slab.py::plan_boundaries / exchange_plan ---------------------------

inline bool plan_boundaries(const std::vector<int64_t>& plane_totals, int world, int min_width, const std::vector<int>* old,
                            int reach, std::vector<int>* out) {
    const int planes = (int)plane_totals.size();
    if ((int64_t)world * min_width > planes) return false;
    std::vector<int64_t> cum(planes + 1, 0);
    for (int x = 0; x < planes; x++) cum[x + 1] = cum[x] + plane_totals[x];
    const int64_t total = cum[planes];
    std::vector<int> b(world + 1, 0), lo(world + 1, 0), hi(world + 1, 0);
    b[world] = lo[world] = hi[world] = planes;
    for (int r = 1; r < world; r++) {
        const double target = (double)total * r / world;
        int x = 0;
        while (x < planes && (double)cum[x] < target) x++;            // searchsorted(cum, target, "left")
        if (x > 0) {
            const double dl = (double)cum[x - 1] - target, dr = (double)cum[x < planes ? x : planes] - target;
            if ((dl < 0 ? -dl : dl) <= (dr < 0 ? -dr : dr)) x--;
        }
        b[r] = x;
        lo[r] = r * min_width;
        hi[r] = planes - (world - r) * min_width;
        if (old) {
            if ((*old)[r - 1] + reach > lo[r]) lo[r] = (*old)[r - 1] + reach;
            if ((*old)[r + 1] - reach < hi[r]) hi[r] = (*old)[r + 1] - reach;
        }
    }
    for (int r = 1; r < world; r++) {
        int v = b[r];
        if (v < lo[r]) v = lo[r];
        if (v < b[r - 1] + min_width) v = b[r - 1] + min_width;
        if (v > hi[r]) v = hi[r];
        b[r] = v;
    }
    for (int r = world - 1; r > 0; r--) {
        int v = b[r];
        if (v > b[r + 1] - min_width) v = b[r + 1] - min_width;
        if (v < lo[r]) v = lo[r];
        b[r] = v;
    }
    bool ok = true;
    for (int r = 0; r < world; r++) ok = ok && b[r + 1] - b[r] >= min_width;
    for (int r = 1; r < world; r++) ok = ok && lo[r] <= b[r] && b[r] <= hi[r];
    if (!ok) {
        if (!old) return false;
        *out = *old;
        return true;
    }
    *out = b;
    return true;
}

struct ExchangePlan { int64_t send_left_end, send_right_begin, m_left, m_right, pull_left_first; };

inline int clipi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// counts: world rows of `planes` entries (row r = particles rank r owns per plane after the previous sort)
inline ExchangePlan exchange_plan(const std::vector<int64_t>& counts, int planes, const std::vector<int>& old,
                                  const std::vector<int>& nw, int rank, int reach) {
    const int world = (int)old.size() - 1;
    auto off = [&](int r, int x) { int64_t s = 0; for (int k = 0; k < x; k++) s += counts[(size_t)r * planes + k]; return s; };
    auto sum = [&](int r, int a, int b) { int64_t s = 0; for (int k = a; k < b; k++) s += counts[(size_t)r * planes + k]; return s; };
    ExchangePlan p;
    const int64_t n_own = off(rank, planes);
    p.send_left_end = 0; p.send_right_begin = n_own; p.m_left = p.m_right = 0; p.pull_left_first = 0;
    if (rank > 0) {
        p.send_left_end = off(rank, clipi(nw[rank] + reach, old[rank], old[rank + 1]));
        const int first = clipi(nw[rank] - reach, old[rank - 1], old[rank]);
        p.m_left = sum(rank - 1, first, old[rank]);
        p.pull_left_first = off(rank - 1, first);
    }
    if (rank < world - 1) {
        p.send_right_begin = off(rank, clipi(nw[rank + 1] - reach, old[rank], old[rank + 1]));
        const int last = clipi(nw[rank + 1] + reach, old[rank + 1], old[rank + 2]);
        p.m_right = sum(rank + 1, old[rank + 1], last);
    }
    return p;
}

// ---- the run ---------------------------------------------------------------------------------------------

class SlabRun {
public:
    // devices[r] = CUDA device of rank r (repeat a device to emulate several ranks on it)
    SlabRun(const pbf_params& params, const Scene& scene, const std::vector<int>& devices, int ghost, int margin, int replan_every)
        : m_params(params), m_scene(scene), m_dev(devices), m_world((int)devices.size()), m_ghost(ghost), m_margin(margin),
          m_reach(ghost + margin), m_replan(replan_every), m_barrier((int)devices.size()), m_steps_done(0) {
        for (int a = 0; a < 3; a++) {
            const float diff = scene.ulim[a] - scene.llim[a];
            m_dims[a] = (int)ceilf(diff / params.h);
        }
        m_planes = m_dims[0];
        m_rank.resize(m_world);
        m_counts[0].assign((size_t)m_world * m_planes, 0);
        m_counts[1].assign((size_t)m_world * m_planes, 0);
        m_hist.assign((size_t)m_world * m_planes, 0);
    }
    ~SlabRun() { destroy(); }

    int world() const { return m_world; }
    int planes() const { return m_planes; }
    const std::vector<int>& bounds() const { return m_bounds; }
    const std::string& error() const { return m_error; }
    int64_t particles(int r) const { return m_rank[r].n_own; }
    int64_t launches(int r) const { return pbf_launch_count(m_rank[r].sim); }
    pbf_sim* handle(int r) { return m_rank[r].sim; }

    // Plans the slabs from the scene's per-plane particle counts, lets every rank generate and adopt its
    // part, attaches the neighbours. false + error() on failure.
    bool init() { return run_all([this](int r) { return init_rank(r); }); }
    // Advances the scene `n` steps on all ranks; returns after the device work of the last step completed.
    bool step(int n) {
        const bool ok = run_all([this, n](int r) { return step_rank(r, n); });
        m_steps_done += n;
        return ok;
    }
    // Final state of rank r (host copies): positions, velocities, ids of its own particles, cell-sorted.
    bool download(int r, std::vector<float>* pos, std::vector<float>* vel, std::vector<uint32_t>* iid) {
        Rank& k = m_rank[r];
        pos->resize((size_t)k.n_own * 3); vel->resize((size_t)k.n_own * 3); iid->resize((size_t)k.n_own);
        if (k.n_own == 0) return true;
        return pbf_copy_d2h(pos->data(), k.pos, k.n_own * 12) == PBF_OK && pbf_copy_d2h(vel->data(), k.vel, k.n_own * 12) == PBF_OK &&
               pbf_copy_d2h(iid->data(), k.iid, k.n_own * 4) == PBF_OK;
    }
    bool stats(int r, pbf_stats* out) { Rank& k = m_rank[r]; return pbf_get_stats(k.sim, k.pos, k.vel, k.n_own, out) == PBF_OK; }

private:
    struct Rank {
        pbf_sim* sim = nullptr;
        void* stream = nullptr;
        float *pos = nullptr, *npos = nullptr, *vel = nullptr, *nvel = nullptr;   // current / next state (swapped per step)
        uint32_t* iid = nullptr;
        void* alloc[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        int64_t capacity = 0, n_own = 0;
        pbf_slab_peer_info info;
        // the NEXT step's plan, made during this step (its boundaries and this rank's exchange), when this step's kernels
        // were told to deliver the boundary planes' raw state themselves (pbf_slab_push_state)
        bool planned = false;
        std::vector<int> next_bounds;
        ExchangePlan next_xp;
    };

    template <class F> bool run_all(F f) {
        std::vector<std::thread> th;
        std::vector<int> ok(m_world, 0);
        for (int r = 0; r < m_world; r++)
            th.emplace_back([&, r] { ok[r] = f(r) ? 1 : 0; if (!ok[r]) m_barrier.fail(); });
        for (auto& t : th) t.join();
        for (int r = 0; r < m_world; r++) if (!ok[r]) return false;
        return true;
    }
    bool fail(int r, const char* what) {
        std::lock_guard<std::mutex> lk(m_mu);
        if (m_error.empty()) m_error = std::string("rank ") + std::to_string(r) + ": " + what + ": " + pbf_last_error();
        return false;
    }
#define SLAB_TRY(call) do { if ((call) != PBF_OK) return fail(r, #call); } while (0)

    // lattice layers of block b that can reach planes [x0, x1): x = origin + (ix + 0.5) d + [0, 0.2 d)
    void layers_for(const Block& b, int x0, int x1, int* lo, int* hi) const {
        const double d = 0.05, h = m_params.h;
        int a = (int)floor((x0 * h + m_scene.llim[0] - b.origin[0]) / d - 0.7) - 1;
        int e = (int)ceil((x1 * h + m_scene.llim[0] - b.origin[0]) / d - 0.5) + 1;
        *lo = clipi(a, 0, b.n[0]);
        *hi = clipi(e, 0, b.n[0]);
    }

    bool init_rank(int r) {
        Rank& k = m_rank[r];
        const int dev = m_dev[r];
        const bool L = r > 0, R = r < m_world - 1;
        SLAB_TRY(pbf_stream_create(dev, &k.stream));
        // 1. histogram of my share of the lattice layers over ALL planes, with a throw-away handle
        int64_t share = 0;
        std::vector<int> sa(m_scene.blocks.size()), sb(m_scene.blocks.size());
        for (size_t b = 0; b < m_scene.blocks.size(); b++) {
            const Block& bl = m_scene.blocks[b];
            sa[b] = (int)((int64_t)bl.n[0] * r / m_world); sb[b] = (int)((int64_t)bl.n[0] * (r + 1) / m_world);
            share += (int64_t)(sb[b] - sa[b]) * bl.n[1] * bl.n[2];
        }
        {
            pbf_sim* tmp = nullptr;
            SLAB_TRY(pbf_create(&m_params, m_scene.ulim, m_scene.llim, share > 0 ? share : 1, dev, &tmp));
            void* buf[5];
            for (int q = 0; q < 5; q++) SLAB_TRY(pbf_device_alloc(dev, (share > 0 ? share : 1) * (q < 4 ? 12 : 4), &buf[q]));
            int64_t off = 0, first_iid = 0;
            for (size_t b = 0; b < m_scene.blocks.size(); b++) {
                const Block& bl = m_scene.blocks[b];
                SLAB_TRY(pbf_scene_block_slice_device(bl.origin, bl.n, 0.05f, 27, (uint32_t)first_iid, sa[b], sb[b], (float*)buf[0] + 3 * off,
                                                      (float*)buf[2] + 3 * off, (uint32_t*)buf[4] + off, k.stream));
                off += (int64_t)(sb[b] - sa[b]) * bl.n[1] * bl.n[2];
                first_iid += (int64_t)bl.n[0] * bl.n[1] * bl.n[2];
            }
            if (share > 0) {
                SLAB_TRY(pbf_slab_sort_state(tmp, 0, m_planes, 0, 0, (float*)buf[0], (float*)buf[1], (float*)buf[2], (float*)buf[3],
                                             (uint32_t*)buf[4], share, k.stream));
                SLAB_TRY(pbf_slab_plane_counts(tmp, 0, m_planes, &m_hist[(size_t)r * m_planes]));
            }
            SLAB_TRY(pbf_stream_sync(dev, k.stream));
            for (int q = 0; q < 5; q++) pbf_device_free(dev, buf[q]);
            pbf_destroy(tmp);
        }
        if (!m_barrier.wait()) return false;
        // 2. the plan (every rank computes the same one), capacity, the handle
        std::vector<int64_t> totals(m_planes, 0);
        int64_t most = 0;
        for (int x = 0; x < m_planes; x++) {
            for (int q = 0; q < m_world; q++) totals[x] += m_hist[(size_t)q * m_planes + x];
            if (totals[x] > most) most = totals[x];
        }
        std::vector<int> b;
        if (!plan_boundaries(totals, m_world, m_world > 1 ? 2 * m_reach : 1, nullptr, m_reach, &b)) {
            std::lock_guard<std::mutex> lk(m_mu);
            if (m_error.empty()) m_error = "the grid has too few cell planes for this many ranks with this ghost + margin";
            return false;
        }
        if (r == 0) m_bounds = b;
        int64_t per_rank = 0;
        for (int q = 0; q < m_world; q++) {
            int64_t c = 0;
            for (int x = b[q]; x < b[q + 1]; x++) c += totals[x];
            if (c > per_rank) per_rank = c;
        }
        k.capacity = (int64_t)(1.3 * (double)per_rank) + 2 * (2 * m_ghost + m_margin) * most + 4096;
        SLAB_TRY(pbf_create(&m_params, m_scene.ulim, m_scene.llim, k.capacity, dev, &k.sim));
        for (int q = 0; q < 5; q++) SLAB_TRY(pbf_device_alloc(dev, k.capacity * (q < 4 ? 12 : 4), &k.alloc[q]));
        k.pos = (float*)k.alloc[0]; k.npos = (float*)k.alloc[1]; k.vel = (float*)k.alloc[2]; k.nvel = (float*)k.alloc[3];
        k.iid = (uint32_t*)k.alloc[4];
        // 3. neighbours (same process: plain pointers), before the first state is published
        SLAB_TRY(pbf_slab_register_state(k.sim, k.pos, k.npos, k.vel, k.nvel, k.iid));
        SLAB_TRY(pbf_slab_peer_export(k.sim, &k.info));
        if (!m_barrier.wait()) return false;
        if (L) SLAB_TRY(pbf_slab_peer_attach(k.sim, 0, &m_rank[r - 1].info));
        if (R) SLAB_TRY(pbf_slab_peer_attach(k.sim, 1, &m_rank[r + 1].info));
        if (!m_barrier.wait()) return false;
        // 4. my particles: the lattice layers that can reach my planes, generated generously, adopted exactly
        int64_t n = 0, first_iid = 0;
        for (size_t q = 0; q < m_scene.blocks.size(); q++) {
            const Block& bl = m_scene.blocks[q];
            int lo, hi;
            layers_for(bl, b[r], b[r + 1], &lo, &hi);
            const int64_t m = (int64_t)(hi - lo) * bl.n[1] * bl.n[2];
            if (n + m > k.capacity) return fail(r, "scene part exceeds the rank's capacity");
            SLAB_TRY(pbf_scene_block_slice_device(bl.origin, bl.n, 0.05f, 27, (uint32_t)first_iid, lo, hi, k.pos + 3 * n, k.vel + 3 * n,
                                                  k.iid + n, k.stream));
            n += m;
            first_iid += (int64_t)bl.n[0] * bl.n[1] * bl.n[2];
        }
        SLAB_TRY(pbf_slab_adopt_state(k.sim, b[r], b[r + 1], L, R, k.pos, k.npos, k.vel, k.nvel, k.iid, n, &k.n_own, k.stream));
        std::swap(k.pos, k.npos); std::swap(k.vel, k.nvel);
        SLAB_TRY(pbf_slab_plane_counts(k.sim, 0, m_planes, &m_counts[0][(size_t)r * m_planes]));
        SLAB_TRY(pbf_stream_sync(dev, k.stream));
        return m_barrier.wait();
    }

    bool step_rank(int r, int nsteps) {
        Rank& k = m_rank[r];
        const bool L = r > 0, R = r < m_world - 1;
        std::vector<int> old = m_bounds, nw;
        for (int s = 0; s < nsteps; s++) {
            const int64_t t = m_steps_done + s;
            const std::vector<int64_t>& cur = m_counts[t & 1];
            std::vector<int64_t>& nxt = m_counts[(t + 1) & 1];
            nw = old;
            ExchangePlan xp;
            const bool pushed = k.planned;   // the neighbours' velocity / XSPH kernels delivered the raw state last step
            if (pushed) {
                nw = k.next_bounds;
                xp = k.next_xp;
                k.planned = false;
            } else {
                if (m_replan && t && t % m_replan == 0 && m_world > 1) {
                    std::vector<int64_t> totals(m_planes, 0);
                    for (int x = 0; x < m_planes; x++) for (int q = 0; q < m_world; q++) totals[x] += cur[(size_t)q * m_planes + x];
                    plan_boundaries(totals, m_world, 2 * m_reach, &old, m_reach, &nw);
                }
                xp = exchange_plan(cur, m_planes, old, nw, r, m_reach);
            }
            if (k.n_own + xp.m_left + xp.m_right > k.capacity) return fail(r, "the rank would hold more particles than its capacity");
            pbf_slab_step st;
            memset(&st, 0, sizeof(st));
            st.x_begin = nw[r]; st.x_end = nw[r + 1]; st.ghost = m_ghost; st.has_left = L; st.has_right = R;
            st.n_own = k.n_own; st.m_left = xp.m_left; st.m_right = xp.m_right;
            st.send_left_end = xp.send_left_end; st.send_right_begin = xp.send_right_begin; st.pull_left_first = pushed ? (int64_t)PBF_SLAB_STATE_PUSHED : xp.pull_left_first;
            SLAB_TRY(pbf_slab_begin(k.sim, &st, k.pos, k.npos, k.vel, k.nvel, k.iid, k.stream));   // pulls the raw boundary state (or waits for the pushed one)
            SLAB_TRY(pbf_stage_advect(k.sim));
            SLAB_TRY(pbf_stage_build_grid(k.sim));                                               // sort, layout (one host sync)
            pbf_slab_layout lay;
            SLAB_TRY(pbf_slab_get_layout(k.sim, &lay));
            if (lay.flags) return flag_error(r, lay.flags);
            for (int it = 0; it < m_params.niter; it++) {
                SLAB_TRY(pbf_stage_lambda(k.sim));
                if (it == 0) {   // replicate the per-plane counts while the device works on the first pass
                    SLAB_TRY(pbf_slab_plane_counts(k.sim, 0, m_planes, &nxt[(size_t)r * m_planes]));
                    if (!m_barrier.wait()) return false;
                }
                SLAB_TRY(pbf_slab_halo_sync(k.sim));
                SLAB_TRY(pbf_stage_delta_p(k.sim));
                SLAB_TRY(pbf_slab_halo_sync(k.sim));
            }
            if (m_params.niter == 0) {
                SLAB_TRY(pbf_slab_plane_counts(k.sim, 0, m_planes, &nxt[(size_t)r * m_planes]));
                if (!m_barrier.wait()) return false;
            }
            if (m_world > 1 && m_push) {
                // the next step's plan is a function of `nxt` alone (replicated since the barrier above): make it now and
                // let this step's velocity / XSPH kernels store the boundary planes' final state straight into the
                // neighbours' next input arrays — behind the neighbour's own particles, and on its right-hand side behind
                // what ITS left neighbour delivers (slab.py SlabSimulator._plan_next is the same code)
                std::vector<int> nb = nw;
                if (m_replan && (t + 1) % m_replan == 0) {
                    std::vector<int64_t> totals(m_planes, 0);
                    for (int x = 0; x < m_planes; x++) for (int q = 0; q < m_world; q++) totals[x] += nxt[(size_t)q * m_planes + x];
                    plan_boundaries(totals, m_world, 2 * m_reach, &nw, m_reach, &nb);
                }
                auto owned = [&](int q) { int64_t c = 0; for (int x = 0; x < m_planes; x++) c += nxt[(size_t)q * m_planes + x]; return c; };
                const ExchangePlan nx = exchange_plan(nxt, m_planes, nw, nb, r, m_reach);
                int64_t left_dst = 0, right_dst = 0;
                if (L) left_dst = owned(r - 1) + exchange_plan(nxt, m_planes, nw, nb, r - 1, m_reach).m_left;
                if (R) right_dst = owned(r + 1);
                SLAB_TRY(pbf_slab_push_state(k.sim, L ? nx.send_left_end : 0, left_dst, R ? nx.send_right_begin : owned(r), right_dst));
                k.planned = true;
                k.next_bounds = nb;
                k.next_xp = nx;
            }
            SLAB_TRY(pbf_stage_update_velocity(k.sim));
            SLAB_TRY(pbf_slab_halo_sync(k.sim));
            SLAB_TRY(pbf_stage_correct_velocity(k.sim));
            SLAB_TRY(pbf_stage_end(k.sim));
            std::swap(k.pos, k.npos); std::swap(k.vel, k.nvel);
            k.n_own = lay.own_count;
            old = nw;
        }
        uint32_t flags = 0;
        SLAB_TRY(pbf_slab_flags(k.sim, &flags));   // synchronises the rank's stream
        if (flags) return flag_error(r, flags);
        if (!m_barrier.wait()) return false;
        if (r == 0) m_bounds = old;
        return m_barrier.wait();
    }
    bool flag_error(int r, uint32_t f) {
        std::lock_guard<std::mutex> lk(m_mu);
        if (m_error.empty())
            m_error = "rank " + std::to_string(r) + ": " +
                      (f & PBF_SLAB_FLAG_MIGRATION ? "a particle travelled more planes in one step than `margin` covers; " : "") +
                      (f & PBF_SLAB_FLAG_GHOST ? "a particle drifted beyond the ghost planes during the iterations; " : "") +
                      (f & PBF_SLAB_FLAG_TIMEOUT ? "a neighbour's halo flag did not arrive; " : "") + "results are not exact";
        return false;
    }
    void destroy() {
        for (int r = 0; r < m_world; r++) {
            Rank& k = m_rank[r];
            if (k.stream) pbf_stream_sync(m_dev[r], k.stream);
        }
        for (int r = 0; r < m_world; r++) {
            Rank& k = m_rank[r];
            if (k.sim) pbf_destroy(k.sim);
            for (int q = 0; q < 5; q++) if (k.alloc[q]) pbf_device_free(m_dev[r], k.alloc[q]);
            if (k.stream) pbf_stream_destroy(m_dev[r], k.stream);
            k = Rank();
        }
    }
#undef SLAB_TRY

    pbf_params m_params;
    Scene m_scene;
    std::vector<int> m_dev;
    int m_world, m_ghost, m_margin, m_reach, m_replan;
    bool m_push = getenv("PBF_SLAB_PUSH") == nullptr || getenv("PBF_SLAB_PUSH")[0] != '0';   // (A/B switch, like slab.py)
    int m_dims[3], m_planes;
    std::vector<Rank> m_rank;
    std::vector<int> m_bounds;
    std::vector<int64_t> m_counts[2], m_hist;   // replicated per-plane counts (double-buffered by step parity)
    Barrier m_barrier;
    std::mutex m_mu;
    std::string m_error;
    int64_t m_steps_done;
};

}  // namespace pbfslab
