// pbf_headless.cpp — headless C++ harness: the reference's application loop without the viewer.
//
// Restates what FluidSystem does around the solver (fluids/FluidSystem.cpp:8-120, fluids/main.cpp):
// default parameters, the box, the double-dam scene, five particle buffers sized MAX_PARTICLE_NUM,
// initSource(), then `stepSimulate()` in a loop — loadParams(), optional sweeping wall
// (ulim + A_ulim*sin(w*(frame-start))), step() with the ping-pong roles, frameCount++ — written
// against the shim headers so it reads like the reference's own code. Calls CUDA only through the
// C-ABI (libpbf_b200.so); compiles with plain g++.
//
//   pbf_headless [steps=100] [moving=0] [dump.bin] [--resume FILE] [--save FILE] [--save-every K]
//                [--stats FILE] [--stats-every K] [--emitter TOTAL]
// prints per-run statistics and, with a third argument, dumps the final
// (npos, nvel, iid) so tests can compare it with the Python path bit for bit.
// SURVEY.md 8(f) rank 1 (the reference has neither): --save writes a state file (include/pbf.h
// pbf_checkpoint_save) after the last step and, with --save-every, every K steps; --resume continues
// such a file — same bits as the uninterrupted run, because the state file carries the particle
// order, the parameters, the box and the frame counter of the wall schedule; --stats appends one JSON
// line per K steps (density error, kinetic energy, max speed, mean height, ms per step).
// --emitter TOTAL replaces the double dam by an EmitterSource (a 16x16 jet, one
// layer every 2 steps up to TOTAL particles) and calls stepSource() before every step — the reference
// defines FluidSystem::stepSource() (FluidSystem.cpp:92-97) for exactly this but its main loop never
// calls it, both shipped sources being static.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <string>
#include <vector>

#include "ParticleSource.h"

struct FluidSystemHeadless {
    ParticleSource* m_source;
    Simulator* m_simulator;
    uint d_pos, d_npos, d_vel, d_nvel, d_iid;
    bool m_tictoc;
    int m_nparticle, frameCount, startMovingFrame;
    bool moving;
    float3 m_ulim, m_llim, m_A_ulim, m_A_llim;
    float m_w;

    FluidSystemHeadless() : m_tictoc(false), frameCount(0), startMovingFrame(0), moving(false) {
        GUIParams& params = GUIParams::getInstance();   // defaults of FluidSystem.cpp:15-32 set on first use
        m_ulim = make_float3(2.f, 2.f, 4.f);             // FluidSystem.cpp:34-38
        m_llim = make_float3(-2.f, -2.f, 0.f);
        m_A_llim = make_float3(0.f, 0.f, 0.f);
        m_A_ulim = make_float3(2.f, 0.f, 0.f);
        m_w = 0.05;
        // the sweep widens the box to ulim.x = 4: construct on the widest box so the cell table holds it
        m_simulator = new Simulator(params, make_float3(4.f, 2.f, 4.f), m_llim);
        m_simulator->setLim(m_ulim, m_llim);
        float dd = 1.f / 20;                              // FluidSystem.cpp:55-61
        float d1 = dd * 20, d2 = dd * 20, d3 = dd * 40;
        m_source = new DoubleDamSource(
            make_float3(-1.8f, 1.8f, 3.8f), make_float3(-1.8f + d1, 1.8f - d2, 3.8f - d3), make_int3(20, 20, 40),
            make_float3(1.8f - d1, -1.8f + d2, 3.8f), make_float3(1.8f, -1.8f, 3.8f - d3), make_int3(20, 20, 40));
        m_nparticle = 2 * 20 * 20 * 40;
        DeviceBuffers& b = DeviceBuffers::getInstance();  // FluidSystem.cpp:64-84
        d_pos = b.create(MAX_PARTICLE_NUM * sizeof(float3));
        d_npos = b.create(MAX_PARTICLE_NUM * sizeof(float3));
        d_vel = b.create(MAX_PARTICLE_NUM * sizeof(float3));
        d_nvel = b.create(MAX_PARTICLE_NUM * sizeof(float3));
        d_iid = b.create(MAX_PARTICLE_NUM * sizeof(uint));
    }
    ~FluidSystemHeadless() { delete m_simulator; delete m_source; }

    void initSource() { m_nparticle = m_source->initialize(d_pos, d_vel, d_iid, MAX_PARTICLE_NUM); }   // FluidSystem.cpp:87-90
    void stepSource() {                                   // FluidSystem.cpp:92-97
        if (!m_tictoc) m_nparticle = m_source->update(d_pos, d_vel, d_iid, MAX_PARTICLE_NUM);
        else m_nparticle = m_source->update(d_npos, d_nvel, d_iid, MAX_PARTICLE_NUM);
    }
    // --emitter: a jet from the x = -1.9 wall instead of the two blocks (EmitterSource, ParticleSource.h)
    void useEmitter(int total) {
        delete m_source;
        m_source = new EmitterSource(make_float3(-1.9f, -0.4f, 2.0f), 16, 16, 0.05f, make_float3(3.f, 0.f, 0.f), 2, total);
    }

    void stepSimulate() {                                 // FluidSystem.cpp:99-120
        m_simulator->loadParams();
        if (moving) {
            float t = m_w * (frameCount - startMovingFrame);
            float phi = sin(t);
            m_simulator->setLim(m_ulim + m_A_ulim * phi, m_llim + m_A_llim * phi);
        }
        if (!m_tictoc) m_simulator->step(d_pos, d_npos, d_vel, d_nvel, d_iid, m_nparticle);
        else m_simulator->step(d_npos, d_pos, d_nvel, d_vel, d_iid, m_nparticle);
        m_tictoc = !m_tictoc;
        frameCount++;
    }
    uint currentPos() const { return m_tictoc ? d_npos : d_pos; }   // what render() would draw (FluidSystem.cpp:131-135)
    uint currentVel() const { return m_tictoc ? d_nvel : d_vel; }
};

static void writeStats(FILE* f, FluidSystemHeadless& fluids, int step, double ms_per_step) {
    DeviceBuffers& b = DeviceBuffers::getInstance();
    pbf_stats st;
    checkPbf(pbf_get_stats(fluids.m_simulator->handle(), (const float*)b.ptr(fluids.currentPos()),
                           (const float*)b.ptr(fluids.currentVel()), fluids.m_nparticle, &st));
    fprintf(f, "{\"step\": %d, \"ms_per_step\": %.6f, \"density_err_mean\": %.9g, \"density_err_max\": %.9g, "
               "\"kinetic_energy\": %.9g, \"max_speed\": %.9g, \"mean_z\": %.9g}\n",
            step, ms_per_step, st.density_err_mean, st.density_err_max, st.kinetic_energy, st.max_speed, st.mean_z);
    fflush(f);
}

int main(int argc, char** argv) {
    // positional arguments first, then --flags
    std::vector<const char*> posarg;
    const char *resume = 0, *save = 0, *stats_path = 0;
    int save_every = 0, stats_every = 10, emitter_total = 0;
    for (int a = 1; a < argc; a++) {
        const std::string k = argv[a];
        const bool has_val = a + 1 < argc;
        if (k == "--resume" && has_val) resume = argv[++a];
        else if (k == "--save" && has_val) save = argv[++a];
        else if (k == "--save-every" && has_val) save_every = atoi(argv[++a]);
        else if (k == "--stats" && has_val) stats_path = argv[++a];
        else if (k == "--stats-every" && has_val) stats_every = atoi(argv[++a]);
        else if (k == "--emitter" && has_val) emitter_total = atoi(argv[++a]);
        else if (k.size() > 2 && k[0] == '-' && k[1] == '-') { fprintf(stderr, "unknown or incomplete option %s\n", argv[a]); return 2; }
        else posarg.push_back(argv[a]);
    }
    const int steps = posarg.size() > 0 ? atoi(posarg[0]) : 100;
    const int moving = posarg.size() > 1 ? atoi(posarg[1]) : 0;
    const char* dump = posarg.size() > 2 ? posarg[2] : 0;
    if (stats_every < 1) stats_every = 1;
    FluidSystemHeadless fluids;
    fluids.moving = moving != 0;
    if (emitter_total > 0) fluids.useEmitter(emitter_total);
    pbf_sim* h = fluids.m_simulator->handle();
    DeviceBuffers& b = DeviceBuffers::getInstance();
    if (resume) {
        int64_t n = 0, frame = 0;
        checkPbf(pbf_checkpoint_load(h, resume, (float*)b.ptr(fluids.d_pos), (float*)b.ptr(fluids.d_vel),
                                     (uint32_t*)b.ptr(fluids.d_iid), MAX_PARTICLE_NUM, &n, &frame));
        fluids.m_simulator->saveParams();   // the file's parameters become the GUIParams the loop re-loads every step
        fluids.m_nparticle = (int)n;
        fluids.frameCount = (int)frame;
        fluids.m_tictoc = false;
        if (emitter_total > 0) static_cast<EmitterSource*>(fluids.m_source)->restore((int)n, (int)frame);
    } else {
        fluids.initSource();
    }
    FILE* stats = 0;
    if (stats_path) {
        stats = fopen(stats_path, resume ? "a" : "w");
        if (!stats) { perror(stats_path); return 1; }
    }
    struct timespec t0, t1, ts0, ts1;
    checkPbf(pbf_device_sync(0));
    clock_gettime(CLOCK_MONOTONIC, &t0);
    ts0 = t0;
    for (int s = 0; s < steps; s++) {
        if (emitter_total > 0 && fluids.frameCount > 0) fluids.stepSource();   // (frame 0: initialize() emitted the first layer)
        fluids.stepSimulate();
        const int done = s + 1;
        if (stats && (done % stats_every == 0 || done == steps)) {
            checkPbf(pbf_device_sync(0));
            clock_gettime(CLOCK_MONOTONIC, &ts1);
            const int span = done % stats_every == 0 ? stats_every : done % stats_every;
            writeStats(stats, fluids, fluids.frameCount, ((ts1.tv_sec - ts0.tv_sec) * 1e3 + 1e-6 * (ts1.tv_nsec - ts0.tv_nsec)) / span);
            clock_gettime(CLOCK_MONOTONIC, &ts0);
        }
        if (save && save_every > 0 && done % save_every == 0 && done != steps)
            checkPbf(pbf_checkpoint_save(h, save, (const float*)b.ptr(fluids.currentPos()), (const float*)b.ptr(fluids.currentVel()),
                                         (const uint32_t*)b.ptr(fluids.d_iid), fluids.m_nparticle, fluids.frameCount));
    }
    checkPbf(pbf_device_sync(0));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (stats) fclose(stats);
    if (save)
        checkPbf(pbf_checkpoint_save(h, save, (const float*)b.ptr(fluids.currentPos()), (const float*)b.ptr(fluids.currentVel()),
                                     (const uint32_t*)b.ptr(fluids.d_iid), fluids.m_nparticle, fluids.frameCount));
    const double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    pbf_stats st;
    checkPbf(pbf_get_stats(h, (const float*)b.ptr(fluids.currentPos()), (const float*)b.ptr(fluids.currentVel()), fluids.m_nparticle, &st));
    printf("{\"particles\": %d, \"steps\": %d, \"frame\": %d, \"moving\": %d, \"seconds\": %.6f, \"particle_steps_per_s\": %.1f, "
           "\"density_err_mean\": %.9g, \"density_err_max\": %.9g, \"kinetic_energy\": %.9g, \"max_speed\": %.9g, \"mean_z\": %.9g, "
           "\"launches\": %lld}\n",
           fluids.m_nparticle, steps, fluids.frameCount, moving, sec, fluids.m_nparticle * (double)steps / sec, st.density_err_mean, st.density_err_max,
           st.kinetic_energy, st.max_speed, st.mean_z, (long long)pbf_launch_count(h));
    if (dump) {
        const int n = fluids.m_nparticle;
        std::vector<float> pos((size_t)n * 3), vel((size_t)n * 3);
        std::vector<uint32_t> iid((size_t)n);
        b.getSubData(fluids.currentPos(), 0, (size_t)n * 12, pos.data());
        b.getSubData(fluids.currentVel(), 0, (size_t)n * 12, vel.data());
        b.getSubData(fluids.d_iid, 0, (size_t)n * 4, iid.data());
        FILE* f = fopen(dump, "wb");
        if (!f) { perror(dump); return 1; }
        fwrite(&n, sizeof(int), 1, f);
        fwrite(pos.data(), 4, pos.size(), f);
        fwrite(vel.data(), 4, vel.size(), f);
        fwrite(iid.data(), 4, iid.size(), f);
        fclose(f);
    }
    return 0;
}
