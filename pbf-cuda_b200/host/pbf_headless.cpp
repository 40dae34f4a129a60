// pbf_headless.cpp — headless C++ harness: the reference's application loop without the viewer.
//
// Restates what FluidSystem does around the solver (fluids/FluidSystem.cpp:8-120, fluids/main.cpp):
// default parameters, the box, the double-dam scene, five particle buffers sized MAX_PARTICLE_NUM,
// initSource(), then `stepSimulate()` in a loop — loadParams(), optional sweeping wall
// (ulim + A_ulim*sin(w*(frame-start))), step() with the ping-pong roles, frameCount++ — written
// against the shim headers so it reads like the reference's own code. Calls CUDA only through the
// C-ABI (libpbf_b200.so); compiles with plain g++.
//
//   pbf_headless [steps=100] [moving=0] [dump.bin]
// prints per-run statistics and, with a third argument, dumps the final
// (npos, nvel, iid) so tests can compare it with the Python path bit for bit.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

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

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 100;
    const int moving = argc > 2 ? atoi(argv[2]) : 0;
    FluidSystemHeadless fluids;
    fluids.initSource();
    fluids.moving = moving != 0;
    pbf_sim* h = fluids.m_simulator->handle();
    struct timespec t0, t1;
    checkPbf(pbf_device_sync(0));
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int s = 0; s < steps; s++) fluids.stepSimulate();
    checkPbf(pbf_device_sync(0));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    DeviceBuffers& b = DeviceBuffers::getInstance();
    pbf_stats st;
    checkPbf(pbf_get_stats(h, (const float*)b.ptr(fluids.currentPos()), (const float*)b.ptr(fluids.currentVel()), fluids.m_nparticle, &st));
    printf("{\"particles\": %d, \"steps\": %d, \"moving\": %d, \"seconds\": %.6f, \"particle_steps_per_s\": %.1f, "
           "\"density_err_mean\": %.9g, \"density_err_max\": %.9g, \"kinetic_energy\": %.9g, \"max_speed\": %.9g, \"mean_z\": %.9g, "
           "\"launches\": %lld}\n",
           fluids.m_nparticle, steps, moving, sec, fluids.m_nparticle * (double)steps / sec, st.density_err_mean, st.density_err_max,
           st.kinetic_energy, st.max_speed, st.mean_z, (long long)pbf_launch_count(h));
    if (argc > 3) {
        const int n = fluids.m_nparticle;
        std::vector<float> pos((size_t)n * 3), vel((size_t)n * 3);
        std::vector<uint32_t> iid((size_t)n);
        b.getSubData(fluids.currentPos(), 0, (size_t)n * 12, pos.data());
        b.getSubData(fluids.currentVel(), 0, (size_t)n * 12, vel.data());
        b.getSubData(fluids.d_iid, 0, (size_t)n * 4, iid.data());
        FILE* f = fopen(argv[3], "wb");
        if (!f) { perror(argv[3]); return 1; }
        fwrite(&n, sizeof(int), 1, f);
        fwrite(pos.data(), 4, pos.size(), f);
        fwrite(vel.data(), 4, vel.size(), f);
        fwrite(iid.data(), 4, iid.size(), f);
        fclose(f);
    }
    return 0;
}
