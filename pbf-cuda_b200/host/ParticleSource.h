// ParticleSource.h — mirror of the reference's scene interface and its two generators
// (fluids/ParticleSource.h:4-17, DoubleDamSource.{h,cpp}, FixedCubeSource.{h,cpp}): same class
// names, constructors and initialize/update/reset(uint pos, uint vel, uint iid, int max) -> count.
// Generation is pbf_scene_cube of the C-ABI (the reference's loop, MSVC rand() LCG, srand(27));
// the upload is DeviceBuffers::subData where the reference calls glBufferSubData.
#pragma once
#include <vector>

#include "Simulator.h"

class ParticleSource {
public:
    ParticleSource() {}
    virtual ~ParticleSource() {}
    /* Params: buffer names of pos, vel, iid + max_nparticle; return number of particles */
    virtual int initialize(uint, uint, uint, int) = 0;
    virtual int update(uint, uint, uint, int) = 0;
    virtual int reset(uint, uint, uint, int) = 0;
};

namespace pbf_host {
struct Block { float3 ulim, llim; int3 ns; };
inline int generate_and_upload(const std::vector<Block>& blocks, uint pos, uint vel, uint iid, int max_nparticle) {
    std::vector<float> h_pos((size_t)max_nparticle * 3), h_vel((size_t)max_nparticle * 3);
    std::vector<uint32_t> h_iid((size_t)max_nparticle);
    uint32_t rng = 27;  // srand(27), DoubleDamSource.cpp:25 / FixedCubeSource.cpp:8
    int64_t count = 0;
    for (size_t b = 0; b < blocks.size(); b++) {
        const float u[3] = {blocks[b].ulim.x, blocks[b].ulim.y, blocks[b].ulim.z};
        const float l[3] = {blocks[b].llim.x, blocks[b].llim.y, blocks[b].llim.z};
        const int32_t ns[3] = {blocks[b].ns.x, blocks[b].ns.y, blocks[b].ns.z};
        int64_t c = 0;
        checkPbf(pbf_scene_cube(u, l, ns, &rng, (uint32_t)count, h_pos.data() + 3 * count, h_vel.data() + 3 * count,
                                h_iid.data() + count, max_nparticle - count, &c));
        count += c;
    }
    DeviceBuffers& d = DeviceBuffers::getInstance();
    d.subData(pos, 0, (size_t)count * 12, h_pos.data());
    d.subData(vel, 0, (size_t)count * 12, h_vel.data());
    d.subData(iid, 0, (size_t)count * 4, h_iid.data());
    return (int)count;
}
}  // namespace pbf_host

class FixedCubeSource : public ParticleSource {   // fluids/FixedCubeSource.h:9-16
public:
    FixedCubeSource(float3 ulim, float3 llim, int3 ns) : m_count(0) { pbf_host::Block b = {ulim, llim, ns}; m_blocks.push_back(b); }
    int initialize(uint pos, uint vel, uint iid, int max_nparticle) { return m_count = pbf_host::generate_and_upload(m_blocks, pos, vel, iid, max_nparticle); }
    int update(uint, uint, uint, int) { return m_count; }
    int reset(uint pos, uint vel, uint iid, int max_nparticle) { return initialize(pos, vel, iid, max_nparticle); }
private:
    std::vector<pbf_host::Block> m_blocks;
    int m_count;
};

class DoubleDamSource : public ParticleSource {   // fluids/DoubleDamSource.h:8-20
public:
    DoubleDamSource(float3 ulim1, float3 llim1, int3 ns1, float3 ulim2, float3 llim2, int3 ns2) : m_count(0) {
        pbf_host::Block a = {ulim1, llim1, ns1}, b = {ulim2, llim2, ns2};
        m_blocks.push_back(a); m_blocks.push_back(b);
    }
    int initialize(uint pos, uint vel, uint iid, int max_nparticle) { return m_count = pbf_host::generate_and_upload(m_blocks, pos, vel, iid, max_nparticle); }
    int update(uint, uint, uint, int) { return m_count; }
    int reset(uint pos, uint vel, uint iid, int max_nparticle) { return initialize(pos, vel, iid, max_nparticle); }
private:
    std::vector<pbf_host::Block> m_blocks;
    int m_count;
};

// An emitter: the use the reference's interface was made for but never shipped — FluidSystem::stepSource()
// (fluids/FluidSystem.cpp:92-97, defined but never called) hands update() the CURRENT ping-pong buffers and
// takes the returned count, while both shipped sources return a constant (DoubleDamSource.cpp:43-45).
// EmitterSource appends one lattice layer (ns.x = 1 layers of ny x nz particles, spacing `d`, at `origin`,
// all with velocity `v0`) every `period` calls, at the end of the caller's buffers — the solver accepts any
// order and permutes iid along — until `total` particles exist. iid continues the running count.
// Deterministic (no jitter): a run is a function of (origin, ny, nz, d, v0, period, total) and the call count.
class EmitterSource : public ParticleSource {
public:
    EmitterSource(float3 origin, int ny, int nz, float d, float3 v0, int period, int total)
        : m_origin(origin), m_ny(ny), m_nz(nz), m_d(d), m_v0(v0), m_period(period < 1 ? 1 : period), m_total(total),
          m_count(0), m_calls(0) {}
    int initialize(uint pos, uint vel, uint iid, int max_nparticle) {
        m_count = 0;
        m_calls = 0;
        return update(pos, vel, iid, max_nparticle);
    }
    int update(uint pos, uint vel, uint iid, int max_nparticle) {
        const int layer = m_ny * m_nz;
        const int limit = m_total < max_nparticle ? m_total : max_nparticle;
        if (m_calls % m_period == 0 && m_count + layer <= limit) {
            std::vector<float> h_pos((size_t)layer * 3), h_vel((size_t)layer * 3);
            std::vector<uint32_t> h_iid((size_t)layer);
            for (int j = 0; j < m_ny; j++)
                for (int k = 0; k < m_nz; k++) {
                    const int e = j * m_nz + k;
                    h_pos[3 * e] = m_origin.x;
                    h_pos[3 * e + 1] = m_origin.y + m_d * (float)j;
                    h_pos[3 * e + 2] = m_origin.z + m_d * (float)k;
                    h_vel[3 * e] = m_v0.x; h_vel[3 * e + 1] = m_v0.y; h_vel[3 * e + 2] = m_v0.z;
                    h_iid[e] = (uint32_t)(m_count + e);
                }
            DeviceBuffers& b = DeviceBuffers::getInstance();
            b.subData(pos, (size_t)m_count * 12, (size_t)layer * 12, h_pos.data());
            b.subData(vel, (size_t)m_count * 12, (size_t)layer * 12, h_vel.data());
            b.subData(iid, (size_t)m_count * 4, (size_t)layer * 4, h_iid.data());
            m_count += layer;
        }
        m_calls++;
        return m_count;
    }
    int reset(uint pos, uint vel, uint iid, int max_nparticle) { return initialize(pos, vel, iid, max_nparticle); }
    // resume from a state file: `count` particles exist after `calls` update() calls
    void restore(int count, int calls) { m_count = count; m_calls = calls; }
private:
    float3 m_origin;
    int m_ny, m_nz;
    float m_d;
    float3 m_v0;
    int m_period, m_total, m_count, m_calls;
};
