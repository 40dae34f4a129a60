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
