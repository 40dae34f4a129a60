// Simulator.h — drop-in C++ mirror of the reference's Simulator (fluids/Simulator.h:7-61,
// fluids/Simulator.cpp:10-136) on top of the C-ABI (include/pbf.h).
//
// Same constructor, same five public members, same meaning:
//     Simulator(const GUIParams&, float3 ulim, float3 llim)
//     void step(uint d_pos, uint d_npos, uint d_vel, uint d_nvel, uint d_iid, int nparticle)
//     void loadParams(); void saveParams(); void setLim(const float3&, const float3&)
// The reference's `uint` arguments are OpenGL buffer names that step() maps to device pointers
// with cudaGraphicsGLRegisterBuffer/Map/GetMappedPointer on every call (Simulator.cpp:20-36).
// Headless there is no GL: the names are handles of a small buffer registry (DeviceBuffers) that
// resolve to device pointers; a GL viewer would register its mapped VBO pointers under the VBO
// names instead. Errors keep the reference's convention (checkCudaErrors,
// common/cuda_inc/helper_cuda.h:999-1011): print "file:line message" and exit(EXIT_FAILURE).
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include <map>

#include "../../include/pbf.h"
#include "GUIParams.h"
#include "helper.h"

#define checkPbf(call)                                                                     \
    do {                                                                                   \
        int rc__ = (call);                                                                 \
        if (rc__ != PBF_OK) {                                                              \
            fprintf(stderr, "PBF error at %s:%d code=%d \"%s\" : %s\n", __FILE__, __LINE__, rc__, #call, pbf_last_error()); \
            exit(EXIT_FAILURE);                                                            \
        }                                                                                  \
    } while (0)

// uint name -> device pointer; replaces glGenBuffers/glBufferData + the per-step GL interop.
class DeviceBuffers {
public:
    static DeviceBuffers& getInstance() { static DeviceBuffers inst; return inst; }
    uint create(size_t bytes, int device = 0) {   // glGenBuffers + glBufferData(NULL)
        void* p = 0;
        checkPbf(pbf_device_alloc(device, (int64_t)bytes, &p));
        uint name = ++m_next;
        m_ptr[name] = p;
        return name;
    }
    uint adopt(void* device_ptr) {                // e.g. a mapped GL VBO pointer owned by a viewer
        uint name = ++m_next;
        m_ptr[name] = device_ptr;
        m_foreign[name] = true;
        return name;
    }
    void* ptr(uint name) const {
        std::map<uint, void*>::const_iterator it = m_ptr.find(name);
        if (it == m_ptr.end()) { fprintf(stderr, "DeviceBuffers: unknown buffer name %u\n", name); exit(EXIT_FAILURE); }
        return it->second;
    }
    void subData(uint name, size_t offset, size_t bytes, const void* host) {   // glBufferSubData
        checkPbf(pbf_copy_h2d((char*)ptr(name) + offset, host, (int64_t)bytes));
    }
    void getSubData(uint name, size_t offset, size_t bytes, void* host) const {
        checkPbf(pbf_copy_d2h(host, (const char*)ptr(name) + offset, (int64_t)bytes));
    }
    void destroy(uint name, int device = 0) {
        if (!m_foreign.count(name)) pbf_device_free(device, ptr(name));
        m_ptr.erase(name); m_foreign.erase(name);
    }
private:
    DeviceBuffers() : m_next(0) {}
    std::map<uint, void*> m_ptr;
    std::map<uint, bool> m_foreign;
    uint m_next;
};

class Simulator {
public:
    // As in the reference the `params` argument is not read: the constructor calls loadParams(),
    // which pulls from the GUIParams singleton (Simulator.h:10-12, Simulator.cpp:103).
    Simulator(const GUIParams& /*params*/, float3 ulim, float3 llim, int max_particles = MAX_PARTICLE_NUM, int device = 0)
        : m_sim(0), m_ulim(ulim), m_llim(llim) {
        pbf_params p = GUIParams::getInstance().toC();
        const float u[3] = {ulim.x, ulim.y, ulim.z}, l[3] = {llim.x, llim.y, llim.z};
        checkPbf(pbf_create(&p, u, l, max_particles, device, &m_sim));
        loadParams();
    }
    ~Simulator() { pbf_destroy(m_sim); }

    void step(uint d_pos, uint d_npos, uint d_vel, uint d_nvel, uint d_iid, int nparticle) {
        DeviceBuffers& b = DeviceBuffers::getInstance();
        checkPbf(pbf_step(m_sim, (float*)b.ptr(d_pos), (float*)b.ptr(d_npos), (float*)b.ptr(d_vel), (float*)b.ptr(d_nvel),
                          (uint32_t*)b.ptr(d_iid), nparticle, /*stream*/ 0));
    }
    void loadParams() {                              // Simulator.cpp:101-115
        pbf_params p = GUIParams::getInstance().toC();
        checkPbf(pbf_set_params(m_sim, &p));
    }
    void saveParams() {                              // Simulator.cpp:117-130
        pbf_params p;
        checkPbf(pbf_get_params(m_sim, &p));
        GUIParams::getInstance().fromC(p);
    }
    void setLim(const float3& ulim, const float3& llim) {   // Simulator.cpp:132-136
        m_ulim = ulim; m_llim = llim;
        const float u[3] = {ulim.x, ulim.y, ulim.z}, l[3] = {llim.x, llim.y, llim.z};
        checkPbf(pbf_set_lim(m_sim, u, l));
    }
    pbf_sim* handle() { return m_sim; }              // for read-backs / statistics through the C-ABI
private:
    Simulator(const Simulator&);
    Simulator& operator=(const Simulator&);
    pbf_sim* m_sim;
    float3 m_ulim, m_llim;
};
