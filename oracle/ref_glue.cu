/*
 * ref_glue.cu — headless driver for the reference's OWN Simulator.cu. TEST INFRASTRUCTURE ONLY.
 *
 * The reference's Simulator.cu / Simulator_kernel.cuh / Simulator.h / helper.h /
 * GUIParams.{h,cpp} are compiled UNCHANGED from where they lie under /root/reference
 * (see oracle/Makefile, target _ref); nothing of the reference is copied into this repo.
 * This file only supplies what fluids/Simulator.cpp would have supplied had it not been
 * tied to OpenGL (it includes <glad\glad.h> and cuda_gl_interop.h, Simulator.cpp:6-8):
 *   - Simulator::loadParams / saveParams / setLim   (restated from Simulator.cpp:101-136)
 *   - the stage sequence of Simulator::step with its cudaDeviceSynchronize fences
 *     (Simulator.cpp:44-79) on raw device pointers instead of mapped GL buffers
 * and a C interface for ctypes. The private stage methods are reached with the usual
 * `#define private public` (class layout is unaffected).
 *
 * Scratch capacity: the reference hard-codes MAX_PARTICLE_NUM = 130000 (helper.h:10) and a
 * cell capacity derived from the constructor box (Simulator.h:13-14). For larger scenes the
 * glue frees and re-allocates those six scratch arrays after construction — no source patch.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <helper_math.h>
#include <helper_cuda.h>

#define private public
#include "Simulator.h"
#undef private

/* ---- restated from fluids/Simulator.cpp:101-136 (the file itself needs OpenGL) ---- */
void Simulator::loadParams() {
    const GUIParams& params = GUIParams::getInstance();
    m_dt = params.dt;
    m_gravity = params.g;
    m_h = params.h;
    m_pho0 = params.pho0;
    m_lambda_eps = params.lambda_eps;
    m_delta_q = params.delta_q;
    m_k_corr = params.k_corr;
    m_n_corr = params.n_corr;
    m_k_boundaryDensity = params.k_boundaryDensity;
    m_c_XSPH = params.c_XSPH;
    m_niter = params.niter;
}
void Simulator::saveParams() {
    GUIParams& params = GUIParams::getInstance();
    params.dt = m_dt;
    params.g = m_gravity;
    params.h = m_h;
    params.pho0 = m_pho0;
    params.lambda_eps = m_lambda_eps;
    params.delta_q = m_delta_q;
    params.k_corr = m_k_corr;
    params.n_corr = m_n_corr;
    params.k_boundaryDensity = m_k_boundaryDensity;
    params.c_XSPH = m_c_XSPH;
    params.niter = m_niter;
}
void Simulator::setLim(const float3& ulim, const float3& llim) {
    m_llim = llim;
    m_ulim = ulim;
}

struct ref_params {  /* same layout as pbf_params / orc_params */
    int32_t niter;
    float pho0, g, h, dt, lambda_eps, delta_q, k_corr, n_corr, k_boundaryDensity, c_XSPH;
};

struct ref_handle {
    Simulator* sim;
    int64_t max_particles;
    int64_t ngrid;
};

static void write_singleton(const ref_params* p) {
    GUIParams& g = GUIParams::getInstance();
    g.niter = p->niter; g.pho0 = p->pho0; g.g = p->g; g.h = p->h; g.dt = p->dt;
    g.lambda_eps = p->lambda_eps; g.delta_q = p->delta_q; g.k_corr = p->k_corr;
    g.n_corr = p->n_corr; g.k_boundaryDensity = p->k_boundaryDensity; g.c_XSPH = p->c_XSPH;
}

extern "C" {

__attribute__((visibility("default")))
void* ref_create(const ref_params* p, const float* ulim, const float* llim, int64_t max_particles) {
    write_singleton(p);
    /* construct on the reference's own box so that its constructor memset (Simulator.h:24-25,
     * MAX_PARTICLE_NUM*4 bytes into ngrid*4-byte arrays) stays in bounds, then re-provision */
    Simulator* sim = new Simulator(GUIParams::getInstance(), make_float3(2.f, 2.f, 4.f), make_float3(-2.f, -2.f, 0.f));
    ref_handle* hd = new ref_handle;
    hd->sim = sim;
    hd->max_particles = max_particles < MAX_PARTICLE_NUM ? MAX_PARTICLE_NUM : max_particles;
    float3 u = make_float3(ulim[0], ulim[1], ulim[2]), l = make_float3(llim[0], llim[1], llim[2]);
    float3 d = (u - l) / 0.1;
    int64_t ngrid = 4 * (int64_t)((double)d.x * d.y * d.z);      /* Simulator.h:13-14 */
    if (ngrid < 256000) ngrid = 256000;
    hd->ngrid = ngrid;
    cudaFree(sim->dc_gridId); cudaFree(sim->dc_gridStart); cudaFree(sim->dc_gridEnd);
    cudaFree(sim->dc_lambda); cudaFree(sim->dc_pho); cudaFree(sim->dc_tpos);
    size_t nid = (size_t)(hd->max_particles > ngrid ? hd->max_particles : ngrid);
    cudaError_t e = cudaSuccess;
    e = cudaMalloc(&sim->dc_gridId, sizeof(uint) * nid);              if (e) goto fail;
    e = cudaMalloc(&sim->dc_gridStart, sizeof(uint) * (size_t)ngrid); if (e) goto fail;
    e = cudaMalloc(&sim->dc_gridEnd, sizeof(uint) * (size_t)ngrid);   if (e) goto fail;
    e = cudaMalloc(&sim->dc_lambda, sizeof(float) * (size_t)hd->max_particles);  if (e) goto fail;
    e = cudaMalloc(&sim->dc_pho, sizeof(float) * (size_t)hd->max_particles);     if (e) goto fail;
    e = cudaMalloc(&sim->dc_tpos, sizeof(float3) * (size_t)hd->max_particles);   if (e) goto fail;
    cudaMemset(sim->dc_gridStart, 0, sizeof(uint) * (size_t)ngrid);
    cudaMemset(sim->dc_gridEnd, 0, sizeof(uint) * (size_t)ngrid);
    sim->setLim(u, l);
    return hd;
fail:
    fprintf(stderr, "ref_create: %s\n", cudaGetErrorString(e));
    return nullptr;
}

__attribute__((visibility("default")))
void ref_destroy(void* h) {
    ref_handle* hd = (ref_handle*)h;
    delete hd->sim;
    delete hd;
}

__attribute__((visibility("default")))
void ref_set_params(void* h, const ref_params* p) {
    write_singleton(p);
    ((ref_handle*)h)->sim->loadParams();   /* FluidSystem.cpp:102 */
}

__attribute__((visibility("default")))
void ref_set_lim(void* h, const float* ulim, const float* llim) {
    ((ref_handle*)h)->sim->setLim(make_float3(ulim[0], ulim[1], ulim[2]), make_float3(llim[0], llim[1], llim[2]));
}

__attribute__((visibility("default")))
void ref_bind(void* h, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n) {
    Simulator* s = ((ref_handle*)h)->sim;
    s->m_nparticle = (int)n;
    s->dc_pos = (float3*)pos; s->dc_npos = (float3*)npos;
    s->dc_vel = (float3*)vel; s->dc_nvel = (float3*)nvel;
    s->dc_iid = iid;
}

/* stage: 0 advect, 1 buildGridHash, 2 correctDensity (one iteration), 3 updateVelocity, 4 correctVelocity */
__attribute__((visibility("default")))
int ref_stage(void* h, int stage) {
    Simulator* s = ((ref_handle*)h)->sim;
    switch (stage) {
        case 0: s->advect(); break;
        case 1: s->buildGridHash(); break;
        case 2: s->correctDensity(); break;
        case 3: s->updateVelocity(); break;
        case 4: s->correctVelocity(); break;
        default: return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "ref_stage %d: %s\n", stage, cudaGetErrorString(e)); return 2; }
    return 0;
}

/* The stage sequence of Simulator::step (Simulator.cpp:44-79): same calls, same fences. */
__attribute__((visibility("default")))
int ref_step(void* h, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n) {
    ref_bind(h, pos, npos, vel, nvel, iid, n);
    Simulator* s = ((ref_handle*)h)->sim;
    cudaDeviceSynchronize();
    s->advect();
    cudaDeviceSynchronize();
    s->buildGridHash();
    cudaDeviceSynchronize();
    for (uint i = 0; i < (uint)s->m_niter; i++) s->correctDensity();
    cudaDeviceSynchronize();
    s->updateVelocity();
    cudaDeviceSynchronize();
    s->correctVelocity();
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : 2;
}

/* what: 0 gridId[n], 1 gridStart[cells], 2 gridEnd[cells], 3 lambda[n], 4 pho[n], 5 tpos[3n] */
__attribute__((visibility("default")))
int ref_read(void* h, int what, void* dst, int64_t count) {
    Simulator* s = ((ref_handle*)h)->sim;
    const void* src = nullptr;
    size_t elt = 4;
    switch (what) {
        case 0: src = s->dc_gridId; break;
        case 1: src = s->dc_gridStart; break;
        case 2: src = s->dc_gridEnd; break;
        case 3: src = s->dc_lambda; break;
        case 4: src = s->dc_pho; break;
        case 5: src = s->dc_tpos; elt = 12; break;
        default: return 1;
    }
    return cudaMemcpy(dst, src, elt * (size_t)count, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}

__attribute__((visibility("default")))
void ref_grid_dim(void* h, int32_t* dim) {
    Simulator* s = ((ref_handle*)h)->sim;
    dim[0] = s->m_gridHashDim.x; dim[1] = s->m_gridHashDim.y; dim[2] = s->m_gridHashDim.z;
}

__attribute__((visibility("default")))
float ref_coef_corr(void* h) { return ((ref_handle*)h)->sim->m_coef_corr; }

}  /* extern "C" */
