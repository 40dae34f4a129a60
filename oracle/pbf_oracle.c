/*
 * pbf_oracle.c — scalar CPU restatement of naeioi/PBF-CUDA's Simulator::step.
 * TEST INFRASTRUCTURE ONLY (see pbf_oracle.h). Pinned against the reference's own
 * CUDA build through tests/golden/ (made by tests/golden/make_golden.py on a B200).
 *
 * Compile:  gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared   (see oracle/Makefile)
 * -ffp-contract=off matters: every fused multiply-add the reference's device code
 * performs is written out with fmaf() below, nothing else may be contracted.
 * Host-side constants are computed the way the reference's host code computes them
 * (plain float / double arithmetic, glibc powf).
 *
 * Paths in the citations are relative to /root/reference/fluids/.
 */
#include "pbf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* helper.h:4-9. M_PI is redefined by the reference (the later #define wins). */
#define REF_PI 3.14159265359
#define REF_LIM_EPS 1e-3
#define REF_KERNAL_EPS 1e-4
#define REF_MAX_DP 0.1

struct orc_sim {
    orc_params p;
    float ulim[3], llim[3];
    int32_t dim[3];
    int64_t max_particles;
    int64_t ngrid;          /* capacity of the cell arrays */
    int nthreads;
    /* bound caller arrays (Simulator.h:50-51) */
    float *pos, *npos, *vel, *nvel;
    uint32_t* iid;
    int64_t n;
    /* scratch (Simulator.h:16-21) */
    uint32_t *grid_id, *grid_start, *grid_end;
    float *lambda, *pho, *tpos;
    /* sort scratch */
    uint32_t* perm;
    float* tmp3;
    uint32_t* tmp1;
    uint32_t* cell_count;
    float coef_corr;
};

/* ------------------------------------------------------------------------------------------
 * small pieces
 * ---------------------------------------------------------------------------------------- */

/* cvt.rzi.s32.f32: what `(int)(float)` compiles to on the device. NaN -> 0, saturating. */
static inline int32_t cvt_rzi(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

/* r2 as nvcc contracts `x*x + y*y + z*z` (helper.h:19 norm2, helper_math.h dot/length and the
 * open-coded sum at Simulator_kernel.cuh:89): fma(z,z, fma(x,x, y*y)). */
static inline float sumsq(float x, float y, float z) { return fmaf(z, z, fmaf(x, x, y * y)); }

typedef struct { float coef, h2; } poly6_t;
typedef struct { float h, coef; } spiky_t;

/* getPoly6::getPoly6 (Simulator.cu:77-83), host arithmetic. */
static poly6_t make_poly6(float h) {
    poly6_t k;
    k.h2 = h * h;
    float ih = 1.f / h;
    float ih3 = ih * ih * ih;
    float ih9 = ih3 * ih3 * ih3;
    k.coef = (float)((double)(315.f * ih9) / ((double)64.f * REF_PI));
    return k;
}
/* getPoly6::operator() (Simulator.cu:85-89). */
static inline float poly6_eval(poly6_t k, float r2) {
    if (r2 >= k.h2) return 0.f;
    float d = k.h2 - r2;
    return k.coef * d * d * d;
}
/* getSpikyGrad::getSpikyGrad (Simulator.cu:94-98). */
static spiky_t make_spiky(float h) {
    spiky_t k;
    k.h = h;
    float h6 = h * h;
    h6 = h6 * h6 * h6;
    k.coef = (float)((double)-45.f / (REF_PI * (double)h6));
    return k;
}
/* getSpikyGrad::operator() (Simulator.cu:101-106). r2 is shared with the caller's norm
 * (the compiler CSEs length(r)'s dot product with it). */
static inline void spiky_eval(spiky_t k, float rx, float ry, float rz, float r2, float g[3]) {
    float rlen = sqrtf(r2);
    if (rlen >= k.h || (double)rlen < REF_KERNAL_EPS) {
        g[0] = g[1] = g[2] = 0.f;
        return;
    }
    float d = k.h - rlen;
    float s = k.coef * d * d / rlen;
    g[0] = rx * s;
    g[1] = ry * s;
    g[2] = rz * s;
}

/* getGridxyz::operator() (Simulator.cu:30-35) == the coordinate part of getGridId (:66-71). */
static inline void cell_of(const orc_sim* s, const float* p, int32_t c[3]) {
    for (int a = 0; a < 3; a++) {
        float diff = p[a] - s->llim[a];
        c[a] = imin(imax(cvt_rzi(diff / s->p.h), 0), s->dim[a] - 1);
    }
}
/* xyzToId::operator() (Simulator.cu:45-53). */
static inline int32_t cell_id(const orc_sim* s, int32_t x, int32_t y, int32_t z) {
    return x * s->dim[1] * s->dim[2] + y * s->dim[2] + z;
}

/* DensityBoundary::densityAt (Simulator.cu:144-149): float in, double inside, float out. */
static inline float density_at(float h, float d) {
    if (d > h) return 0.f;
    if (d <= 0.f) return (float)(2 * REF_PI / 3);
    return (float)((2 * REF_PI / 3) * (double)(h - d) * (double)(h - d) * (double)(h + d));
}
/* DensityBoundary::operator() (Simulator.cu:152-160). */
static inline float boundary_density(const orc_sim* s, const float* p) {
    float h = s->p.h;
    return density_at(h, s->ulim[0] - p[0]) + density_at(h, p[0] - s->llim[0]) +
           density_at(h, s->ulim[1] - p[1]) + density_at(h, p[1] - s->llim[1]) +
           density_at(h, s->ulim[2] - p[2]) + density_at(h, p[2] - s->llim[2]);
}

/* ------------------------------------------------------------------------------------------
 * lifetime / params
 * ---------------------------------------------------------------------------------------- */

void orc_default_params(orc_params* p) { /* FluidSystem.cpp:15-25 */
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
}

static void compute_dim(orc_sim* s) { /* Simulator.cu:187-188 */
    for (int a = 0; a < 3; a++) {
        float diff = s->ulim[a] - s->llim[a];
        s->dim[a] = (int32_t)ceilf(diff / s->p.h);
    }
}

orc_sim* orc_create(const orc_params* p, const float ulim[3], const float llim[3],
                    int64_t max_particles) {
    orc_sim* s = (orc_sim*)calloc(1, sizeof(orc_sim));
    s->p = *p;
    memcpy(s->ulim, ulim, sizeof(float) * 3);
    memcpy(s->llim, llim, sizeof(float) * 3);
    s->max_particles = max_particles;
    compute_dim(s);
    /* Simulator.h:13-14 sizes the cell arrays as 4*(int)(dx*dy*dz) with d=(ulim-llim)/0.1;
     * never less than the actual number of cells (the reference would overflow there). */
    double dx = (ulim[0] - llim[0]) / 0.1, dy = (ulim[1] - llim[1]) / 0.1, dz = (ulim[2] - llim[2]) / 0.1;
    int64_t ngrid = 4 * (int64_t)(dx * dy * dz);
    int64_t cells = (int64_t)s->dim[0] * s->dim[1] * s->dim[2];
    if (ngrid < 2 * cells) ngrid = 2 * cells;
    s->ngrid = ngrid;
    s->grid_id = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)max_particles);
    s->grid_start = (uint32_t*)calloc((size_t)ngrid, sizeof(uint32_t));
    s->grid_end = (uint32_t*)calloc((size_t)ngrid, sizeof(uint32_t));
    s->cell_count = (uint32_t*)calloc((size_t)ngrid + 1, sizeof(uint32_t));
    s->lambda = (float*)calloc((size_t)max_particles, sizeof(float));
    s->pho = (float*)calloc((size_t)max_particles, sizeof(float));
    s->tpos = (float*)calloc((size_t)max_particles * 3, sizeof(float));
    s->perm = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)max_particles);
    s->tmp3 = (float*)malloc(sizeof(float) * 3 * (size_t)max_particles);
    s->tmp1 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)max_particles);
    s->nthreads = 1;
    return s;
}

void orc_destroy(orc_sim* s) {
    if (!s) return;
    free(s->grid_id); free(s->grid_start); free(s->grid_end); free(s->cell_count);
    free(s->lambda); free(s->pho); free(s->tpos); free(s->perm); free(s->tmp3); free(s->tmp1);
    free(s);
}

void orc_set_params(orc_sim* s, const orc_params* p) { s->p = *p; }
void orc_set_lim(orc_sim* s, const float ulim[3], const float llim[3]) {
    memcpy(s->ulim, ulim, sizeof(float) * 3);
    memcpy(s->llim, llim, sizeof(float) * 3);
}
void orc_set_threads(orc_sim* s, int nthreads) { s->nthreads = nthreads < 1 ? 1 : nthreads; }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_bind(orc_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n) {
    s->pos = pos; s->npos = npos; s->vel = vel; s->nvel = nvel; s->iid = iid; s->n = n;
}

/* ------------------------------------------------------------------------------------------
 * stages
 * ---------------------------------------------------------------------------------------- */

/* advect_kernel (Simulator_kernel.cuh:12-15): vel += dt*g ; npos = pos + dt*vel, both fma. */
void orc_advect(orc_sim* s) {
    const float dt = s->p.dt;
    const float g[3] = {0.f, 0.f, -s->p.g};
    const int64_t n = s->n;
#pragma omp parallel for schedule(static) num_threads(s->nthreads)
    for (int64_t i = 0; i < n; i++) {
        for (int a = 0; a < 3; a++) {
            float v = fmaf(dt, g[a], s->vel[3 * i + a]);
            s->vel[3 * i + a] = v;
            s->npos[3 * i + a] = fmaf(dt, v, s->pos[3 * i + a]);
        }
    }
}

static void permute3(orc_sim* s, float* a) {
    const int64_t n = s->n;
    for (int64_t i = 0; i < n; i++) {
        const float* src = a + 3 * (size_t)s->perm[i];
        s->tmp3[3 * i] = src[0]; s->tmp3[3 * i + 1] = src[1]; s->tmp3[3 * i + 2] = src[2];
    }
    memcpy(a, s->tmp3, sizeof(float) * 3 * (size_t)n);
}

/* buildGridHash (Simulator.cu:178-211) + computeGridRange (Simulator_kernel.cuh:21-50). */
void orc_build_grid(orc_sim* s) {
    const int64_t n = s->n;
    compute_dim(s);
    const int64_t cells = (int64_t)s->dim[0] * s->dim[1] * s->dim[2];
    /* getGridId on npos (Simulator.cu:190-193) */
#pragma omp parallel for schedule(static) num_threads(s->nthreads)
    for (int64_t i = 0; i < n; i++) {
        int32_t c[3];
        cell_of(s, s->npos + 3 * i, c);
        s->tmp1[i] = (uint32_t)cell_id(s, c[0], c[1], c[2]);
    }
    /* sort_by_key carrying (pos, vel, npos, nvel, iid) (Simulator.cu:196-198): radix sort,
     * stable. Restated as a stable counting sort. */
    memset(s->cell_count, 0, sizeof(uint32_t) * ((size_t)cells + 1));
    for (int64_t i = 0; i < n; i++) s->cell_count[s->tmp1[i] + 1]++;
    for (int64_t c = 0; c < cells; c++) s->cell_count[c + 1] += s->cell_count[c];
    for (int64_t i = 0; i < n; i++) {
        uint32_t dst = s->cell_count[s->tmp1[i]]++;
        s->perm[dst] = (uint32_t)i;
        s->grid_id[dst] = s->tmp1[i];
    }
    permute3(s, s->pos);
    permute3(s, s->vel);
    permute3(s, s->npos);
    permute3(s, s->nvel);
    for (int64_t i = 0; i < n; i++) s->tmp1[i] = s->iid[s->perm[i]];
    memcpy(s->iid, s->tmp1, sizeof(uint32_t) * (size_t)n);
    /* cudaMemset + computeGridRange (Simulator.cu:200-204) */
    memset(s->grid_start, 0, sizeof(uint32_t) * (size_t)cells);
    memset(s->grid_end, 0, sizeof(uint32_t) * (size_t)cells);
    for (int64_t i = 0; i < n; i++) {
        uint32_t cur = s->grid_id[i];
        uint32_t last = i == 0 ? (uint32_t)-1 : s->grid_id[i - 1];
        if (cur != last) {
            s->grid_start[cur] = (uint32_t)i;
            if (last != (uint32_t)-1) s->grid_end[last] = (uint32_t)i;
        }
        if (i == n - 1) s->grid_end[cur] = (uint32_t)n;
    }
}

/* computeLambda (Simulator_kernel.cuh:52-129). */
static void lambda_pass(orc_sim* s) {
    const int64_t n = s->n;
    const poly6_t poly6 = make_poly6(s->p.h);
    const spiky_t spiky = make_spiky(s->p.h);
    const float pho0 = s->p.pho0;
#pragma omp parallel for schedule(dynamic, 256) num_threads(s->nthreads)
    for (int64_t i = 0; i < n; i++) {
        const float* cp = s->npos + 3 * i;
        int32_t ind[3];
        cell_of(s, cp, ind);
        float pho = 0.f, gradj_l2 = 0.f;
        float gi[3] = {0.f, 0.f, 0.f};
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    int x = ind[0] + dx, y = ind[1] + dy, z = ind[2] + dz;
                    if (x < 0 || x >= s->dim[0] || y < 0 || y >= s->dim[1] || z < 0 || z >= s->dim[2]) continue;
                    int32_t c = cell_id(s, x, y, z);
                    uint32_t start = s->grid_start[c], end = s->grid_end[c];
                    for (int64_t j = start; j < (int64_t)end; j++) {
                        const float* q = s->npos + 3 * j;
                        float ddx = cp[0] - q[0], ddy = cp[1] - q[1], ddz = cp[2] - q[2];
                        float r2 = sumsq(ddx, ddy, ddz);
                        pho += poly6_eval(poly6, r2);
                        float g[3];
                        spiky_eval(spiky, ddx, ddy, ddz, r2, g);
                        g[0] = g[0] / pho0; g[1] = g[1] / pho0; g[2] = g[2] / pho0;
                        gi[0] += g[0]; gi[1] += g[1]; gi[2] += g[2];
                        if (j != i) gradj_l2 += sumsq(g[0], g[1], g[2]);
                    }
                }
        /* pho += k_b * boundaryDensity(cpos) contracts to one fma (Simulator_kernel.cuh:119-120) */
        pho = fmaf(s->p.k_boundaryDensity, boundary_density(s, cp), pho);
        /* grad_l2 = gradj_l2 + gx*gx + gy*gy + gz*gz: three chained fma (:122) */
        float grad_l2 = fmaf(gi[2], gi[2], fmaf(gi[1], gi[1], fmaf(gi[0], gi[0], gradj_l2)));
        s->lambda[i] = -(pho / pho0 - 1.f) / (grad_l2 + s->p.lambda_eps);
        s->pho[i] = pho;
    }
}

/* computetpos (Simulator_kernel.cuh:131-194). */
static void tpos_pass(orc_sim* s) {
    const int64_t n = s->n;
    const poly6_t poly6 = make_poly6(s->p.h);
    const spiky_t spiky = make_spiky(s->p.h);
    const float pho0 = s->p.pho0, coef_corr = s->coef_corr, n_corr = s->p.n_corr;
    const float max_dp = (float)REF_MAX_DP; /* clamp3f takes float limits (helper.h:26) */
#pragma omp parallel for schedule(dynamic, 256) num_threads(s->nthreads)
    for (int64_t i = 0; i < n; i++) {
        const float* cp = s->npos + 3 * i;
        int32_t ind[3];
        cell_of(s, cp, ind);
        const float lambda = s->lambda[i];
        float d[3] = {0.f, 0.f, 0.f};
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    int x = ind[0] + dx, y = ind[1] + dy, z = ind[2] + dz;
                    if (x < 0 || x >= s->dim[0] || y < 0 || y >= s->dim[1] || z < 0 || z >= s->dim[2]) continue;
                    int32_t c = cell_id(s, x, y, z);
                    uint32_t start = s->grid_start[c], end = s->grid_end[c];
                    for (int64_t j = start; j < (int64_t)end; j++) {
                        if (j == i) continue;
                        const float* q = s->npos + 3 * j;
                        float px = cp[0] - q[0], py = cp[1] - q[1], pz = cp[2] - q[2];
                        float r2 = sumsq(px, py, pz);
                        float pw = powf(poly6_eval(poly6, r2), n_corr);
                        /* (lambda + lambdas[j] + corr), corr = coef_corr*pw: one fma (:165-166) */
                        float sc = fmaf(coef_corr, pw, lambda + s->lambda[j]);
                        float g[3];
                        spiky_eval(spiky, px, py, pz, r2, g);
                        d[0] = fmaf(sc, g[0], d[0]);
                        d[1] = fmaf(sc, g[1], d[1]);
                        d[2] = fmaf(sc, g[2], d[2]);
                    }
                }
        float q[3];
        for (int a = 0; a < 3; a++) {
            float v = d[a] / pho0;
            v = fmaxf(fminf(v, max_dp), -max_dp);      /* clamp3f, helper.h:21-28 */
            q[a] = cp[a] + v;
            /* box clamp in double (LIM_EPS is a double literal), Simulator_kernel.cuh:190-192 */
            double hi = (double)s->ulim[a] - REF_LIM_EPS, lo = (double)s->llim[a] + REF_LIM_EPS;
            s->tpos[3 * i + a] = (float)fmax(fmin((double)q[a], hi), lo);
        }
    }
}

/* The two halves of correctDensity, exported separately so that the slab-protocol tests can
 * refresh ghost lambdas between them (tests/_slab_cpu.py): Simulator.cu:222-233 and :235-248. */
void orc_lambda_pass(orc_sim* s) { lambda_pass(s); }
void orc_delta_p_pass(orc_sim* s) {
    poly6_t poly6 = make_poly6(s->p.h);
    s->coef_corr = -s->p.k_corr / powf(poly6_eval(poly6, s->p.delta_q * s->p.delta_q), s->p.n_corr); /* :235 */
    tpos_pass(s);
    memcpy(s->npos, s->tpos, sizeof(float) * 3 * (size_t)s->n); /* thrust::copy_n, :247-248 */
}

/* correctDensity (Simulator.cu:213-249): lambda pass, coef_corr, tpos pass, commit. */
void orc_correct_density(orc_sim* s) {
    orc_lambda_pass(s);
    orc_delta_p_pass(s);
}

/* h_updateVelocity (Simulator.cu:127-137). */
void orc_update_velocity(orc_sim* s) {
    const float inv_dt = 1.f / s->p.dt;
    const int64_t n3 = 3 * s->n;
#pragma omp parallel for schedule(static) num_threads(s->nthreads)
    for (int64_t k = 0; k < n3; k++) s->vel[k] = (s->npos[k] - s->pos[k]) * inv_dt;
}

/* computeXSPH (Simulator_kernel.cuh:196-239). */
void orc_correct_velocity(orc_sim* s) {
    const int64_t n = s->n;
    const poly6_t poly6 = make_poly6(s->p.h);
    const float c_xsph = s->p.c_XSPH;
#pragma omp parallel for schedule(dynamic, 256) num_threads(s->nthreads)
    for (int64_t i = 0; i < n; i++) {
        const float* cp = s->npos + 3 * i;
        const float* cv = s->vel + 3 * i;
        const float cpho = s->pho[i];
        int32_t ind[3];
        cell_of(s, cp, ind);
        float av[3] = {0.f, 0.f, 0.f};
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    int x = ind[0] + dx, y = ind[1] + dy, z = ind[2] + dz;
                    if (x < 0 || x >= s->dim[0] || y < 0 || y >= s->dim[1] || z < 0 || z >= s->dim[2]) continue;
                    int32_t c = cell_id(s, x, y, z);
                    int32_t start = (int32_t)s->grid_start[c], end = (int32_t)s->grid_end[c];
                    for (int32_t j = start; j < end; j++) {
                        const float* q = s->npos + 3 * (size_t)j;
                        const float* vj = s->vel + 3 * (size_t)j;
                        float px = cp[0] - q[0], py = cp[1] - q[1], pz = cp[2] - q[2];
                        float w = poly6_eval(poly6, sumsq(px, py, pz));
                        float den = cpho + s->pho[j];
                        for (int a = 0; a < 3; a++) {
                            float vp = vj[a] - cv[a];
                            av[a] += ((vp + vp) * w) / den;   /* 2.f*vp*poly6/(cpho+phos[j]), :230 */
                        }
                    }
                }
        for (int a = 0; a < 3; a++) s->nvel[3 * i + a] = fmaf(c_xsph, av[a], cv[a]); /* :235 */
    }
}

/* Simulator::step (Simulator.cpp:44-79), GL interop and Logger removed. */
void orc_step(orc_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n) {
    orc_bind(s, pos, npos, vel, nvel, iid, n);
    orc_advect(s);
    orc_build_grid(s);
    for (int i = 0; i < s->p.niter; i++) orc_correct_density(s);
    orc_update_velocity(s);
    orc_correct_velocity(s);
}

/* ------------------------------------------------------------------------------------------
 * accessors / derived quantities
 * ---------------------------------------------------------------------------------------- */
const uint32_t* orc_grid_id(const orc_sim* s) { return s->grid_id; }
const uint32_t* orc_grid_start(const orc_sim* s) { return s->grid_start; }
const uint32_t* orc_grid_end(const orc_sim* s) { return s->grid_end; }
const float* orc_lambda(const orc_sim* s) { return s->lambda; }
const float* orc_pho(const orc_sim* s) { return s->pho; }
const float* orc_tpos(const orc_sim* s) { return s->tpos; }
void orc_grid_dim(const orc_sim* s, int32_t dim[3]) { dim[0] = s->dim[0]; dim[1] = s->dim[1]; dim[2] = s->dim[2]; }
float orc_coef_corr(const orc_sim* s) { return s->coef_corr; }
float orc_poly6_coef(const orc_sim* s) { return make_poly6(s->p.h).coef; }
float orc_spiky_coef(const orc_sim* s) { return make_spiky(s->p.h).coef; }
float orc_poly6(const orc_sim* s, float r2) { return poly6_eval(make_poly6(s->p.h), r2); }

static void count_impl(const orc_sim* s, uint32_t* out, int in_range_only) {
    const int64_t n = s->n;
    const float h2 = s->p.h * s->p.h;
#pragma omp parallel for schedule(dynamic, 256) num_threads(s->nthreads)
    for (int64_t i = 0; i < n; i++) {
        const float* cp = s->npos + 3 * i;
        int32_t ind[3];
        cell_of(s, cp, ind);
        uint32_t cnt = 0;
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    int x = ind[0] + dx, y = ind[1] + dy, z = ind[2] + dz;
                    if (x < 0 || x >= s->dim[0] || y < 0 || y >= s->dim[1] || z < 0 || z >= s->dim[2]) continue;
                    int32_t c = cell_id(s, x, y, z);
                    for (int64_t j = s->grid_start[c]; j < (int64_t)s->grid_end[c]; j++) {
                        const float* q = s->npos + 3 * j;
                        float r2 = sumsq(cp[0] - q[0], cp[1] - q[1], cp[2] - q[2]);
                        if (!in_range_only || r2 < h2) cnt++;
                    }
                }
        out[i] = cnt;
    }
}
void orc_neighbor_count(const orc_sim* s, uint32_t* out) { count_impl(s, out, 1); }
void orc_candidate_count(const orc_sim* s, uint32_t* out) { count_impl(s, out, 0); }

void orc_lambda_allpairs(const orc_sim* s, float* lambda_out, float* pho_out, uint32_t* count_out) {
    const int64_t n = s->n;
    const poly6_t poly6 = make_poly6(s->p.h);
    const spiky_t spiky = make_spiky(s->p.h);
    const float pho0 = s->p.pho0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(s->nthreads)
    for (int64_t i = 0; i < n; i++) {
        const float* cp = s->npos + 3 * i;
        float pho = 0.f, gradj_l2 = 0.f, gi[3] = {0.f, 0.f, 0.f};
        uint32_t cnt = 0;
        for (int64_t j = 0; j < n; j++) {
            const float* q = s->npos + 3 * j;
            float ddx = cp[0] - q[0], ddy = cp[1] - q[1], ddz = cp[2] - q[2];
            float r2 = sumsq(ddx, ddy, ddz);
            if (r2 < poly6.h2) cnt++;
            pho += poly6_eval(poly6, r2);
            float g[3];
            spiky_eval(spiky, ddx, ddy, ddz, r2, g);
            g[0] = g[0] / pho0; g[1] = g[1] / pho0; g[2] = g[2] / pho0;
            gi[0] += g[0]; gi[1] += g[1]; gi[2] += g[2];
            if (j != i) gradj_l2 += sumsq(g[0], g[1], g[2]);
        }
        pho = fmaf(s->p.k_boundaryDensity, boundary_density(s, cp), pho);
        float grad_l2 = fmaf(gi[2], gi[2], fmaf(gi[1], gi[1], fmaf(gi[0], gi[0], gradj_l2)));
        lambda_out[i] = -(pho / pho0 - 1.f) / (grad_l2 + s->p.lambda_eps);
        pho_out[i] = pho;
        count_out[i] = cnt;
    }
}

/* ------------------------------------------------------------------------------------------
 * scenes
 * ---------------------------------------------------------------------------------------- */

/* MSVC rand(): the reference ran on Windows (fluids.vcxproj); RAND_MAX = 32767. */
static inline int msvc_rand(uint32_t* state) {
    *state = *state * 214013u + 2531011u;
    return (int)((*state >> 16) & 0x7fff);
}
#define MSVC_RAND_MAX 32767

/* DoubleDamSource::generate_cube (DoubleDamSource.cpp:5-21) with d from the ctor
 * (DoubleDamSource.h:13-22); identical loop in FixedCubeSource::initialize. */
int64_t orc_scene_cube(const float ulim[3], const float llim[3], const int32_t ns[3],
                       uint32_t* rng_state, uint32_t first_iid,
                       float* pos, float* vel, uint32_t* iid) {
    float d[3];
    for (int a = 0; a < 3; a++) {
        d[a] = ulim[a] - llim[a];
        d[a] /= (float)ns[a];
    }
    float sx = llim[0] + d[0] / 2, sy = llim[1] + d[1] / 2, sz = llim[2] + d[2] / 2;
    int64_t count = 0;
    float x = sx;
    for (int i = 0; i < ns[0]; i++, x += d[0]) {
        float y = sy;
        for (int j = 0; j < ns[1]; j++, y += d[1]) {
            float z = sz;
            for (int k = 0; k < ns[2]; k++, z += d[2], count++) {
                float r1 = 1.f * msvc_rand(rng_state) / MSVC_RAND_MAX;
                float r2 = 1.f * msvc_rand(rng_state) / MSVC_RAND_MAX;
                float r3 = 1.f * msvc_rand(rng_state) / MSVC_RAND_MAX;
                pos[3 * count + 0] = x + 0.1f * (sx * r1);
                pos[3 * count + 1] = y + 0.1f * (sy * r2);
                pos[3 * count + 2] = z + 0.1f * (sz * r3);
                vel[3 * count + 0] = vel[3 * count + 1] = vel[3 * count + 2] = 0.f;
                iid[count] = first_iid + (uint32_t)count;
            }
        }
    }
    return count;
}

/* FluidSystem.cpp:34-35,55-61 + DoubleDamSource::initialize (DoubleDamSource.cpp:23-31). */
int64_t orc_scene_double_dam_reference(float* pos, float* vel, uint32_t* iid, float ulim[3], float llim[3]) {
    ulim[0] = 2.f; ulim[1] = 2.f; ulim[2] = 4.f;
    llim[0] = -2.f; llim[1] = -2.f; llim[2] = 0.f;
    float dd = 1.f / 20;
    float d1 = dd * 20, d2 = dd * 20, d3 = dd * 40;
    float u1[3] = {-1.8f, 1.8f, 3.8f}, l1[3] = {-1.8f + d1, 1.8f - d2, 3.8f - d3};
    float u2[3] = {1.8f - d1, -1.8f + d2, 3.8f}, l2[3] = {1.8f, -1.8f, 3.8f - d3};
    int32_t ns[3] = {20, 20, 40};
    uint32_t rng = 27; /* srand(27) */
    int64_t c1 = orc_scene_cube(u1, l1, ns, &rng, 0, pos, vel, iid);
    int64_t c2 = orc_scene_cube(u2, l2, ns, &rng, (uint32_t)c1, pos + 3 * c1, vel + 3 * c1, iid + c1);
    return c1 + c2;
}

/* Counter-based jitter for the scalable scenes (not in the reference, whose
 * jitter is scaled by the block's start coordinate and degenerates for large boxes). */
static inline uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
void orc_scene_block(const float origin[3], const int32_t n[3], float spacing, uint32_t seed,
                     uint32_t first_iid, float* pos, float* vel, uint32_t* iid) {
    const float jit = 0.2f * spacing;
    for (int32_t ix = 0; ix < n[0]; ix++)
        for (int32_t iy = 0; iy < n[1]; iy++)
            for (int32_t iz = 0; iz < n[2]; iz++) {
                uint32_t local = ((uint32_t)ix * (uint32_t)n[1] + (uint32_t)iy) * (uint32_t)n[2] + (uint32_t)iz;
                uint32_t id = first_iid + local;
                int32_t idx[3] = {ix, iy, iz};
                for (int a = 0; a < 3; a++) {
                    uint32_t hsh = hash32(seed * 0x9E3779B9u + hash32(id * 3u + (uint32_t)a + 0x7F4A7C15u));
                    float u = (float)(hsh >> 8) * (1.0f / 16777216.0f);
                    float base = origin[a] + spacing * ((float)idx[a] + 0.5f);
                    pos[3 * (size_t)local + a] = base + jit * u;
                    vel[3 * (size_t)local + a] = 0.f;
                }
                iid[local] = id;
            }
}

/* FluidSystem.cpp:104-110. */
void orc_wall_lim(const float ulim0[3], const float llim0[3], const float a_ulim[3],
                  const float a_llim[3], float w, int frame, int start_frame,
                  float ulim[3], float llim[3]) {
    float t = w * (float)(frame - start_frame);
    float phi = (float)sin((double)t);
    for (int a = 0; a < 3; a++) {
        ulim[a] = ulim0[a] + a_ulim[a] * phi;
        llim[a] = llim0[a] + a_llim[a] * phi;
    }
}

void orc_stats(const float* pho, const float* npos, const float* nvel, int64_t n, float pho0, double out[5]) {
    double err_sum = 0, err_max = -1e300, ke = 0, vmax = 0, zsum = 0;
    for (int64_t i = 0; i < n; i++) {
        double c = (double)pho[i] / (double)pho0 - 1.0;
        err_sum += fabs(c);
        if (c > err_max) err_max = c;
        double v2 = (double)nvel[3 * i] * nvel[3 * i] + (double)nvel[3 * i + 1] * nvel[3 * i + 1] +
                    (double)nvel[3 * i + 2] * nvel[3 * i + 2];
        ke += 0.5 * v2;
        if (v2 > vmax) vmax = v2;
        zsum += npos[3 * i + 2];
    }
    out[0] = n ? err_sum / (double)n : 0;
    out[1] = n ? err_max : 0;
    out[2] = ke;
    out[3] = sqrt(vmax);
    out[4] = n ? zsum / (double)n : 0;
}
