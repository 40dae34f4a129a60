/*
 * pbf_oracle.h — CPU restatement of the reference's PBF step. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libpbf_b200.so) never links or calls it.
 *
 * Parity status: the reference ships no tests, fixtures or golden vectors for this path
 * (SURVEY.md 8c), so this oracle is pinned against the reference's OWN Simulator.cu,
 * compiled unchanged from /root/reference into oracle/_ref/libpbf_ref.so and run on a
 * B200; the outputs are committed under tests/golden/ (see tests/golden/make_golden.py).
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference's fluids/ directory). Arithmetic follows the PTX nvcc 12.9 emits for the
 * unchanged reference at -O3 (fma contraction spelled out with fmaf; compile this file
 * with -ffp-contract=off).
 */
#ifndef PBF_ORACLE_H_
#define PBF_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {   /* GUIParams.h:7-17 */
    int32_t niter;
    float pho0, g, h, dt, lambda_eps, delta_q, k_corr, n_corr, k_boundaryDensity, c_XSPH;
} orc_params;

typedef struct orc_sim orc_sim;

void orc_default_params(orc_params* p);                         /* FluidSystem.cpp:15-25 */
orc_sim* orc_create(const orc_params* p, const float ulim[3], const float llim[3],
                    int64_t max_particles);                     /* Simulator.h:10-26 */
void orc_destroy(orc_sim* s);                                   /* Simulator.h:27-34 */
void orc_set_params(orc_sim* s, const orc_params* p);           /* Simulator.cpp:101-115 */
void orc_set_lim(orc_sim* s, const float ulim[3], const float llim[3]); /* Simulator.cpp:132-136 */
void orc_set_threads(orc_sim* s, int nthreads);                 /* OpenMP threads for the timed baseline */
int  orc_max_threads(void);

/* Simulator::step on HOST arrays (Simulator.cpp:44-79): all five arrays are permuted in
 * place into cell-sorted order; results in npos / nvel / iid. */
void orc_step(orc_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n);

/* the five stages, after orc_bind (what step() does with the mapped pointers) */
void orc_bind(orc_sim* s, float* pos, float* npos, float* vel, float* nvel, uint32_t* iid, int64_t n);
void orc_advect(orc_sim* s);            /* Simulator.cu:165-176, Simulator_kernel.cuh:7-19 */
void orc_build_grid(orc_sim* s);        /* Simulator.cu:178-211, Simulator_kernel.cuh:21-50 */
void orc_correct_density(orc_sim* s);   /* Simulator.cu:213-249, one iteration */
void orc_lambda_pass(orc_sim* s);       /*   its first half:  Simulator.cu:222-233 */
void orc_delta_p_pass(orc_sim* s);      /*   its second half: Simulator.cu:235-248 */
void orc_update_velocity(orc_sim* s);   /* Simulator.cu:267-274 */
void orc_correct_velocity(orc_sim* s);  /* Simulator.cu:251-265 */

/* scratch accessors (valid until the next stage call) */
const uint32_t* orc_grid_id(const orc_sim* s);      /* dc_gridId, sorted, n entries */
const uint32_t* orc_grid_start(const orc_sim* s);   /* dc_gridStart, cells entries   */
const uint32_t* orc_grid_end(const orc_sim* s);     /* dc_gridEnd                    */
const float* orc_lambda(const orc_sim* s);          /* dc_lambda */
const float* orc_pho(const orc_sim* s);             /* dc_pho    */
const float* orc_tpos(const orc_sim* s);            /* dc_tpos   */
void orc_grid_dim(const orc_sim* s, int32_t dim[3]);
float orc_coef_corr(const orc_sim* s);              /* m_coef_corr after correct_density */
float orc_poly6_coef(const orc_sim* s);
float orc_spiky_coef(const orc_sim* s);
float orc_poly6(const orc_sim* s, float r2);        /* getPoly6::operator(), Simulator.cu:85-89 */

/* Derived quantities of SURVEY.md A.9 on the bound state. */
void orc_neighbor_count(const orc_sim* s, uint32_t* count_out);          /* r2 < h2 in the 27 cells */
void orc_candidate_count(const orc_sim* s, uint32_t* count_out);         /* all j in the 27 cells   */

/* The reference's dead DEBUG_NO_HASH_GRID idea (helper.h:37, Simulator_kernel.cuh:106-117):
 * all-pairs evaluation of rho / lambda and of the neighbour count on the bound npos,
 * to check the grid search against. O(n^2). */
void orc_lambda_allpairs(const orc_sim* s, float* lambda_out, float* pho_out, uint32_t* count_out);

/* Scene generators. */
int64_t orc_scene_cube(const float ulim[3], const float llim[3], const int32_t ns[3],
                       uint32_t* rng_state, uint32_t first_iid,
                       float* pos, float* vel, uint32_t* iid);  /* DoubleDamSource.cpp:5-21 */
int64_t orc_scene_double_dam_reference(float* pos, float* vel, uint32_t* iid,
                                       float ulim[3], float llim[3]); /* FluidSystem.cpp:34-35,55-61 */
void orc_scene_block(const float origin[3], const int32_t n[3], float spacing, uint32_t seed,
                     uint32_t first_iid, float* pos, float* vel, uint32_t* iid); /* SURVEY.md 8d */

/* Moving wall schedule (FluidSystem.cpp:104-110): ulim + A_ulim*sin(w*(frame-start)). */
void orc_wall_lim(const float ulim0[3], const float llim0[3], const float a_ulim[3],
                  const float a_llim[3], float w, int frame, int start_frame,
                  float ulim[3], float llim[3]);

/* Run statistics, f64 accumulation in index order. */
void orc_stats(const float* pho, const float* npos, const float* nvel, int64_t n, float pho0,
               double out[5]);

#ifdef __cplusplus
}
#endif
#endif
