// GUIParams.h — mirror of the reference's parameter singleton (fluids/GUIParams.h:3-44,
// GUIParams.cpp:5-12). Same class, same field names, same lazily-created process-wide instance;
// the fluid half maps 1:1 onto pbf_params of the C-ABI. The renderer half is kept only so that
// code touching those fields still compiles (the viewer is out of scope).
#pragma once
#include "../../include/pbf.h"

class GUIParams {
public:
    /* fluid params (GUIParams.h:7-17) */
    int niter;
    float pho0, g, h, dt, lambda_eps, delta_q, k_corr, n_corr, k_boundaryDensity, c_XSPH;
    /* renderer params (GUIParams.h:20-38), unused by the solver */
    int kernel_r; float sigma_r, sigma_z; int smooth_niter, keep_edge, blur_option;
    enum ShadeOption { Full = 0, Depth, Thick, Normal, Fresnel, Reflect, Refract, RefractBL };
    ShadeOption shading_option;

    static GUIParams& getInstance() {
        static GUIParams* instance = 0;
        if (instance == 0) {
            instance = new GUIParams();
            instance->setDefaults();
        }
        return *instance;
    }
    // defaults FluidSystem::FluidSystem writes (FluidSystem.cpp:15-32)
    void setDefaults() {
        pbf_params p;
        pbf_default_params(&p);
        fromC(p);
        smooth_niter = 2; kernel_r = 10; sigma_r = 6.f; sigma_z = 0.1f;
        shading_option = Full; keep_edge = 1; blur_option = 0;
    }
    pbf_params toC() const {
        pbf_params p;
        p.niter = niter; p.pho0 = pho0; p.g = g; p.h = h; p.dt = dt; p.lambda_eps = lambda_eps; p.delta_q = delta_q;
        p.k_corr = k_corr; p.n_corr = n_corr; p.k_boundaryDensity = k_boundaryDensity; p.c_XSPH = c_XSPH;
        return p;
    }
    void fromC(const pbf_params& p) {
        niter = p.niter; pho0 = p.pho0; g = p.g; h = p.h; dt = p.dt; lambda_eps = p.lambda_eps; delta_q = p.delta_q;
        k_corr = p.k_corr; n_corr = p.n_corr; k_boundaryDensity = p.k_boundaryDensity; c_XSPH = p.c_XSPH;
    }
};
