// pbf_slab_headless.cpp — headless C++ harness for ONE scene on several GPUs (SlabSimulator.h).
//
// The multi-GPU sibling of pbf_headless.cpp: builds a named dam-break scene, cuts it into x-slabs over
// `--ranks` ranks (one host thread each; `--devices` maps ranks to CUDA devices, several ranks may share
// one), steps it through the library's fused peer-memory transport and prints throughput and run
// statistics as JSON. `--dump file` writes the final (pos, vel, iid) of all ranks in rank order — the
// single-rank run of the same scene writes the same bytes (tests/test_slab_gpu.py).
//
//   pbf_slab_headless [--scene small|dam_1m|double_dam_16m|dam_8m|dam_64m] [--ranks G] [--devices 0,1,..]
//                     [--weak] [--steps N] [--warmup W] [--ghost 5] [--margin 6] [--replan 20] [--dump file]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>

#include "GUIParams.h"
#include "SlabSimulator.h"

using namespace pbfslab;

static Block block(float ox, float oy, float oz, int nx, int ny, int nz) {
    Block b; b.origin[0] = ox; b.origin[1] = oy; b.origin[2] = oz; b.n[0] = nx; b.n[1] = ny; b.n[2] = nz; return b;
}
static bool make_scene(const std::string& name, Scene* sc) {   // the named scenes of pbf-cuda_b200/__init__.py SCENES
    sc->llim[0] = sc->llim[1] = sc->llim[2] = 0.f;
    sc->blocks.clear();
    if (name == "small") { sc->ulim[0] = 3.2f; sc->ulim[1] = 0.6f; sc->ulim[2] = 1.2f; sc->blocks.push_back(block(0.35f, 0.05f, 0.05f, 40, 8, 14)); }
    else if (name == "dam_1m") { sc->ulim[0] = 16.0f; sc->ulim[1] = 3.6f; sc->ulim[2] = 9.6f; sc->blocks.push_back(block(0.2f, 0.2f, 0.2f, 128, 64, 128)); }
    else if (name == "dam_8m") { sc->ulim[0] = 9.6f; sc->ulim[1] = 26.0f; sc->ulim[2] = 9.6f; sc->blocks.push_back(block(0.2f, 0.2f, 0.2f, 128, 512, 128)); }
    else if (name == "dam_64m") { sc->ulim[0] = 76.8f; sc->ulim[1] = 26.0f; sc->ulim[2] = 9.6f; sc->blocks.push_back(block(0.2f, 0.2f, 0.2f, 1024, 512, 128)); }
    else if (name == "double_dam_16m") {
        sc->ulim[0] = 38.4f; sc->ulim[1] = 38.4f; sc->ulim[2] = 9.6f;
        sc->blocks.push_back(block(0.2f, 25.4f, 0.2f, 256, 256, 128));
        sc->blocks.push_back(block(25.4f, 0.2f, 0.2f, 256, 256, 128));
    } else return false;
    return true;
}

int main(int argc, char** argv) {
    std::string scene_name = "dam_1m", dump, devices_arg;
    int ranks = 1, steps = 20, warmup = 5, ghost = 5, margin = 6, replan = 20;
    bool weak = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--scene") scene_name = next();
        else if (a == "--ranks") ranks = atoi(next());
        else if (a == "--devices") devices_arg = next();
        else if (a == "--steps") steps = atoi(next());
        else if (a == "--warmup") warmup = atoi(next());
        else if (a == "--ghost") ghost = atoi(next());
        else if (a == "--margin") margin = atoi(next());
        else if (a == "--replan") replan = atoi(next());
        else if (a == "--dump") dump = next();
        else if (a == "--weak") weak = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    Scene sc;
    if (!make_scene(scene_name, &sc)) { fprintf(stderr, "unknown scene %s\n", scene_name.c_str()); return 2; }
    if (weak) {   // per-rank work fixed: the block and the box repeated `ranks` times along x
        if (sc.blocks.size() != 1) { fprintf(stderr, "--weak needs a single-block scene\n"); return 2; }
        sc.blocks[0].n[0] *= ranks;
        sc.ulim[0] *= ranks;
    }
    int ndev = 0;
    if (pbf_device_count(&ndev) != PBF_OK || ndev < 1) { fprintf(stderr, "no CUDA device: %s\n", pbf_last_error()); return 1; }
    std::vector<int> devices;
    if (!devices_arg.empty()) {
        char* copy = strdup(devices_arg.c_str());
        for (char* t = strtok(copy, ","); t; t = strtok(NULL, ",")) devices.push_back(atoi(t));
        free(copy);
        if ((int)devices.size() != ranks) { fprintf(stderr, "--devices lists %zu devices for %d ranks\n", devices.size(), ranks); return 2; }
    } else {
        for (int r = 0; r < ranks; r++) devices.push_back(r % ndev);
    }
    const pbf_params params = GUIParams::getInstance().toC();   // the reference's defaults (FluidSystem.cpp:15-25)
    SlabRun run(params, sc, devices, ghost, margin, replan);
    if (!run.init()) { fprintf(stderr, "init failed: %s\n", run.error().c_str()); return 1; }
    int64_t n_total = 0, n_min = -1, n_max = 0;
    for (int r = 0; r < ranks; r++) {
        const int64_t n = run.particles(r);
        n_total += n; if (n > n_max) n_max = n; if (n_min < 0 || n < n_min) n_min = n;
    }
    if (warmup > 0 && !run.step(warmup)) { fprintf(stderr, "step failed: %s\n", run.error().c_str()); return 1; }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (steps > 0 && !run.step(steps)) { fprintf(stderr, "step failed: %s\n", run.error().c_str()); return 1; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    double ke = 0, err_sum = 0, err_max = -1e300, vmax = 0;
    int64_t launches = 0;
    for (int r = 0; r < ranks; r++) {
        pbf_stats st;
        if (run.particles(r) > 0 && run.stats(r, &st)) {
            ke += st.kinetic_energy; err_sum += st.density_err_mean * run.particles(r);
            if (st.density_err_max > err_max) err_max = st.density_err_max;
            if (st.max_speed > vmax) vmax = st.max_speed;
        }
        launches += run.launches(r);
    }
    printf("{\"scene\": \"%s\", \"weak\": %d, \"ranks\": %d, \"devices\": %d, \"particles\": %lld, \"particles_per_rank\": [%lld, %lld], "
           "\"steps\": %d, \"warmup\": %d, \"seconds\": %.6f, \"ms_per_step\": %.5f, \"particle_steps_per_s\": %.1f, "
           "\"density_err_mean\": %.9g, \"density_err_max\": %.9g, \"kinetic_energy\": %.9g, \"max_speed\": %.9g, \"launches\": %lld, "
           "\"ghost\": %d, \"margin\": %d, \"boundaries\": [",
           scene_name.c_str(), weak ? 1 : 0, ranks, ndev, (long long)n_total, (long long)n_min, (long long)n_max, steps, warmup, sec,
           steps ? 1e3 * sec / steps : 0.0, steps ? n_total * (double)steps / sec : 0.0, err_sum / (double)n_total, err_max, ke, vmax,
           (long long)launches, ghost, margin);
    for (size_t i = 0; i < run.bounds().size(); i++) printf("%s%d", i ? ", " : "", run.bounds()[i]);
    printf("]}\n");
    if (!dump.empty()) {
        FILE* f = fopen(dump.c_str(), "wb");
        if (!f) { perror(dump.c_str()); return 1; }
        const int n32 = (int)n_total;
        fwrite(&n32, sizeof(int), 1, f);
        std::vector<std::vector<float> > pos(ranks), vel(ranks);
        std::vector<std::vector<uint32_t> > iid(ranks);
        for (int r = 0; r < ranks; r++)
            if (!run.download(r, &pos[r], &vel[r], &iid[r])) { fprintf(stderr, "download failed: %s\n", pbf_last_error()); return 1; }
        for (int r = 0; r < ranks; r++) fwrite(pos[r].data(), 4, pos[r].size(), f);
        for (int r = 0; r < ranks; r++) fwrite(vel[r].data(), 4, vel[r].size(), f);
        for (int r = 0; r < ranks; r++) fwrite(iid[r].data(), 4, iid[r].size(), f);
        fclose(f);
    }
    return 0;
}
