// demo_main.cpp -- the reference's start-up sequence (main.cpp:94-108) without the window: load the scene with the
// reference's own MeshS/Material code, build form factors and converge the lighting through the shim classes.
//   shim_demo <scene.obj> <mtl_dir/> <method 0|1|2> <emission_value> <seed> [rands.bin] [matrix cache file]   (cwd must hold color_tables/srgb.coeff)
// Prints one line of results the parity test compares with the Python path.
#include <cstdio>
#include <vector>
#define TINYOBJLOADER_IMPLEMENTATION
#include <tiny_obj_loader.h>
#undef TINYOBJLOADER_IMPLEMENTATION
#include "OptixPrimeFunctionality.h"
#include "Lightning.h"

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: shim_demo scene.obj mtl_dir/ method emission seed\n"); return 2; }
    std::vector<float> wavelengths = { 200.0, 250.0, 300.0, 350.0, 400.0, 450.0, 500.0, 550.0, 600.0 }; // main.cpp:94
    MeshS mesh(argv[1], argv[2], wavelengths);
    for (auto &m : mesh.materials) // the UV lamp's M is undefined behaviour in the reference (Material.cpp:52-54): use M = 0
        if (m.spectral_values.size() && m.spectral_values[0] == 0.0f && m.spectral_values.back() == 0.0f && m.rgbcolor == glm::vec3(0.f))
            m.M.setZero();
    OptixPrimeFunctionality optixP(mesh, 0, atol(argv[5]));
    if (argc > 6) { // optional: a binary file of S x {u,v} floats replaces the rand() pattern (repeatable across C runtimes)
        std::vector<UV> r(RAYS_PER_PATCH);
        FILE *f = fopen(argv[6], "rb");
        if (f && fread(r.data(), sizeof(UV), r.size(), f) == r.size()) optixP.setSamples(r);
        if (f) fclose(f);
    }
    float emission = (float)atof(argv[4]);
    int method = atoi(argv[3]);
    Lightning *l = Lightning::get_lightning(method, mesh, optixP, emission, wavelengths, true, argc > 7 ? argv[7] : nullptr); // main.cpp:108 passes matfile
    double sum = 0;
    for (auto &band : l->lightningvalues) for (float v : band) sum += v;
    glm::vec3 c = l->get_color_of_patch(mesh.numtriangles / 2);
    // the batched visibility entry point on its own (OptixPrimeFunctionality.cpp:169-242): count and sum of the triplets
    std::vector<parallellism::Tripl> unoccluded;
    std::vector<Eigen::Triplet<double>> tr = optixP.calculateAllVisibility(unoccluded, mesh, optixP.rands);
    double tsum = 0;
    for (auto &t : tr) tsum += t.value();
    // the per-pair debugging entry points (InputHandler.cpp:184, OptixPrimeFunctionality.cpp:456-469) on two fixed patches
    const int pa = 3, pb = mesh.numtriangles / 2 + 5;
    float nus = optixP.p2pFormfactorNusselt(pa, pb, mesh), p2p = optixP.p2pFormfactor(pa, pb, mesh);
    std::vector<optix_functionality::Hit> picks(2);
    picks[0].t = 1; picks[0].triangleId = pa; picks[0].uv.x = 0.25f; picks[0].uv.y = 0.5f;
    picks[1].t = 1; picks[1].triangleId = pb; picks[1].uv.x = 0.3f; picks[1].uv.y = 0.3f;
    bool shot = optixP.shootPatchRay(picks, mesh);
    printf("RESULT passes=%d sumB=%.9e color=%.6f,%.6f,%.6f rand0=%.9g tripl=%zu,%.12e nusselt=%.9e p2p=%.9e shoot=%d\n", l->numpasses, sum, c.x,
           c.y, c.z, optixP.rands[0].u, tr.size(), tsum, nus, p2p, shot ? 1 : 0);
    delete l;
    return 0;
}
