// OptixPrimeFunctionality.h -- drop-in replacement of the reference class of the same name for the form-factor path
// (reference: visual studio/OptixPrimeFunctionality.h:24-48, .cpp:6-81, :133-271, :311-366), forwarding to the
// C-ABI of include/daisy_b200.h.  Host code written against DaisyRiot keeps compiling: same class, same method
// names and argument lists for optixQuery / cudaCalculateRadiosityMatrix / calculateRadiosityMatrix /
// calculateVisibility / p2pFormfactor / traceScreen / intersectMouse.  The OptiX Prime context/model members are gone
// (nothing here uses OptiX).  One extra constructor argument, `ndevices`, spreads the matrix and the solve over that many
// GPUs of the box from this one process (daisy_group_*); the default 1 is the reference's single-device behaviour.
//
// Error behaviour follows the reference: failures are printed to std::cerr and execution continues
// (.cpp:45-53, :72-79, parallellism.cuh:21-26); the C layer underneath is strict and keeps the message.
#pragma once
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <iostream>
#include <vector>
#include <cmath>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <Eigen/Sparse>
#include "Vertex.h"   // reference headers, unchanged: vertex::TriangleIndex
#include "Defines.h"  // UV, RAYS_PER_PATCH
#include "MeshS.h"
#include "daisy_b200.h"

typedef Eigen::SparseMatrix<float> SpMat;
typedef Eigen::Triplet<double> Tripl;

// the optix:: vector vocabulary the reference's host code uses (Camera.h, optix_functionality.h); the OptiX SDK headers are
// no longer needed.  Skipped when the includer already has them (define DAISY_HAVE_OPTIX_MATH).
#ifndef DAISY_HAVE_OPTIX_MATH
namespace optix {
struct float2 { float x, y; };
struct float3 { float x, y, z; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float3 make_float3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline float3 operator+(const float3 &a, const float3 &b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(const float3 &a, const float3 &b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(const float3 &a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
static inline float3 operator*(float s, const float3 &a) { return make_float3(a.x * s, a.y * s, a.z * s); }
static inline float dot(const float3 &a, const float3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float3 normalize(const float3 &v) { return v * (1.0f / sqrtf(dot(v, v))); }
} // namespace optix
#endif
namespace optix_functionality {
struct Hit { float t; int triangleId; optix::float2 uv; };                                  // optix_functionality.h:10-14
static inline optix::float3 glm2optixf3(glm::vec3 v) { return optix::make_float3(v.x, v.y, v.z); } // optix_functionality.cpp
static inline glm::vec3 optix2glmf3(optix::float3 v) { return glm::vec3(v.x, v.y, v.z); }
} // namespace optix_functionality
namespace parallellism { struct Tripl { int m_row, m_col; double m_value; }; }               // parallellism.cuh:30-33

static_assert(sizeof(optix_functionality::Hit) == sizeof(daisy_hit), "Hit layout");
static_assert(sizeof(parallellism::Tripl) == sizeof(daisy_tripl), "Tripl layout");
static_assert(sizeof(vertex::TriangleIndex) == 6 * sizeof(int), "TriangleIndex layout");

class OptixPrimeFunctionality {
public:
    daisy_group *group = nullptr; // replaces optix::prime::Context contextP + optix::prime::Model model: one context per GPU
    daisy_ctx *ctx = nullptr;     // the first device's context (whole mesh + LBVH): closest-hit queries, per-pair entry points

    // .cpp:36-64: build the acceleration structure over the triangle soup and fix the 50-sample pattern.
    // seed < 0 keeps the reference's behaviour (srand(time)); pass a seed to make runs repeatable.  ndevices > 1: devices
    // device .. device + ndevices - 1 share the matrix (row blocks) and the solve.
    explicit OptixPrimeFunctionality(MeshS &mesh, int device = 0, long seed = -1, int ndevices = 1) {
        std::vector<int> ids;
        for (int i = 0; i < (ndevices > 1 ? ndevices : 1); i++) ids.push_back(device + i);
        if (report(daisy_group_create(reinterpret_cast<const float *>(mesh.vertices.data()), (int)mesh.vertices.size(),
                                      reinterpret_cast<const float *>(mesh.normals.data()), (int)mesh.normals.size(),
                                      reinterpret_cast<const int32_t *>(mesh.triangleIndices.data()), (int)mesh.triangleIndices.size(), ids.data(),
                                      (int)ids.size(), &group)))
            ctx = daisy_group_ctx(group, 0);
        rands.resize(RAYS_PER_PATCH);
        std::srand(seed < 0 ? (unsigned)std::time(nullptr) : (unsigned)seed);
        for (size_t i = 0; i < RAYS_PER_PATCH; i++) {
            UV uv = UV();
            uv.u = ((float)(rand() % RAND_MAX)) / RAND_MAX;
            uv.v = ((float)(rand() % RAND_MAX)) / RAND_MAX;
            uv.v = uv.v * (1 - uv.u);
            rands[i] = uv;
        }
        setSamples(rands);
    }
    ~OptixPrimeFunctionality() { daisy_group_destroy(group); } // destroy the Lightning objects first: their solvers live on these contexts
    OptixPrimeFunctionality(const OptixPrimeFunctionality &) = delete;
    OptixPrimeFunctionality &operator=(const OptixPrimeFunctionality &) = delete;

    void setSamples(const std::vector<UV> &r) {
        rands = r;
        if (group) report(daisy_group_set_samples(group, reinterpret_cast<const float *>(rands.data()), (int)rands.size()));
    }

    // .cpp:66-81 -- rays: origin,direction pairs; hits: one Hit per ray (miss: t < 0)
    void optixQuery(int number_of_rays, std::vector<optix::float3> &rays, std::vector<optix_functionality::Hit> &hits) {
        report(daisy_query_closest(ctx, number_of_rays, reinterpret_cast<const float *>(rays.data()), reinterpret_cast<daisy_hit *>(hits.data())));
    }

    // .cpp:6-34 -- unoccluded form factors + visibility + RadMat.setFromTriplets, all on the GPU; the matrix stays
    // resident for the Lightning classes and RadMat is refilled only if the caller's matrix can hold it
    void cudaCalculateRadiosityMatrix(SpMat &RadMat, MeshS &mesh) { build(RadMat, mesh, DAISY_FF_DEVICE); }
    // .cpp:311-366 -- the cuda_on = false arithmetic (float pi, mirrored entry by reciprocity)
    void calculateRadiosityMatrix(SpMat &RadMat, MeshS &mesh) { build(RadMat, mesh, DAISY_FF_HOST); }

    // parallellism::runCalculateRadiosityMatrix(SimpleMesh&) -- parallellism.cu:4-89
    std::vector<parallellism::Tripl> runCalculateRadiosityMatrix(MeshS &mesh) {
        std::vector<parallellism::Tripl> out((size_t)mesh.numtriangles * mesh.numtriangles);
        report(daisy_unoccluded_rows(ctx, DAISY_FF_DEVICE, 0, mesh.numtriangles, reinterpret_cast<daisy_tripl *>(out.data())));
        return out;
    }

    // .cpp:244-271
    float calculateVisibility(int originPatch, int destPatch, MeshS &mesh) {
        std::vector<optix::float3> rays(2 * rands.size());
        std::vector<optix_functionality::Hit> hits(rands.size());
        for (size_t i = 0; i < rands.size(); i++) {
            glm::vec3 o = uv2xyz(originPatch, rands[i], mesh), d = uv2xyz(destPatch, rands[i], mesh);
            glm::vec3 diff = d - o;
            float inv = 1.0f / sqrtf(diff.x * diff.x + diff.y * diff.y + diff.z * diff.z);
            glm::vec3 n(diff.x * inv, diff.y * inv, diff.z * inv);
            glm::vec3 org = o + n * 0.000001f;
            rays[2 * i] = { org.x, org.y, org.z };
            rays[2 * i + 1] = { n.x, n.y, n.z };
        }
        optixQuery((int)rands.size(), rays, hits);
        float visibility = 0;
        for (auto &hit : hits) visibility += (hit.t > 0 && hit.triangleId == destPatch) ? 1 : 0;
        return visibility / rands.size();
    }

    // .cpp:169-242 -- every pair row < col of the triplet list with a positive unoccluded factor is traced with the sample
    // pattern; pairs with visibility > 0 yield (row, col, vis * ff(row,col)) and (col, row, vis * ff(col,row)).  The reference
    // signature also takes the OptiX context and model; here the fused GPU kernel does list, rays and reduction in one go, so
    // the unoccluded list and the pattern arguments are implied by the context (same values by construction).  The order of the
    // returned triplets is column-major instead of the reference's push order; setFromTriplets does not depend on it.
    std::vector<Eigen::Triplet<double>> calculateAllVisibility(std::vector<parallellism::Tripl> & /*tripletlist*/, MeshS &mesh,
                                                               std::vector<UV> & /*rands*/) {
        std::vector<Eigen::Triplet<double>> out;
        if (!report(daisy_group_formfactors_build(group, DAISY_FF_DEVICE))) return out;
        int64_t nnz = 0;
        if (!report(daisy_group_formfactors_to_csc(group, &nnz, nullptr, nullptr, nullptr))) return out;
        std::vector<float> val((size_t)nnz);
        std::vector<int> inner((size_t)nnz), outer((size_t)mesh.numtriangles + 1);
        if (!report(daisy_group_formfactors_to_csc(group, &nnz, val.data(), inner.data(), outer.data()))) return out;
        out.reserve((size_t)nnz);
        for (int c = 0; c < mesh.numtriangles; c++)
            for (int i = outer[c]; i < outer[c + 1]; i++) out.emplace_back(inner[i], c, (double)val[i]);
        return out;
    }

    // .cpp:133-167 (unoccluded 4x4 rule times visibility, host arithmetic)
    float p2pFormfactor(int originPatch, int destPatch, MeshS &mesh) {
        daisy_tripl t;
        std::vector<daisy_tripl> row((size_t)mesh.numtriangles);
        report(daisy_unoccluded_rows(ctx, DAISY_FF_HOST, originPatch, 1, row.data()));
        t = row[destPatch];
        return (float)t.m_value * calculateVisibility(originPatch, destPatch, mesh);
    }

    // .cpp:273-306 -- Nusselt analogue (debug read-out of InputHandler.cpp:184): destination triangle projected onto the unit
    // hemisphere around the origin patch's centre, then onto its plane; area / pi times the sampled visibility.  The reference
    // guards with `if (isFacingBack(a), isFacingBack(b))`, a comma expression: only the second test counts (kept).
    float p2pFormfactorNusselt(int originPatch, int destPatch, MeshS &mesh) {
        glm::vec3 center_origin = centre(originPatch, mesh), center_dest = centre(destPatch, mesh);
        glm::vec3 normal_origin = avgNormal(originPatch, mesh);
        if (isFacingBack(center_dest, originPatch, mesh)) return 0.0f;
        glm::vec3 proj[3];
        for (int i = 0; i < 3; i++) {
            glm::vec3 hemi = center_origin + glm::normalize(mesh.vertices[mesh.triangleIndices[destPatch].vertex[i]] - center_origin);
            proj[i] = hemi - glm::dot(normal_origin, hemi - center_origin) * normal_origin;
        }
        float surface = 0.5 * glm::length(glm::cross(proj[1] - proj[0], proj[2] - proj[0]));
        return (surface / M_PIf) * calculateVisibility(originPatch, destPatch, mesh);
    }

    // .cpp:456-469 -- one ray between two picked surface points: does it reach the second patch?
    bool shootPatchRay(std::vector<optix_functionality::Hit> &patches, MeshS &mesh) {
        UV ua = { patches[0].uv.x, patches[0].uv.y }, ub = { patches[1].uv.x, patches[1].uv.y };
        glm::vec3 a = uv2xyz(patches[0].triangleId, ua, mesh), b = uv2xyz(patches[1].triangleId, ub, mesh);
        glm::vec3 d = glm::normalize(b - a), o = a + d * 0.000001f;
        std::vector<optix::float3> ray(2);
        ray[0] = { o.x, o.y, o.z };
        ray[1] = { d.x, d.y, d.z };
        std::vector<optix_functionality::Hit> hit(1);
        optixQuery(1, ray, hit);
        return hit[0].triangleId == patches[1].triangleId;
    }

    // .cpp:83-131 -- camera ray cast + shading.  RenderContext is the reference's Drawer::RenderContext (taken by value like
    // there; its members are references) or any struct with the same members: camera, optixView, trianglesonScreen, mesh,
    // lightning, radiosityRendering, antialiasing, supersampling.  Closest hit, isFacingBack, Drawer::interpolate
    // (Drawer.cpp:161-186), the supersample average and the clamp run in one kernel; trianglesonScreen (the picking index)
    // is rebuilt from the hit records in the reference's x-major pixel order.
    template <class RenderContext>
    void traceScreen(RenderContext cntxt) {
        const int W = cntxt.camera.pixwidth, H = cntxt.camera.pixheight;
        const int samples = cntxt.antialiasing ? cntxt.supersampling : 1;
        cntxt.optixView.resize((size_t)W * H);
        std::vector<optix::float3> rays;
        cntxt.camera.gen_rays_for_screen(rays, cntxt.antialiasing);
        const int N = (int)cntxt.mesh.triangleIndices.size();
        std::vector<glm::vec3> rgb((size_t)N);
        for (int i = 0; i < N; i++)
            rgb[(size_t)i] = cntxt.radiosityRendering ? cntxt.lightning.get_color_of_patch(i) : cntxt.mesh.materials[cntxt.mesh.materialIndexPerTriangle[i]].rgbcolor;
        std::vector<optix_functionality::Hit> hits((size_t)W * H * samples);
        const glm::vec3 eye = optix_functionality::optix2glmf3(cntxt.camera.eye);
        if (!report(daisy_trace_screen(ctx, W, H, samples, reinterpret_cast<const float *>(rays.data()), &eye.x, reinterpret_cast<const float *>(rgb.data()),
                                       cntxt.radiosityRendering ? 1 : 0, reinterpret_cast<float *>(cntxt.optixView.data()),
                                       reinterpret_cast<daisy_hit *>(hits.data()))))
            return;
        cntxt.trianglesonScreen.clear();
        cntxt.trianglesonScreen.resize((size_t)N);
        for (int x = 0; x < W; x++)
            for (int y = 0; y < H; y++)
                for (int i = 0; i < samples; i++) {
                    const optix_functionality::Hit &h = hits[((size_t)y * W + x) * samples + i];
                    if (h.t > 0 && !isFacingBack(eye, h.triangleId, cntxt.mesh)) {
                        MatrixIndex index = {};
                        index.col = x; index.row = y; index.uv = { h.uv.x, h.uv.y };
                        cntxt.trianglesonScreen[(size_t)h.triangleId].push_back(index);
                    }
                }
    }

    // .cpp:471-517 -- picking: one ray through the cursor; the first pick selects patch 0, the second patch 1 and shoots a
    // ray between them.  DebugLine / Camera are the reference's types (or look-alikes with left, debugtriangles / eye, dir,
    // up, viewport).
    template <class DebugLine, class CameraT>
    bool intersectMouse(DebugLine &debugline, double xpos, double ypos, CameraT &camera, std::vector<std::vector<MatrixIndex>> & /*trianglesonScreen*/,
                        std::vector<glm::vec3> & /*optixView*/, std::vector<optix_functionality::Hit> &patches, MeshS &mesh) {
        bool hitB = true;
        std::vector<optix::float3> ray(2);
        std::vector<optix_functionality::Hit> hit(1);
        glm::mat4x4 lookat = glm::lookAt(optix_functionality::optix2glmf3(camera.eye), optix_functionality::optix2glmf3(camera.dir), optix_functionality::optix2glmf3(camera.up));
        glm::mat4x4 projection = glm::perspective(45.0f, (float)(800) / (float)(600), 0.1f, 1000.0f);
        ray[0] = optix_functionality::glm2optixf3(glm::unProject(glm::vec3(xpos, ypos, 0.0), lookat, projection, camera.viewport));
        ray[1] = optix_functionality::glm2optixf3(glm::unProject(glm::vec3(xpos, ypos, 1.0), lookat, projection, camera.viewport));
        optixQuery(1, ray, hit);
        if (hit[0].t > 0) {
            printf("\nhit triangle: %i ", hit[0].triangleId);
            if (debugline.left) patches[0] = hit[0];
            else {
                patches[1] = hit[0];
                printf("\nshoot ray between patches \n");
                printf("patch triangle 1: %i \n", patches[0].triangleId);
                printf("patch triangle 2: %i \n", patches[1].triangleId);
                hitB = shootPatchRay(patches, mesh);
                printf("\ndid it hit? %i", hitB);
            }
            debugline.left = !debugline.left;
            debugline.debugtriangles.push_back(hit[0].triangleId);
        } else {
            printf("miss!");
            hitB = false;
            patches.clear();
            patches.resize(2);
            debugline.left = true;
        }
        return hitB;
    }

    std::vector<UV> rands;
    bool refill_RadMat = true; // false: leave the caller's Eigen matrix empty after a build (large scenes: the matrix lives on the GPUs)

private:
    static glm::vec3 centre(int tri, MeshS &mesh) { // triangle_math.cpp:16-21
        glm::vec3 c = mesh.vertices[mesh.triangleIndices[tri].vertex.x] + mesh.vertices[mesh.triangleIndices[tri].vertex.y] +
                      mesh.vertices[mesh.triangleIndices[tri].vertex.z];
        return glm::vec3(c.x / 3, c.y / 3, c.z / 3);
    }
    static glm::vec3 avgNormal(int tri, MeshS &mesh) { // triangle_math.cpp:23-29
        glm::vec3 n = mesh.normals[mesh.triangleIndices[tri].normal.x] + mesh.normals[mesh.triangleIndices[tri].normal.y] +
                      mesh.normals[mesh.triangleIndices[tri].normal.z];
        return glm::normalize(glm::vec3(n.x / 3, n.y / 3, n.z / 3));
    }
    static bool isFacingBack(glm::vec3 origin, int destPatch, MeshS &mesh) { // triangle_math.cpp:76-86
        return glm::dot(glm::normalize(centre(destPatch, mesh) - origin), avgNormal(destPatch, mesh)) >= 0;
    }
    static glm::vec3 uv2xyz(int tri, const UV &uv, MeshS &mesh) { // triangle_math.cpp:3-9
        glm::vec3 a = mesh.vertices[mesh.triangleIndices[tri].vertex.x];
        glm::vec3 b = mesh.vertices[mesh.triangleIndices[tri].vertex.y];
        glm::vec3 c = mesh.vertices[mesh.triangleIndices[tri].vertex.z];
        return a + uv.u * (b - a) + uv.v * (c - a);
    }
    void build(SpMat &RadMat, MeshS &mesh, int variant) {
        std::cout << "Calculating radiosity matrix..." << std::endl;
        std::cout << "Number of triangles: " << mesh.triangleIndices.size() << std::endl;
        if (!report(daisy_group_formfactors_build(group, variant))) return;
        int64_t pairs = 0, rays = 0; double lbvh = 0, ff = 0;
        daisy_group_formfactors_stats(group, &pairs, &rays, &lbvh, &ff);
        std::cout << "Calculation time of form factors + visibility: " << ff * 1e-3 << " s (" << rays << " rays, " << daisy_group_size(group) << " GPU(s))" << std::endl;
        if (!refill_RadMat) { std::cout << "... done! (matrix resident on the GPUs, RadMat not refilled)" << std::endl; return; }
        int64_t nnz = 0;
        if (!report(daisy_group_formfactors_to_csc(group, &nnz, nullptr, nullptr, nullptr))) return;
        if (nnz > 0x7fffffffLL) { std::cerr << "RadMat not refilled: more non-zeros than Eigen's int index holds (matrix stays on the GPU)" << std::endl; return; }
        RadMat.resize(mesh.numtriangles, mesh.numtriangles);
        RadMat.makeCompressed();
        RadMat.resizeNonZeros((int)nnz);
        report(daisy_group_formfactors_to_csc(group, &nnz, RadMat.valuePtr(), RadMat.innerIndexPtr(), RadMat.outerIndexPtr()));
        std::cout << "... done!" << std::endl;
    }
    static bool report(int rc) {
        if (rc != DAISY_OK) std::cerr << "An error occurred with error code " << rc << " and message " << daisy_last_error() << std::endl;
        return rc == DAISY_OK;
    }
};
