// OptixPrimeFunctionality.h -- drop-in replacement of the reference class of the same name for the form-factor path
// (reference: visual studio/OptixPrimeFunctionality.h:24-48, .cpp:6-81, :133-271, :311-366), forwarding to the
// C-ABI of include/daisy_b200.h.  Host code written against DaisyRiot keeps compiling: same class, same method
// names and argument lists for optixQuery / cudaCalculateRadiosityMatrix / calculateRadiosityMatrix /
// calculateVisibility / p2pFormfactor.  The OptiX Prime context/model members are gone (nothing here uses OptiX);
// the camera/picking helpers traceScreen and intersectMouse are UI code outside this path (they only need optixQuery).
//
// Error behaviour follows the reference: failures are printed to std::cerr and execution continues
// (.cpp:45-53, :72-79, parallellism.cuh:21-26); the C layer underneath is strict and keeps the message.
#pragma once
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <iostream>
#include <vector>
#include <glm/glm.hpp>
#include <Eigen/Sparse>
#include "Vertex.h"   // reference headers, unchanged: vertex::TriangleIndex
#include "Defines.h"  // UV, RAYS_PER_PATCH
#include "MeshS.h"
#include "daisy_b200.h"

typedef Eigen::SparseMatrix<float> SpMat;
typedef Eigen::Triplet<double> Tripl;

namespace optix { struct float2 { float x, y; }; struct float3 { float x, y, z; }; }
namespace optix_functionality { struct Hit { float t; int triangleId; optix::float2 uv; }; } // optix_functionality.h:10-14
namespace parallellism { struct Tripl { int m_row, m_col; double m_value; }; }               // parallellism.cuh:30-33

static_assert(sizeof(optix_functionality::Hit) == sizeof(daisy_hit), "Hit layout");
static_assert(sizeof(parallellism::Tripl) == sizeof(daisy_tripl), "Tripl layout");
static_assert(sizeof(vertex::TriangleIndex) == 6 * sizeof(int), "TriangleIndex layout");

class OptixPrimeFunctionality {
public:
    daisy_ctx *ctx = nullptr; // replaces optix::prime::Context contextP + optix::prime::Model model

    // .cpp:36-64: build the acceleration structure over the triangle soup and fix the 50-sample pattern.
    // seed < 0 keeps the reference's behaviour (srand(time)); pass a seed to make runs repeatable.
    explicit OptixPrimeFunctionality(MeshS &mesh, int device = 0, long seed = -1) {
        report(daisy_ctx_create(reinterpret_cast<const float *>(mesh.vertices.data()), (int)mesh.vertices.size(),
                                reinterpret_cast<const float *>(mesh.normals.data()), (int)mesh.normals.size(),
                                reinterpret_cast<const int32_t *>(mesh.triangleIndices.data()), (int)mesh.triangleIndices.size(), device, &ctx));
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
    ~OptixPrimeFunctionality() { daisy_ctx_destroy(ctx); }
    OptixPrimeFunctionality(const OptixPrimeFunctionality &) = delete;
    OptixPrimeFunctionality &operator=(const OptixPrimeFunctionality &) = delete;

    void setSamples(const std::vector<UV> &r) {
        rands = r;
        if (ctx) report(daisy_ctx_set_samples(ctx, reinterpret_cast<const float *>(rands.data()), (int)rands.size()));
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
        if (!report(daisy_formfactors_build(ctx, DAISY_FF_DEVICE))) return out;
        int64_t nnz = 0;
        if (!report(daisy_formfactors_to_csc(ctx, &nnz, nullptr, nullptr, nullptr))) return out;
        std::vector<float> val((size_t)nnz);
        std::vector<int> inner((size_t)nnz), outer((size_t)mesh.numtriangles + 1);
        if (!report(daisy_formfactors_to_csc(ctx, &nnz, val.data(), inner.data(), outer.data()))) return out;
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

    std::vector<UV> rands;

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
        if (!report(daisy_formfactors_build(ctx, variant))) return;
        int64_t pairs = 0, owned = 0, rays = 0; double lbvh = 0, ff = 0;
        daisy_formfactors_stats(ctx, &pairs, &owned, &rays, &lbvh, &ff);
        std::cout << "Calculation time of form factors + visibility: " << ff * 1e-3 << " s (" << rays << " rays)" << std::endl;
        int64_t nnz = 0;
        if (!report(daisy_formfactors_to_csc(ctx, &nnz, nullptr, nullptr, nullptr))) return;
        if (nnz > 0x7fffffffLL) { std::cerr << "RadMat not refilled: more non-zeros than Eigen's int index holds (matrix stays on the GPU)" << std::endl; return; }
        RadMat.resize(mesh.numtriangles, mesh.numtriangles);
        RadMat.makeCompressed();
        RadMat.resizeNonZeros((int)nnz);
        report(daisy_formfactors_to_csc(ctx, &nnz, RadMat.valuePtr(), RadMat.innerIndexPtr(), RadMat.outerIndexPtr()));
        std::cout << "... done!" << std::endl;
    }
    static bool report(int rc) {
        if (rc != DAISY_OK) std::cerr << "An error occurred with error code " << rc << " and message " << daisy_last_error() << std::endl;
        return rc == DAISY_OK;
    }
};
