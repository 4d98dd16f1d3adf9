// ref_wrap.cpp -- extern "C" doorway into the reference's OWN sources, compiled unmodified from
// /root/reference (triangle_math.cpp, MeshS.cpp, Material.cpp, rgb2spec.cpp, Lightning.h + vendored
// Eigen 3.2.10 / glm 0.9.8.4 / tinyobjloader).  TEST INFRASTRUCTURE ONLY: it pins oracle/daisy_oracle.c and
// serves as the "reference" CPU baseline.  Nothing here is linked into the product.
//
// What cannot be compiled is OptixPrimeFunctionality.cpp (needs the closed OptiX Prime SDK).  The stub class
// below stands in for it: its two matrix builders fill RadMat from triplets handed in by the caller, exactly
// like RadMat.setFromTriplets at OptixPrimeFunctionality.cpp:25 / :364.
#include <vector>
#include <string>
#include <iostream>
#include <fstream>
#include <sstream>
#include <cstdint>
#include <glm/glm.hpp>
#include <Eigen/Sparse>
#include <Eigen/Dense>
#define TINYOBJLOADER_IMPLEMENTATION // reference does this in main.cpp:15
#include <tiny_obj_loader.h>
#undef TINYOBJLOADER_IMPLEMENTATION
#include "MeshS.h"
#include "triangle_math.h"
#include "Defines.h"

typedef Eigen::SparseMatrix<float> SpMat;
typedef Eigen::Triplet<double> Tripl;

class OptixPrimeFunctionality {
public:
    std::vector<Tripl> triplets; // what calculateAllVisibility / the per-pair loop would have produced
    void cudaCalculateRadiosityMatrix(SpMat &RadMat, MeshS &mesh) { RadMat.setFromTriplets(triplets.begin(), triplets.end()); }
    void calculateRadiosityMatrix(SpMat &RadMat, MeshS &mesh) { RadMat.setFromTriplets(triplets.begin(), triplets.end()); }
};

// expose the band vectors of the Lightning classes without touching the header
#define private public
#define protected public
#include "Lightning.h"
#undef private
#undef protected

// Camera.h (ray generation of traceScreen) needs the two float3 converters of optix_functionality.cpp:78-84; that file
// itself cannot be compiled (full OptiX API), so the two one-liners are restated here.
#include <glm/gtc/matrix_transform.hpp>
namespace optix_functionality {
inline optix::float3 glm2optixf3(glm::vec3 v) { return optix::make_float3(v.x, v.y, v.z); }
inline glm::vec3 optix2glmf3(optix::float3 v) { return glm::vec3(v.x, v.y, v.z); }
}
#include "Camera.h"

struct RefScene {
    std::vector<float> wavelengths; // Material keeps a reference to this vector
    MeshS mesh;
    OptixPrimeFunctionality optixP;
    Lightning *lightning = nullptr;
    int method = -1;
    float emission_value = 0;
};

static std::streambuf *g_cout_saved = nullptr;
static std::ostringstream g_sink;
static void quiet(bool on) {
    if (on && !g_cout_saved) g_cout_saved = std::cout.rdbuf(g_sink.rdbuf());
    if (!on && g_cout_saved) { std::cout.rdbuf(g_cout_saved); g_cout_saved = nullptr; g_sink.str(""); }
}

extern "C" {

void ref_flush_stdout(void) { fflush(stdout); }

// ---- scene -------------------------------------------------------------------------------------
// MeshS::loadFromFile (MeshS.cpp:22-128).  cwd must hold color_tables/srgb.coeff (Material.cpp:11).
void *ref_scene_load(const char *obj, const char *mtl_dir, const float *wavelengths, int nw) {
    RefScene *s = new RefScene();
    s->wavelengths.assign(wavelengths, wavelengths + nw);
    quiet(true);
    s->mesh.loadFromFile((char *)obj, (char *)mtl_dir, s->wavelengths);
    quiet(false);
    return s;
}
// geometry only (no materials): fill the public vectors directly (MeshS.h:14-20)
void *ref_scene_from_arrays(const float *v, int nv, const float *n, int nn, const int *tri, int ntri) {
    RefScene *s = new RefScene();
    for (int i = 0; i < nv; i++) s->mesh.vertices.push_back(glm::vec3(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
    for (int i = 0; i < nn; i++) s->mesh.normals.push_back(glm::vec3(n[3 * i], n[3 * i + 1], n[3 * i + 2]));
    for (int i = 0; i < ntri; i++) {
        vertex::TriangleIndex t;
        t.vertex = glm::ivec3(tri[6 * i], tri[6 * i + 1], tri[6 * i + 2]);
        t.normal = glm::ivec3(tri[6 * i + 3], tri[6 * i + 4], tri[6 * i + 5]);
        s->mesh.triangleIndices.push_back(t);
        s->mesh.materialIndexPerTriangle.push_back(0);
    }
    s->mesh.numtriangles = ntri;
    return s;
}
void ref_scene_free(void *h) { RefScene *s = (RefScene *)h; delete s->lightning; delete s; }
void ref_scene_counts(void *h, int *nv, int *nn, int *ntri, int *nmat) {
    RefScene *s = (RefScene *)h;
    *nv = (int)s->mesh.vertices.size(); *nn = (int)s->mesh.normals.size();
    *ntri = (int)s->mesh.triangleIndices.size(); *nmat = (int)s->mesh.materials.size();
}
void ref_scene_arrays(void *h, float *v, float *n, int *tri, int *mat_idx) {
    RefScene *s = (RefScene *)h;
    memcpy(v, s->mesh.vertices.data(), s->mesh.vertices.size() * sizeof(glm::vec3));
    memcpy(n, s->mesh.normals.data(), s->mesh.normals.size() * sizeof(glm::vec3));
    memcpy(tri, s->mesh.triangleIndices.data(), s->mesh.triangleIndices.size() * sizeof(vertex::TriangleIndex));
    memcpy(mat_idx, s->mesh.materialIndexPerTriangle.data(), s->mesh.materialIndexPerTriangle.size() * sizeof(int));
}
// per material: rgbcolor[3], emission[3], spectral_values[K], spectral_emission[K], M[K*K] column-major
void ref_material(void *h, int i, float *rgb, float *emis, float *spec, float *spec_emis, float *M) {
    RefScene *s = (RefScene *)h;
    Material &m = s->mesh.materials[i];
    int K = (int)s->wavelengths.size();
    for (int k = 0; k < 3; k++) { rgb[k] = m.rgbcolor[k]; emis[k] = m.emission[k]; }
    for (int k = 0; k < K; k++) { spec[k] = m.spectral_values[k]; spec_emis[k] = m.spectral_emission[k]; }
    memcpy(M, m.M.data(), sizeof(float) * K * K);
}
// the UV lamp's M is undefined behaviour in the reference (Material.cpp:52-54 indexes a vec3 with i>=3);
// callers overwrite it with the documented restatement before running the gather
void ref_material_set_M(void *h, int i, const float *M) {
    RefScene *s = (RefScene *)h;
    int K = (int)s->wavelengths.size();
    memcpy(s->mesh.materials[i].M.data(), M, sizeof(float) * K * K);
}

// ---- triangle_math.cpp ---------------------------------------------------------------------------
float ref_calculateSurface3(const float *a, const float *b, const float *c) {
    return triangle_math::calculateSurface(glm::vec3(a[0], a[1], a[2]), glm::vec3(b[0], b[1], b[2]), glm::vec3(c[0], c[1], c[2]));
}
float ref_calculateSurface(void *h, int tri) { return triangle_math::calculateSurface((float)tri, ((RefScene *)h)->mesh); }
void ref_calculateCentre(void *h, int tri, float *out) { glm::vec3 c = triangle_math::calculateCentre((float)tri, ((RefScene *)h)->mesh); out[0] = c.x; out[1] = c.y; out[2] = c.z; }
void ref_avgNormal(void *h, int tri, float *out) { glm::vec3 c = triangle_math::avgNormal((float)tri, ((RefScene *)h)->mesh); out[0] = c.x; out[1] = c.y; out[2] = c.z; }
void ref_divideInFourTriangles(void *h, int tri, float *out36) {
    auto r = triangle_math::divideInFourTriangles((float)tri, ((RefScene *)h)->mesh);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) out36[(i * 3 + j) * 3 + k] = r[i][j][k];
}
void ref_uv2xyz(void *h, int tri, float u, float v, float *out) {
    optix::float2 uv = optix::make_float2(u, v);
    optix::float3 p = triangle_math::uv2xyz(tri, uv, ((RefScene *)h)->mesh);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
float ref_calcPointFormfactor(const float *op, const float *on, const float *dp, const float *dn, float surface) {
    vertex::Vertex o = { glm::vec3(op[0], op[1], op[2]), glm::vec3(on[0], on[1], on[2]) };
    vertex::Vertex d = { glm::vec3(dp[0], dp[1], dp[2]), glm::vec3(dn[0], dn[1], dn[2]) };
    return triangle_math::calcPointFormfactor(o, d, surface);
}
// the 16-term loop of OptixPrimeFunctionality::p2pFormfactor (OptixPrimeFunctionality.cpp:133-161, that file
// itself needs OptiX), re-assembled here from the reference's real triangle_math functions, without visibility
float ref_p2pFormfactor_unoccluded(void *h, int originPatch, int destPatch) {
    MeshS &mesh = ((RefScene *)h)->mesh;
    std::vector<std::vector<glm::vec3>> origintriangles = triangle_math::divideInFourTriangles(originPatch, mesh);
    std::vector<std::vector<glm::vec3>> destinationtriangles = triangle_math::divideInFourTriangles(destPatch, mesh);
    std::vector<glm::vec3> originpoints(4), destinationpoints(4);
    glm::vec3 originNormal = triangle_math::avgNormal(originPatch, mesh);
    glm::vec3 destNormal = triangle_math::avgNormal(destPatch, mesh);
    for (int i = 0; i < 4; i++) {
        originpoints[i] = triangle_math::calculateCentre(origintriangles[i]);
        destinationpoints[i] = triangle_math::calculateCentre(destinationtriangles[i]);
    }
    float formfactor = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            formfactor = formfactor + triangle_math::calcPointFormfactor({ originpoints[i], originNormal }, { destinationpoints[j], destNormal },
                triangle_math::calculateSurface(origintriangles[i]) * triangle_math::calculateSurface(destinationtriangles[j]));
    formfactor = formfactor / triangle_math::calculateSurface(originPatch, mesh);
    return formfactor;
}
// ray of OptixPrimeFunctionality.cpp:191-196 from the reference's uv2xyz + the shimmed optix:: operators
void ref_pair_ray(void *h, int row, int col, float u, float v, float *ray6) {
    MeshS &mesh = ((RefScene *)h)->mesh;
    optix::float2 uv = optix::make_float2(u, v);
    optix::float3 origin = triangle_math::uv2xyz(row, uv, mesh);
    optix::float3 dest = triangle_math::uv2xyz(col, uv, mesh);
    optix::float3 o = origin + optix::normalize(dest - origin) * 0.000001f;
    optix::float3 d = optix::normalize(dest - origin);
    ray6[0] = o.x; ray6[1] = o.y; ray6[2] = o.z; ray6[3] = d.x; ray6[4] = d.y; ray6[5] = d.z;
}

// Camera::gen_rays_for_screen (Camera.h:54-80): out = W*H*samples rays of 6 floats
int ref_camera_rays(int width, int height, int supersampling, int antialiasing, float *out) {
    Camera cam(width, height, supersampling);
    std::vector<optix::float3> rays;
    cam.gen_rays_for_screen(rays, antialiasing != 0);
    memcpy(out, rays.data(), rays.size() * sizeof(optix::float3));
    return (int)(rays.size() / 2);
}

// ---- Lightning.h -----------------------------------------------------------------------------------
// triplets (row, col, value as double) are what calculateAllVisibility returns (OptixPrimeFunctionality.cpp:214-217)
void ref_set_triplets(void *h, const int *rows, const int *cols, const double *vals, int64_t nnz) {
    RefScene *s = (RefScene *)h;
    s->optixP.triplets.clear();
    s->optixP.triplets.reserve(nnz);
    for (int64_t i = 0; i < nnz; i++) s->optixP.triplets.push_back(Tripl(rows[i], cols[i], vals[i]));
}
// Lightning::get_lightning(method, ...) (Lightning.h:446-457).  The constructors call converge_lightning();
// returns the pass count that took.
int ref_lightning_create(void *h, int method, float emission_value) {
    RefScene *s = (RefScene *)h;
    delete s->lightning;
    s->method = method; s->emission_value = emission_value;
    quiet(true);
    s->lightning = Lightning::get_lightning(method, s->mesh, s->optixP, s->emission_value, s->wavelengths, true, nullptr);
    quiet(false);
    return s->lightning->numpasses;
}
void ref_lightning_reset(void *h) { quiet(true); ((RefScene *)h)->lightning->reset(); quiet(false); }
int ref_lightning_increment(void *h) { RefScene *s = (RefScene *)h; quiet(true); s->lightning->increment_lightpass(); quiet(false); return s->lightning->numpasses; }
int ref_lightning_converge(void *h) { RefScene *s = (RefScene *)h; quiet(true); s->lightning->converge_lightning(); quiet(false); return s->lightning->numpasses; }
int ref_lightning_bands(void *h) {
    RefScene *s = (RefScene *)h;
    return s->method == 0 ? 1 : (s->method == 1 ? 3 : (int)s->wavelengths.size());
}
// B ("lightningvalues") and residual, band-major K x N
void ref_lightning_read(void *h, float *B, float *residual) {
    RefScene *s = (RefScene *)h;
    int N = s->mesh.numtriangles;
    if (s->method == 0) {
        BWLightning *l = (BWLightning *)s->lightning;
        memcpy(B, l->lightningvalues.data(), sizeof(float) * N);
        memcpy(residual, l->residualvector.data(), sizeof(float) * N);
    } else if (s->method == 1) {
        RGBLightning *l = (RGBLightning *)s->lightning;
        for (int k = 0; k < 3; k++) {
            memcpy(B + (size_t)k * N, l->lightningvalues[k].data(), sizeof(float) * N);
            memcpy(residual + (size_t)k * N, l->residualvector[k].data(), sizeof(float) * N);
        }
    } else {
        SpectralLightning *l = (SpectralLightning *)s->lightning;
        for (int k = 0; k < l->numsamples; k++) {
            memcpy(B + (size_t)k * N, l->lightningvalues[k].data(), sizeof(float) * N);
            memcpy(residual + (size_t)k * N, l->residualvector[k].data(), sizeof(float) * N);
        }
    }
}
void ref_lightning_color(void *h, int patch, float *rgb) {
    glm::vec3 c = ((RefScene *)h)->lightning->get_color_of_patch(patch);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
// one gather pass only, timed by the caller (the CPU baseline): increment_lightpass without the colour cache for
// the spectral class is private API (increment_light_fluorescent, Lightning.h:196-226)
void ref_lightning_pass_only(void *h) {
    RefScene *s = (RefScene *)h;
    if (s->method == 2) ((SpectralLightning *)s->lightning)->increment_light_fluorescent();
    else if (s->method == 1) ((RGBLightning *)s->lightning)->increment_lightpass();
    else { BWLightning *l = (BWLightning *)s->lightning; l->residualvector = l->RadMat * l->residualvector; l->lightningvalues = l->lightningvalues + l->residualvector; l->numpasses++; }
}
int64_t ref_radmat_nnz(void *h) { return (int64_t)((RefScene *)h)->lightning->RadMat.nonZeros(); }

} // extern "C"
