// ref_cuda_wrap.cu -- extern "C" doorway to the reference's only first-party CUDA kernel, compiled UNMODIFIED from
// /root/reference/visual studio/parallellism.cu (calculateRow / p2pFormfactor, :91-227) for sm_100a.
// TEST INFRASTRUCTURE ONLY: the GPU-side oracle for the unoccluded form factors (device arithmetic, double pi) and
// the incumbent kernel that gets timed next to ours.  Two builds (see Makefile): default nvcc flags, and -fmad=false,
// whose results are bit-comparable with the operation-order restatement in daisy_oracle.c / formfactor.cu.
#include <vector>
#include <cstring>
#include <chrono>
#include "parallellism.cuh"

extern "C" {
// out[r*N + c] = F_unoccluded(r -> c) exactly as parallellism::runCalculateRadiosityMatrix returns it (Tripl.m_value);
// returns wall seconds of the reference driver (managed-memory chunks + host copies included, as in the reference)
double ref_cuda_runCalculateRadiosityMatrix(const float *v, int nv, const float *n, int nn, const int *tri, int ntri, double *out) {
    SimpleMesh mesh;
    mesh.numtriangles = ntri;
    for (int i = 0; i < nv; i++) mesh.vertices.push_back(glm::vec3(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
    for (int i = 0; i < nn; i++) mesh.normals.push_back(glm::vec3(n[3 * i], n[3 * i + 1], n[3 * i + 2]));
    for (int i = 0; i < ntri; i++) {
        vertex::TriangleIndex t;
        t.vertex = glm::ivec3(tri[6 * i], tri[6 * i + 1], tri[6 * i + 2]);
        t.normal = glm::ivec3(tri[6 * i + 3], tri[6 * i + 4], tri[6 * i + 5]);
        mesh.triangleIndices.push_back(t);
    }
    auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<parallellism::Tripl> res = parallellism::runCalculateRadiosityMatrix(mesh);
    auto t1 = std::chrono::high_resolution_clock::now();
    size_t want = (size_t)ntri * ntri;
    for (size_t i = 0; i < want && i < res.size(); i++) out[(size_t)res[i].m_row * ntri + res[i].m_col] = res[i].m_value;
    return std::chrono::duration<double>(t1 - t0).count();
}
}
