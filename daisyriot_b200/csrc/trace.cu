// trace.cu -- camera ray cast with the shading epilogue fused behind the closest-hit traversal.
//
// Replaces (reference "visual studio/"):
//   OptixPrimeFunctionality::traceScreen           OptixPrimeFunctionality.cpp:83-131  (optixQuery + per-pixel host loop)
//   triangle_math::isFacingBack                    triangle_math.cpp:76-86
//   Drawer::interpolate                            Drawer.cpp:161-186  (per-vertex colour = mean over trianglesPerVertex)
// The reference ships width*height*samples rays to OptiX Prime, copies 16-byte hits back and shades on one CPU thread.
// Here one thread owns a pixel: it traces the pixel's samples in order, tests isFacingBack, interpolates the per-vertex
// colours with the hit's barycentrics, averages and clamps -- only the finished RGB frame (and, if asked for, the hit
// records the picking code wants) leaves the device.
#include "daisy_common.cuh"
#include "closest.cuh"

// per-vertex colour: sum of get_color_of_patch over trianglesPerVertex[v] in ascending triangle order (MeshS.cpp:107-109
// pushes them in that order), divided component-wise by the count (Drawer.cpp:169-180)
__global__ void k_vertex_colors(int nv, const int *__restrict__ off, const int *__restrict__ adj, const float *__restrict__ patch_rgb,
                                float *__restrict__ vcol) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    float r = 0.f, g = 0.f, b = 0.f;
    const int e0 = off[v], e1 = off[v + 1];
    for (int e = e0; e < e1; e++) {
        const float *c = patch_rgb + 3 * (size_t)adj[e];
        r = fa(r, c[0]); g = fa(g, c[1]); b = fa(b, c[2]);
    }
    const float n = (float)(e1 - e0);
    const bool any = e1 > e0; // a vertex no triangle uses is never looked up
    vcol[3 * (size_t)v] = any ? fd(r, n) : 0.f;
    vcol[3 * (size_t)v + 1] = any ? fd(g, n) : 0.f;
    vcol[3 * (size_t)v + 2] = any ? fd(b, n) : 0.f;
}

__global__ void __launch_bounds__(128) k_trace_shade(const BvhNode *__restrict__ nodes, const TriVerts *__restrict__ tv, const PatchGeom *__restrict__ geom,
                                                     const int *__restrict__ tri, int root, int ntri, int npix, int samples,
                                                     const float *__restrict__ rays, f3 eye, const float *__restrict__ vcol,
                                                     const float *__restrict__ patch_rgb, int interpolate, float *__restrict__ out_rgb,
                                                     daisy_hit *__restrict__ hits_out) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    float cr = 0.f, cg = 0.f, cb = 0.f;
    for (int i = 0; i < samples; i++) {
        const size_t ri = (size_t)pix * samples + i;
        const f3 o = mk3(rays[6 * ri], rays[6 * ri + 1], rays[6 * ri + 2]);
        const f3 d = mk3(rays[6 * ri + 3], rays[6 * ri + 4], rays[6 * ri + 5]);
        const daisy_hit h = closest_hit(nodes, tv, root, ntri, o, d);
        if (hits_out) hits_out[ri] = h;
        if (!(h.t > 0.0f)) continue;
        const int k = h.triangleId;
        // isFacingBack(eye, k): dot(normalize(centre - eye), avgNormal) >= 0 ; centre = (a + b + c) / 3 per component
        const TriVerts T = tv[k];
        const f3 s = e_add(e_add(xyz(T.a), xyz(T.b)), xyz(T.c));
        const f3 centre = mk3(fd(s.x, 3.0f), fd(s.y, 3.0f), fd(s.z, 3.0f));
        const f3 desty = e_normalize(e_sub(centre, eye));
        if (e_dot(desty, xyz(geom[k].n)) >= 0.0f) continue;
        if (interpolate) {
            const int *t = tri + 6 * (size_t)k;
            const float *a = vcol + 3 * (size_t)t[0], *b = vcol + 3 * (size_t)t[1], *c = vcol + 3 * (size_t)t[2];
            // w = 1 - u - v ; lightningvalue = u * a + v * b + w * c                       Drawer.cpp:182-184
            const float w = fs(fs(1.0f, h.u), h.v);
            cr = fa(cr, fa(fa(fm(h.u, a[0]), fm(h.v, b[0])), fm(w, c[0])));
            cg = fa(cg, fa(fa(fm(h.u, a[1]), fm(h.v, b[1])), fm(w, c[1])));
            cb = fa(cb, fa(fa(fm(h.u, a[2]), fm(h.v, b[2])), fm(w, c[2])));
        } else {
            const float *c = patch_rgb + 3 * (size_t)k; // materials[materialIndexPerTriangle[k]].rgbcolor, expanded per patch
            cr = fa(cr, c[0]); cg = fa(cg, c[1]); cb = fa(cb, c[2]);
        }
    }
    const float ns = (float)samples;
    out_rgb[3 * (size_t)pix] = fminf(fmaxf(fd(cr, ns), 0.f), 1.f);
    out_rgb[3 * (size_t)pix + 1] = fminf(fmaxf(fd(cg, ns), 0.f), 1.f);
    out_rgb[3 * (size_t)pix + 2] = fminf(fmaxf(fd(cb, ns), 0.f), 1.f);
}

extern "C" int daisy_trace_screen(daisy_ctx *ctx, int width, int height, int samples, const float *rays6, const float *eye3,
                                  const float *patch_rgb, int interpolate, float *out_rgb, daisy_hit *hits_out) {
    DZ_REQUIRE(ctx && rays6 && eye3 && patch_rgb && out_rgb, DAISY_E_INVALID, "daisy_trace_screen: null argument");
    DZ_REQUIRE(width >= 0 && height >= 0 && samples >= 1, DAISY_E_INVALID, "daisy_trace_screen: bad frame size or sample count");
    const int64_t npix64 = (int64_t)width * height;
    DZ_REQUIRE(npix64 * samples <= 0x7fffffffLL, DAISY_E_INVALID, "daisy_trace_screen: frame too large");
    const int npix = (int)npix64;
    if (npix == 0) return DAISY_OK;
    DZ_CUDA(cudaSetDevice(ctx->device));
    nvtxRangePushA("daisy_trace_screen");
    cudaStream_t st = ctx->stream;
    const size_t nrays = (size_t)npix * samples;
    const int N = ctx->N;
    float *d_rays = nullptr, *d_rgb = nullptr, *d_vcol = nullptr, *d_out = nullptr;
    daisy_hit *d_hits = nullptr;
    cudaError_t e = dz_scratch(ctx, 0, sizeof(float) * 6 * nrays, (void **)&d_rays);
    if (e == cudaSuccess) e = dz_scratch(ctx, 2, sizeof(float) * 3 * (size_t)(N > 0 ? N : 1), (void **)&d_rgb);
    if (e == cudaSuccess) e = dz_scratch(ctx, 3, sizeof(float) * 3 * (size_t)(ctx->nv > 0 ? ctx->nv : 1), (void **)&d_vcol);
    if (e == cudaSuccess) e = dz_scratch(ctx, 4, sizeof(float) * 3 * (size_t)npix, (void **)&d_out);
    if (e == cudaSuccess && hits_out) e = dz_scratch(ctx, 1, sizeof(daisy_hit) * nrays, (void **)&d_hits);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rays, rays6, sizeof(float) * 6 * nrays, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && N > 0) e = cudaMemcpyAsync(d_rgb, patch_rgb, sizeof(float) * 3 * (size_t)N, cudaMemcpyHostToDevice, st);
    f3 eye; eye.x = eye3[0]; eye.y = eye3[1]; eye.z = eye3[2];
    if (e == cudaSuccess) {
        if (interpolate && ctx->nv > 0) k_vertex_colors<<<(ctx->nv + 255) / 256, 256, 0, st>>>(ctx->nv, ctx->d_vadj_off, ctx->d_vadj, d_rgb, d_vcol);
        k_trace_shade<<<(npix + 127) / 128, 128, 0, st>>>(ctx->d_nodes, ctx->d_triverts, ctx->d_geom, ctx->d_tri, ctx->root, N, npix, samples, d_rays,
                                                          eye, d_vcol, d_rgb, interpolate, d_out, d_hits);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_rgb, d_out, sizeof(float) * 3 * (size_t)npix, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && hits_out) e = cudaMemcpyAsync(hits_out, d_hits, sizeof(daisy_hit) * nrays, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    nvtxRangePop();
    if (e != cudaSuccess) { daisy_set_error("daisy_trace_screen: %s", cudaGetErrorString(e)); return DAISY_E_CUDA; }
    return DAISY_OK;
}
