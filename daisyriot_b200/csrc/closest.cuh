// closest.cuh -- closest-hit walk of the LBVH for one ray (shared by k_closest and the traceScreen kernel).
#pragma once
#include "daisy_common.cuh"

// (t, triangleId) lexicographic minimum over all triangles the watertight test accepts with finite t > 0; miss => t = -1,
// triangleId = -1.  Near child first, far child pushed; the result does not depend on the visiting order.
__device__ __forceinline__ daisy_hit closest_hit(const BvhNode *__restrict__ nodes, const TriVerts *__restrict__ tv, int root, int ntri, f3 o, f3 d) {
    daisy_hit best; best.t = -1.0f; best.triangleId = -1; best.u = 0.f; best.v = 0.f;
    if (ntri <= 0) return best;
    WRay w = wray_setup(o, d);
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[64];
    int sp = 0;
    int cur = root;
    while (true) {
        if (cur < 0) {
            int k = ~cur;
            TriVerts t = tv[k];
            float tt, uu, vv;
            if (wray_tri(w, xyz(t.a), xyz(t.b), xyz(t.c), tt, uu, vv)) {
                if (best.triangleId < 0 || tt < best.t || (tt == best.t && k < best.triangleId)) {
                    best.t = tt; best.triangleId = k; best.u = uu; best.v = vv;
                }
            }
            if (sp == 0) break;
            cur = stack[--sp];
            continue;
        }
        BvhNode nd = nodes[cur];
        float tmax = best.triangleId >= 0 ? best.t : INFINITY;
        float tl, tr;
        bool hl = ray_box(o, inv, nd.a.x, nd.a.y, nd.a.z, nd.a.w, nd.b.x, nd.b.y, tmax, tl);
        bool hr = ray_box(o, inv, nd.b.z, nd.b.w, nd.c.x, nd.c.y, nd.c.z, nd.c.w, tmax, tr);
        if (hl && hr) {
            int nearc = nd.d.x, farc = nd.d.y;
            if (tr < tl) { nearc = nd.d.y; farc = nd.d.x; }
            DZ_ASSERT(sp < 64);
            stack[sp++] = farc;
            cur = nearc;
        } else if (hl) cur = nd.d.x;
        else if (hr) cur = nd.d.y;
        else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    return best;
}
