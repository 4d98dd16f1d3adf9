// Stand-in for the OptiX SDK header the reference includes (<OptiX_world.h>, closed SDK, absent here).
// Only the optix::float2/float3 vocabulary the reference's host geometry code touches is provided, with the
// arithmetic of optixu_math_namespace.h as documented: dot = x*x'+y*y'+z*z', normalize = v * (1/sqrtf(dot)).
// TEST INFRASTRUCTURE ONLY (lets /root/reference sources compile unmodified into oracle/_ref).
#pragma once
#include <cmath>
#ifndef M_PIf
#define M_PIf 3.14159265358979323846f
#endif
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
static inline float3 normalize(const float3 &v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
}
