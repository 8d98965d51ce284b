// rt_math.h -- small float3 / colour helpers shared by the CUDA kernels and the host-side scene code.
//
// Arithmetic contract: the translation units that include this header are compiled WITHOUT floating-point
// contraction (nvcc --fmad=false, g++ -ffp-contract=off), IEEE division and square root.  Every expression is
// written in the operation order of the reference (tp2/src/vec.cpp, color.cpp) so that hit distances, barycentrics
// and ray origins come out bit-identical to the reference built without FMA contraction.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__
#else
#define RT_HD inline
#define RT_D inline
#endif

namespace rtb {

struct V3 {
    float x, y, z;
};

RT_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }             // vec.cpp:58-61,68-71,88-91
RT_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }             // vec.cpp:37-40,93-96
RT_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }                                  // vec.cpp:63-66
RT_HD V3 operator*(float k, V3 a) { return v3(k * a.x, k * a.y, k * a.z); }                // vec.cpp:42-45,98-101
RT_HD float dot(V3 u, V3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }                  // vec.cpp:164-167
RT_HD V3 cross(V3 u, V3 v)                                                                 // vec.cpp:156-162
{
    return v3((u.y * v.z) - (u.z * v.y), (u.z * v.x) - (u.x * v.z), (u.x * v.y) - (u.y * v.x));
}
RT_HD float length2(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }                    // vec.cpp:174-177
RT_HD V3 normalize(V3 v) { float kk = 1.0f / sqrtf(length2(v)); return kk * v; }           // vec.cpp:150-154,169-172

struct Col {
    float r, g, b;
};
RT_HD Col col(float r, float g, float b) { Col c; c.r = r; c.g = g; c.b = b; return c; }
RT_HD Col col(float v) { return col(v, v, v); }
RT_HD Col operator+(Col a, Col b) { return col(a.r + b.r, a.g + b.g, a.b + b.b); }         // color.cpp:47-50
RT_HD Col operator*(Col a, Col b) { return col(a.r * b.r, a.g * b.g, a.b * b.b); }         // color.cpp:62-65
RT_HD Col operator*(Col c, float k) { return col(c.r * k, c.g * k, c.b * k); }             // color.cpp:67-75
RT_HD Col operator/(Col a, Col b) { return col(a.r / b.r, a.g / b.g, a.b / b.b); }         // color.cpp:77-80

RT_HD float clamp01(float x)                                                               // std::clamp(x, 0.f, 1.f)
{
    return x < 0.0f ? 0.0f : (1.0f < x ? 1.0f : x);
}

// Row-major 4x4 (Transform::m[row][col], tp2/src/mat.h).
struct M4 {
    float m[4][4];
};

// Transform::operator()(const Point&) -- mat.cpp:83-100
RT_HD V3 xform_point(const M4& t, V3 p)
{
    float xt = t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * p.z + t.m[0][3];
    float yt = t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * p.z + t.m[1][3];
    float zt = t.m[2][0] * p.x + t.m[2][1] * p.y + t.m[2][2] * p.z + t.m[2][3];
    float wt = t.m[3][0] * p.x + t.m[3][1] * p.y + t.m[3][2] * p.z + t.m[3][3];
    if (wt == 1.0f) return v3(xt, yt, zt);
    float w = 1.0f / wt;
    return v3(xt * w, yt * w, zt * w);
}

// ImageUtils::gkit_color_to_Qt_ARGB32_uint + qRgb -- imageUtils.h:149-152 (truncation, not rounding)
RT_HD uint32_t quantise_argb(Col c)
{
    int r = (int)(c.r * 255), g = (int)(c.g * 255), b = (int)(c.b * 255);
    return 0xff000000u | (((uint32_t)r & 0xffu) << 16) | (((uint32_t)g & 0xffu) << 8) | ((uint32_t)b & 0xffu);
}

// Per-pixel seed of the shared xorshift32 stream (murmur3 finaliser, never zero).  Same as rt_pixel_seed().
RT_HD uint32_t pixel_seed(uint32_t pixel_index, uint32_t rng_seed)
{
    uint32_t x = pixel_index ^ rng_seed;
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x | 1u;
}

// XorShiftGenerator -- renderer/xorshift.h:37-65
struct XorShift32 {
    uint32_t state;
    RT_HD uint32_t next()
    {
        uint32_t x = state;
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        state = x;
        return x;
    }
    // get_rand() / (float)UINT32_MAX * 2 - 1 ; (float)UINT32_MAX rounds to 2^32
    RT_HD float bilateral() { return (float)next() / 4294967296.0f * 2 - 1; }
};

} // namespace rtb
