// ssao_device.h -- Renderer::post_process_ssao_SIMD (renderer/renderer.cpp:1229-1434), the per-pixel functions: SURVEY.md section 8(f)3.
// (Plain inline code like rt_device.h: ssao.cuh wraps it in kernels, tests/hostsim compiles it with g++.)
//
// The reference estimates screen-space ambient occlusion from two G-buffers that ray_trace() fills for every pixel whose
// primary ray found something (renderer.cpp:1104-1111): z = -(ray.origin.z + ray.direction.z * t) and the shading normal
// (after normal mapping).  Per pixel it draws ssao_sample_count points in the hemisphere around the normal, projects each
// with the camera's perspective matrix, looks the z-buffer up at the pixel the point lands on and counts the samples that
// lie behind the geometry found there (within ssao_radius); the counts are blurred with a 7x7 box and darken the image
// BEFORE the SSAA resolve.  The reference runs 8 pixels per AVX2 register; every lane's float operations are restated here in
// the same order (one explicit fused multiply-add per _mm256_fmadd_ps of SIMD/m256Point.cpp:3-31, everything else unfused:
// this translation unit is compiled with --fmad=false), and the columns its last partial 8-pixel group leaves over use
// the formulas of its scalar loop (renderer.cpp:1358-1407: double arithmetic + truncation for the sampled pixel, unsigned
// random numbers).  Here the G-buffers are written by the shade stage (kernels.cuh: shade_prepare), one thread per pixel
// counts the occluded samples, a second kernel blurs and applies.
//
// Random numbers: the reference seeds its generators from std::rand() and the OpenMP thread number and draws in processing
// order.  The shared reproducible stream here is one xorshift32 per pixel, state = rt_pixel_seed(py * W' + px,
// rng_seed + kSsaoSeedOffset), four draws per sample in the reference's order (x, y, z, length) -- the discipline of
// oracle/oracle.cpp's SSAO_RNG_PER_PIXEL, whose arithmetic is pinned bit-exactly to the compiled reference in its
// reference-order mode (tests/test_oracle_vs_reference.py).
#pragma once

#include "rt_device.h"

namespace rtb {

constexpr uint32_t kSsaoSeedOffset = 0x9e3779b9u;

struct SsaoView {
    int32_t rw, rh;
    const float* z;              // _z_buffer, +inf where the primary ray found nothing
    const V3* n;                 // _normal_buffer
    M4 proj;                     // Camera::_perspective_proj_mat
    float aspect;
    float fov_mult_simd;         // (float)std::tan(fov / 2 / 180 * M_PI), renderer.cpp:1249 (host, double)
    float fov_mult_scalar;       // std::tan(radians(fov / 2)) in float, renderer.cpp:1369 (host)
    int32_t samples;
    float radius, amount;
    uint32_t rng_seed;
};

RT_DEV int32_t ssao_imin(int32_t a, int32_t b) { return a < b ? a : b; }
RT_DEV int32_t ssao_imax(int32_t a, int32_t b) { return a > b ? a : b; }

RT_DEV uint32_t xs_next(uint32_t& st)                           // xorshift.h:13-22,43-52
{
    uint32_t x = st;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    return st = x;
}
// __m256_XorShiftGenerator::get_rand_bilateral / get_rand_lateral -- xorshift.h:24-32 (_mm256_cvtepi32_ps is SIGNED;
// (float)INT32_MAX rounds to 2^31)
RT_DEV float simd_bilateral(uint32_t& st) { return (float)(int32_t)xs_next(st) / 2147483648.0f; }
RT_DEV float simd_lateral(uint32_t& st) { return ((float)(int32_t)xs_next(st) / 2147483648.0f + 1.0f) * 0.5f; }
// XorShiftGenerator::get_rand_bilateral / get_rand_lateral -- xorshift.h:54-62 ((float)UINT32_MAX rounds to 2^32)
RT_DEV float scalar_bilateral(uint32_t& st) { return (float)xs_next(st) / 4294967296.0f * 2 - 1; }
RT_DEV float scalar_lateral(uint32_t& st) { return (float)xs_next(st) / 4294967296.0f; }

// _mm256_cvtps_epi32: round to nearest even; NaN and out-of-range give the x86 "integer indefinite" 0x80000000
RT_DEV int32_t cvtps_epi32(float f)
{
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return (int32_t)0x80000000;
#if defined(__CUDACC__)
    return __float2int_rn(f);
#else
    return (int32_t)lrintf(f);                                                    // the default rounding mode: to nearest even
#endif
}
// (int) of a double as x86 does it (cvttsd2si): truncation, indefinite when out of range
RT_DEV int32_t cvttsd_epi32(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int32_t)0x80000000;
#if defined(__CUDACC__)
    return __double2int_rz(v);
#else
    return (int32_t)v;
#endif
}

// One lane of the AVX2 loop body, renderer.cpp:1283-1355
RT_DEV int ssao_simd_pixel(const SsaoView& f, int x, int y, uint32_t& rng)
{
    const float view_z = f.z[(size_t)y * f.rw + x];
    float y_ndc = (float)y / (float)f.rh;
    y_ndc = y_ndc * 2.0f;
    y_ndc = y_ndc - 1.0f;
    const float xs = (float)(x & 7) + (float)(x & ~7);
    float x_ndc = xs / (float)f.rw;
    x_ndc = x_ndc * 2.0f;
    x_ndc = x_ndc - 1.0f;
    const float view_ray_x = x_ndc * (f.fov_mult_simd * f.aspect);
    const float view_ray_y = y_ndc * f.fov_mult_simd;
    const V3 P = v3(view_z * view_ray_x, view_z * view_ray_y, view_z * -1.0f);
    const V3 nb = f.n[(size_t)y * f.rw + x];
    const float n_inv = 1.0f / sqrtf(nb.x * nb.x + (nb.y * nb.y + nb.z * nb.z));    // _mm256_length adds x + (y + z)
    const V3 normal = v3(nb.x * n_inv, nb.y * n_inv, nb.z * n_inv);
    int occlusion = 0;
    for (int i = 0; i < f.samples; i++) {
        const float rx = simd_bilateral(rng), ry = simd_bilateral(rng), rz = simd_bilateral(rng);
        const float r_inv = 1.0f / sqrtf(rx * rx + (ry * ry + rz * rz));
        V3 rs = v3(rx * r_inv, ry * r_inv, rz * r_inv);
        const float k = simd_lateral(rng) + 0.0001f;
        rs = v3(rs.x * k, rs.y * k, rs.z * k);
        rs = v3(rs.x * f.radius, rs.y * f.radius, rs.z * f.radius);
        rs = v3(rs.x + P.x, rs.y + P.y, rs.z + P.z);
        const V3 vd = v3(rs.x - P.x, rs.y - P.y, rs.z - P.z);
        const float d = vd.x * normal.x + (vd.y * normal.y + vd.z * normal.z);       // _mm256_dot_product: x + (y + z)
        const float flip = d < 0.0f ? 1.0f : 0.0f;
        const V3 back = v3((P.x - rs.x) * 2.0f, (P.y - rs.y) * 2.0f, (P.z - rs.z) * 2.0f);
        rs = v3(rs.x + back.x * flip, rs.y + back.y * flip, rs.z + back.z * flip);
        const float (*m)[4] = f.proj.m;                                              // __m256Point::transform
        const float xt = RT_FMA(m[0][0], rs.x, RT_FMA(m[0][1], rs.y, RT_FMA(m[0][2], rs.z, m[0][3])));
        const float yt = RT_FMA(m[1][0], rs.x, RT_FMA(m[1][1], rs.y, RT_FMA(m[1][2], rs.z, m[1][3])));
        const float wt = RT_FMA(m[3][0], rs.x, RT_FMA(m[3][1], rs.y, RT_FMA(m[3][2], rs.z, m[3][3])));
        const float w = 1.0f / wt;
        const float ndc_x = xt * w, ndc_y = yt * w;
        int32_t px = cvtps_epi32(((ndc_x + 1.0f) * 0.5f) * (float)f.rw);
        int32_t py = cvtps_epi32(((ndc_y + 1.0f) * 0.5f) * (float)f.rh);
        px = ssao_imax(ssao_imin(px, cvtps_epi32((float)f.rw - 1.0f)), 0);
        py = ssao_imax(ssao_imin(py, cvtps_epi32((float)f.rh - 1.0f)), 0);
        const float sample_geometry_depth = -1.0f * f.z[(size_t)px + (size_t)py * f.rw];
        const bool in_range = fabsf(sample_geometry_depth - P.z) <= f.radius;        // _CMP_LE_OQ
        const bool behind = rs.z < sample_geometry_depth;                             // _CMP_LT_OQ
        if (in_range && behind) occlusion++;
    }
    return occlusion;
}

// The scalar loop of the left-over columns, renderer.cpp:1358-1407
RT_DEV int ssao_scalar_pixel(const SsaoView& f, int x, int y, uint32_t& rng)
{
    const float x_ndc = (float)x / (float)f.rw * 2 - 1;
    const float y_ndc = (float)y / (float)f.rh * 2 - 1;
    const float view_z = f.z[(size_t)y * f.rw + x];
    const float view_ray_x = x_ndc * f.aspect * f.fov_mult_scalar;
    const float view_ray_y = y_ndc * f.fov_mult_scalar;
    const V3 P = v3(view_ray_x * view_z, view_ray_y * view_z, -view_z);
    const V3 normal = normalize(f.n[(size_t)y * f.rw + x]);
    int occlusion = 0;
    for (int i = 0; i < f.samples; i++) {
        const float rx = scalar_bilateral(rng);
        const float ry = scalar_bilateral(rng);
        const float rz = scalar_bilateral(rng);
        V3 rs = normalize(v3(rx, ry, rz));
        rs = (scalar_lateral(rng) + 0.0001f) * rs;
        rs = f.radius * rs;
        rs = rs + P;
        if (dot(rs - P, normal) < 0) rs = rs + 2.0f * (P - rs);
        const V3 ndc = xform_point(f.proj, rs);
        int px = cvttsd_epi32((double)(ndc.x + 1) * 0.5 * (double)f.rw);
        int py = cvttsd_epi32((double)(ndc.y + 1) * 0.5 * (double)f.rh);
        px = ssao_imin(ssao_imax(0, px), f.rw - 1);
        py = ssao_imin(ssao_imax(0, py), f.rh - 1);
        const float sample_geometry_depth = -f.z[(size_t)py * f.rw + px];
        if (fabsf(sample_geometry_depth - P.z) > f.radius) continue;
        if (rs.z < sample_geometry_depth) occlusion++;
    }
    return occlusion;
}


// What k_ssao_occlusion computes for pixel i: the occluded-sample count (0 for background pixels, the reference's infinity_mask)
RT_DEV int ssao_count_pixel(const SsaoView& f, size_t i)
{
    if (f.z[i] == INFINITY) return 0;
    const int leftover = f.rw % 8;
    const int x = (int)(i % (size_t)f.rw), y = (int)(i / (size_t)f.rw);
    uint32_t st = pixel_seed((uint32_t)i, f.rng_seed + kSsaoSeedOffset);
    return x < f.rw - leftover ? ssao_simd_pixel(f, x, y, st) : ssao_scalar_pixel(f, x, y, st);
}

// What k_ssao_apply does to pixel i: the 7x7 box blur of the counts, applied to the image (renderer.cpp:1411-1431).
// QColor(int, int, int) built from float expressions truncates.
RT_DEV void ssao_apply_pixel(const SsaoView& f, const int* ao, uint32_t* image, size_t i)
{
    const int half = 3;
    const int x = (int)(i % (size_t)f.rw), y = (int)(i / (size_t)f.rw);
    if (x < half || y < half || x >= f.rw - half || y >= f.rh - half) return;
    if (f.z[i] == INFINITY) return;
    int sum = 0;
    for (int oy = -half; oy <= half; oy++)
        for (int ox = -half; ox <= half; ox++) sum += ao[(size_t)(y + oy) * f.rw + x + ox];
    const float mult = 1 - ((float)sum / (float)49 / (float)f.samples * f.amount);
    const uint32_t c = image[i];
    const int r = (int)((c >> 16) & 0xffu), g = (int)((c >> 8) & 0xffu), b = (int)(c & 0xffu);
    const int nr = (int)((float)r * mult), ng = (int)((float)g * mult), nb = (int)((float)b * mult);
    image[i] = 0xff000000u | (((uint32_t)nr & 0xffu) << 16) | (((uint32_t)ng & 0xffu) << 8) | ((uint32_t)nb & 0xffu);
}

} // namespace rtb
