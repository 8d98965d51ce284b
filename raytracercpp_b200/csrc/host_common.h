// host_common.h -- host-side helpers of the C ABI that do not touch CUDA: settings defaults and validation,
// tile ownership.  Shared by rtb200.cu and by the single-threaded kernel emulation under tests/hostsim.
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/rtb200.h"
#include "rt_math.h"

namespace rtb {

// Perspective(fov, aspect, znear, zfar) -- mat.cpp:307-319 with radians() of mat.cpp:13-16, row-major.
inline M4 perspective_matrix(float fov, float aspect, float znear, float zfar)
{
    const float rad = ((float)3.14159265358979323846 / 180) * fov;
    const float itan = 1 / tanf(rad * 0.5f);
    const float id = 1 / (znear - zfar);
    M4 p;
    memset(&p, 0, sizeof(p));
    p.m[0][0] = itan / aspect;
    p.m[1][1] = itan;
    p.m[2][2] = (zfar + znear) * id;
    p.m[2][3] = 2.f * zfar * znear * id;
    p.m[3][2] = -1;
    return p;
}

// Transform::inverse() -- mat.cpp:378-447: Gauss-Jordan elimination with full pivoting, in place.  The order of
// the float operations follows the reference so that Camera::_perspective_proj_mat_inv comes out bit-identical.
inline M4 invert_matrix(const M4& src)
{
    M4 a = src;
    int pivot_row[4], pivot_col[4], done[4] = {0, 0, 0, 0};
    for (int step = 0; step < 4; step++) {
        int r = 0, c = 0;
        float largest = 0.f;
        for (int j = 0; j < 4; j++) {
            if (done[j] == 1) continue;
            for (int k = 0; k < 4; k++)
                if (done[k] == 0 && fabsf(a.m[j][k]) >= largest) { largest = fabsf(a.m[j][k]); r = j; c = k; }
        }
        done[c]++;
        if (r != c)
            for (int k = 0; k < 4; k++) { float tmp = a.m[r][k]; a.m[r][k] = a.m[c][k]; a.m[c][k] = tmp; }
        pivot_row[step] = r;
        pivot_col[step] = c;
        const float inv_pivot = 1.f / a.m[c][c];
        a.m[c][c] = 1.f;
        for (int k = 0; k < 4; k++) a.m[c][k] *= inv_pivot;
        for (int j = 0; j < 4; j++) {
            if (j == c) continue;
            const float factor = a.m[j][c];
            a.m[j][c] = 0;
            for (int k = 0; k < 4; k++) a.m[j][k] -= a.m[c][k] * factor;
        }
    }
    for (int step = 3; step >= 0; step--)
        if (pivot_row[step] != pivot_col[step])
            for (int k = 0; k < 4; k++) {
                float tmp = a.m[k][pivot_row[step]];
                a.m[k][pivot_row[step]] = a.m[k][pivot_col[step]];
                a.m[k][pivot_col[step]] = tmp;
            }
    return a;
}

constexpr int kMaxRecursionDepth = 8;             // the device stack of k_reflect is sized for this

inline void default_settings(RtSettings* s)
{
    memset(s, 0, sizeof(*s));                     // rendererSettings.h:29-102
    s->image_width = 1024;
    s->image_height = 1024;
    s->ssaa_factor = 2;
    s->shading_method = RT_SHADING;
    s->max_recursion_depth = 5;
    s->enable_bvh = 1;
    s->bvh_max_depth = 12;
    s->bvh_leaf_object_count = 40;
    s->enable_ambient = s->enable_diffuse = s->enable_specular = s->enable_emissive = 1;
    s->rough_reflections_sample_count = 3;
    s->displacement_mapping_strength = 0.02f;
    s->parallax_mapping_steps = 32;
    s->ssao_sample_count = 64;                    // rendererSettings.h:69-73
    s->ssao_radius = 0.5f;
    s->ssao_amount = 1.0f;
    s->enable_clipping = 1;                       // rendererSettings.h:40
}

// Returns RT_OK or an error code with a message in `why`.
inline int check_settings(const RtSettings* s, std::string& why)
{
    char buf[256];
    auto bad = [&](int code, const char* msg) { why = msg; return code; };
    if (!s) return bad(RT_ERR_INVALID, "settings is NULL");
    if (s->image_width <= 0 || s->image_height <= 0) {
        snprintf(buf, sizeof(buf), "image size %dx%d", s->image_width, s->image_height);
        return bad(RT_ERR_INVALID, buf);
    }
    if (s->enable_ssaa && s->ssaa_factor < 1) return bad(RT_ERR_INVALID, "ssaa_factor < 1");
    if (s->enable_ssao && s->ssao_sample_count < 0) return bad(RT_ERR_INVALID, "ssao_sample_count < 0");
    if (s->enable_displacement_mapping && s->parallax_mapping_steps < 1) return bad(RT_ERR_INVALID, "parallax_mapping_steps < 1");
    if (!s->enable_bvh) return bad(RT_ERR_UNSUPPORTED, "enable_bvh = false (brute force) is not offered");
    if (s->shading_method < RT_SHADING || s->shading_method > RT_VISUALIZE_AO) return bad(RT_ERR_INVALID, "shading_method out of range");
    if (s->max_recursion_depth > kMaxRecursionDepth) return bad(RT_ERR_INVALID, "max_recursion_depth above the supported maximum (8)");
    if (s->rough_reflections_sample_count < 0) return bad(RT_ERR_INVALID, "rough_reflections_sample_count < 0");
    return RT_OK;
}

struct SceneFacts {
    bool bvh_valid, camera_set;
    uint32_t n_tris;
    int n_mats, min_mat_index, max_mat_index;
    int tex_format[RT_TEX_COUNT];
    int n_shapes = 0, shape_min_mat = 0, shape_max_mat = -1;     // analytic shapes (rt_add_sphere / rt_add_plane)
};

// One analytic shape as the device reads it: 2 x float4 (rt_device.h, shape_intersect).
struct HostShape {
    float a[3];          // centre / point
    float radius2;       // Sphere::_radius2 = radius * radius (analyticShape.cpp:7), in float
    float n[3];          // plane normal
    uint32_t bits;       // material index | kind << 31 (0 = sphere, 1 = plane)
};
static_assert(sizeof(HostShape) == 32, "two float4 per shape");

inline HostShape make_sphere(const float c[3], float radius, int32_t mat)
{
    HostShape s;
    s.a[0] = c[0]; s.a[1] = c[1]; s.a[2] = c[2];
    s.radius2 = radius * radius;
    s.n[0] = s.n[1] = s.n[2] = 0.0f;
    s.bits = (uint32_t)mat & 0x7fffffffu;
    return s;
}

inline HostShape make_plane(const float p[3], const float n[3], int32_t mat)
{
    HostShape s;
    s.a[0] = p[0]; s.a[1] = p[1]; s.a[2] = p[2];
    s.radius2 = 0.0f;
    s.n[0] = n[0]; s.n[1] = n[1]; s.n[2] = n[2];
    s.bits = ((uint32_t)mat & 0x7fffffffu) | 0x80000000u;
    return s;
}

inline int check_scene_for_render(const SceneFacts& f, const RtSettings* s, std::string& why)
{
    auto bad = [&](int code, const char* msg) { why = msg; return code; };
    if (!f.bvh_valid) return bad(RT_ERR_STATE, "rt_build_bvh has not been called for the current triangles");
    if (!f.camera_set) return bad(RT_ERR_STATE, "rt_set_camera has not been called");
    const bool rt = s->shading_method == RT_SHADING;
    if (rt && f.n_tris > 0) {
        // the reference asserts on a bad material index (materials.h:117); here it is an error return
        if (f.n_mats == 0) return bad(RT_ERR_STATE, "RT_SHADING needs materials (rt_set_materials)");
        if (f.min_mat_index < 0 || f.max_mat_index >= f.n_mats) return bad(RT_ERR_STATE, "triangle material index out of range");
    }
    if (f.n_shapes > 0) {
        // a shape hit leaves HitInfo::triangle / u / v stale or null in the reference (analyticShape.cpp:9-76): everything
        // that would read them is undefined there and refused here
        if (s->enable_ao_mapping || s->enable_diffuse_mapping || s->enable_normal_mapping || s->enable_roughness_mapping ||
            s->enable_displacement_mapping)
            return bad(RT_ERR_UNSUPPORTED, "texture mapping with analytic shapes reads a stale HitInfo::triangle in the reference");
        if (s->shading_method == RT_BARYCENTRIC_COORDINATES_SHADING || s->shading_method == RT_VISUALIZE_AO)
            return bad(RT_ERR_UNSUPPORTED, "this debug shading mode reads stale HitInfo fields for analytic shapes in the reference");
        if (rt && (f.n_mats == 0 || f.shape_min_mat < 0 || f.shape_max_mat >= f.n_mats))
            return bad(RT_ERR_STATE, "analytic shape material index out of range");
    }
    struct { int on; int slot; const char* msg; } need[] = {
        {(rt || s->shading_method == RT_VISUALIZE_AO) && s->enable_ao_mapping, RT_TEX_AO, "ao mapping enabled but no ao map set"},
        {rt && s->enable_diffuse_mapping, RT_TEX_DIFFUSE, "diffuse mapping enabled but no diffuse map set"},
        {rt && s->enable_normal_mapping, RT_TEX_NORMAL, "normal mapping enabled but no normal map set"},
        {rt && s->enable_roughness_mapping, RT_TEX_ROUGHNESS, "roughness mapping enabled but no roughness map set"},
        {rt && s->enable_displacement_mapping, RT_TEX_DISPLACEMENT, "displacement mapping enabled but no displacement map set"},
        {s->enable_skysphere, RT_TEX_SKYSPHERE, "skysphere enabled but no skysphere set"},
        {s->enable_skybox && !s->enable_skysphere, RT_TEX_SKYBOX_RIGHT, "skybox enabled but its right face is not set"},
        {s->enable_skybox && !s->enable_skysphere, RT_TEX_SKYBOX_LEFT, "skybox enabled but its left face is not set"},
        {s->enable_skybox && !s->enable_skysphere, RT_TEX_SKYBOX_TOP, "skybox enabled but its top face is not set"},
        {s->enable_skybox && !s->enable_skysphere, RT_TEX_SKYBOX_BOTTOM, "skybox enabled but its bottom face is not set"},
        {s->enable_skybox && !s->enable_skysphere, RT_TEX_SKYBOX_BACK, "skybox enabled but its back face is not set"},
        {s->enable_skybox && !s->enable_skysphere, RT_TEX_SKYBOX_FRONT, "skybox enabled but its front face is not set"},
    };
    for (auto& nd : need)
        if (nd.on && f.tex_format[nd.slot] == 0) return bad(RT_ERR_STATE, nd.msg);
    return RT_OK;
}

// Final-resolution tiles owned by shard (tile_mod, tile_rem): dealt round-robin along a row-shifted order so
// that one shard's tiles never line up in a column (sky / geometry cost is spread over the shards).
inline std::vector<uint32_t> owned_tiles(const RtSettings* s, int tile_size, int tile_mod, int tile_rem, int* tiles_x_out)
{
    int tiles_x = (s->image_width + tile_size - 1) / tile_size;
    int tiles_y = (s->image_height + tile_size - 1) / tile_size;
    if (tiles_x_out) *tiles_x_out = tiles_x;
    std::vector<uint32_t> out;
    const int pitch = (tiles_x % tile_mod == 0) ? tiles_x + 1 : tiles_x;
    for (int i = 0; i < tiles_x * tiles_y; i++) {
        int ty = i / tiles_x, tx = i % tiles_x;
        if ((tx + ty * pitch) % tile_mod == tile_rem) out.push_back((uint32_t)i);
    }
    return out;
}

} // namespace rtb
