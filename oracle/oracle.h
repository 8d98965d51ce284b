// oracle.h -- TEST INFRASTRUCTURE ONLY.  C surface of liboracle.so (see oracle.cpp for the restatement).
// Same entry points as the compiled reference shell (ref_harness.cpp) with the prefix orc_ instead of ref_,
// plus counting entry points that the unmodified reference cannot offer.
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/rtb200.h"

#ifdef __cplusplus
extern "C" {
#endif

void* orc_bvh_create(const float* xyz9, size_t n, int max_depth, int leaf_max, double* build_ms);
void orc_bvh_destroy(void* h);
void orc_bvh_stats(void* h, uint64_t* out6);
double orc_bvh_intersect(void* h, const float* o3, const float* d3, size_t n, int32_t* tri_id, float* t, float* u, float* v, int threads);
// out2 = {volume tests, triangle tests} summed over the rays: the V and T of SURVEY.md section 8(d).
double orc_bvh_count(void* h, const float* o3, const float* d3, size_t n, uint64_t* out2, int threads);
int orc_triangle_intersect(const float* xyz9, const float* o3, const float* d3, float* t, float* u, float* v);

void* orc_renderer_create(void);
void orc_renderer_destroy(void* h);
void orc_renderer_configure(void* h, const RtSettings* s, float fov);
void orc_renderer_set_triangles(void* h, const float* xyz9, const float* uv6, const int32_t* mat, size_t n);
void orc_renderer_set_materials(void* h, const RtMaterial* mats, size_t n);
void orc_renderer_set_texture_f32(void* h, int slot, const float* rgba, int w, int hgt);
void orc_renderer_set_texture_u8(void* h, int slot, const uint8_t* rgba, int w, int hgt);
void orc_renderer_set_camera_transform(void* h, const float m[16]);
void orc_renderer_set_light(void* h, const float p[3]);
void orc_renderer_add_sphere(void* h, const float c[3], float radius, int mat);
void orc_renderer_add_plane(void* h, const float p[3], const float n[3], int mat);
void orc_camera_matrices(float fov, float aspect, float znear, float zfar, float* proj16, float* proj_inv16);
void orc_transform_inverse(const float m[16], float* out16);
double orc_renderer_render(void* h, uint32_t* argb_out, int threads);
double orc_renderer_trace_rows(void* h, const float cam_to_world[16], uint32_t* argb_super, int row_begin, int row_end,
                               int row_step, int reseed, int threads);
// out16[0..10]: primary_rays, shadow_rays, reflection_rays, reflection_shadow_rays, primary_hits,
// primary V, primary T, shadow V, shadow T, reflection(+its shadows) V, reflection(+its shadows) T.
void orc_renderer_count_rows(void* h, const float cam_to_world[16], int row_begin, int row_end, int row_step,
                             uint64_t* out16, int threads);
// hybrid_rasterization_tracing: Renderer::raster_trace + post_process in sequential triangle order; rng_mode / ref_seeds9 as in
// orc_renderer_render_ssao; out5 (may be NULL) = shaded fragments, of which hit, shadow / reflection / reflection-shadow rays.
void orc_renderer_raster(void* h, uint32_t* argb_out, int rng_mode, const uint32_t* ref_seeds9, uint64_t* out5);
void orc_downscale(const uint32_t* in, int w, int hgt, int factor, uint32_t* out);
int orc_omp_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
