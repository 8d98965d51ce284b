// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref/libref*.so).
//
// A C-ABI shell around the UNMODIFIED reference sources under /root/reference/tp2 (compiled where they
// lie by oracle/Makefile; nothing is copied into this repo).  It drives the reference exactly through
// its public interface: Renderer::{set_triangles,get_materials,set_*_map,set_skysphere,set_light_position,
// change_camera_fov,set_camera_transform,change_render_size,render_settings,ray_trace,post_process,
// get_image,trace_ray} (renderer.h:38-169) and BVH::BVH / BVH::intersect (bvh.h:302,307).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
#include <cstdint>
#include <cstring>
#include <chrono>
#include <vector>
#include <omp.h>

#include "renderer.h"
#include "bvh.h"
#include "triangle.h"
#include "mesh_io.h"
#include "meshIOUtils.h"
#include "mainUtils.h"
#include "imageUtils.h"
#include "mat.h"

#include "../include/rtb200.h" // RtSettings / RtMaterial PODs shared with the product's ABI

namespace {

std::vector<Triangle> make_triangles(const float* xyz9, const float* uv6, const int32_t* mat, size_t n)
{
    std::vector<Triangle> tris;
    tris.reserve(n);
    for (size_t i = 0; i < n; i++) {
        const float* p = xyz9 + 9 * i;
        Point tu(-1, -1, -1), tv(-1, -1, -1);
        if (uv6) {
            tu = Point(uv6[6 * i + 0], uv6[6 * i + 1], uv6[6 * i + 2]);
            tv = Point(uv6[6 * i + 3], uv6[6 * i + 4], uv6[6 * i + 5]);
        }
        tris.emplace_back(Point(p[0], p[1], p[2]), Point(p[3], p[4], p[5]), Point(p[6], p[7], p[8]),
                          mat ? mat[i] : -1, tu, tv);
    }
    return tris;
}

Transform transform_from(const float m[16])
{
    return Transform(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
}

void apply_settings(RenderSettings& rs, const RtSettings& s)
{
    rs.image_width = s.image_width;
    rs.image_height = s.image_height;
    rs.enable_ssaa = s.enable_ssaa != 0;
    rs.ssaa_factor = s.ssaa_factor;
    rs.hybrid_rasterization_tracing = s.hybrid_rasterization_tracing != 0;
    rs.shading_method = (RenderSettings::ShadingMethod)s.shading_method;
    rs.compute_shadows = s.compute_shadows != 0;
    rs.max_recursion_depth = s.max_recursion_depth;
    rs.enable_bvh = s.enable_bvh != 0;
    rs.bvh_max_depth = s.bvh_max_depth;
    rs.bvh_leaf_object_count = s.bvh_leaf_object_count;
    rs.enable_ssao = s.enable_ssao != 0;
    rs.enable_ambient = s.enable_ambient != 0;
    rs.enable_diffuse = s.enable_diffuse != 0;
    rs.enable_specular = s.enable_specular != 0;
    rs.enable_emissive = s.enable_emissive != 0;
    rs.rough_reflections_sample_count = s.rough_reflections_sample_count;
    rs.enable_ao_mapping = s.enable_ao_mapping != 0;
    rs.enable_diffuse_mapping = s.enable_diffuse_mapping != 0;
    rs.enable_normal_mapping = s.enable_normal_mapping != 0;
    rs.enable_displacement_mapping = s.enable_displacement_mapping != 0;
    rs.enable_roughness_mapping = s.enable_roughness_mapping != 0;
    rs.enable_skysphere = s.enable_skysphere != 0;
    rs.enable_skybox = s.enable_skybox != 0;
    rs.displacement_mapping_strength = s.displacement_mapping_strength;
    rs.parallax_mapping_steps = s.parallax_mapping_steps;
    rs.ssao_sample_count = s.ssao_sample_count;
    rs.ssao_radius = s.ssao_radius;
    rs.ssao_amount = s.ssao_amount;
    rs.enable_clipping = s.enable_clipping != 0;
}

struct RefBvh {
    std::vector<Triangle> tris;
    BVH* bvh = nullptr;
    ~RefBvh() { delete bvh; }
};

struct RefRenderer {
    Renderer renderer;
    RtSettings settings;
    float fov = 45.0f;
    Image skybox_faces[6];      // right, left, top, bottom, back, front (skybox.h:13-17); handed over as one Skybox
};

long long g_last_hits = 0;

uint32_t pixel_seed(uint32_t pixel_index, uint32_t rng_seed)
{
    // must equal rt_pixel_seed() of the product (include/rtb200.h)
    uint32_t x = pixel_index ^ rng_seed;
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x | 1u;
}

void tree_stats(const BVH::OctreeNode* n, int depth, uint64_t* out)
{
    out[0]++;                                     // nodes
    if (n->_is_leaf) {
        out[1]++;                                 // leaves
        if (n->_triangles.empty()) out[2]++;      // empty leaves
        if ((uint64_t)depth > out[4]) out[4] = depth;
        if (n->_triangles.size() > out[5]) out[5] = n->_triangles.size();
        return;
    }
    out[3]++;                                     // interior
    for (int i = 0; i < 8; i++) tree_stats(n->_children[i], depth + 1, out);
}

} // namespace

extern "C" {

// ---- OBJ loading (host I/O of the reference; used only to cut golden fixtures) ---------------------------
// Returns triangle count (or -1).  Call once with NULL outputs to size, again to fill.
int ref_load_obj(const char* path, const float* transform16, float* xyz9, float* uv6, int32_t* mat,
                 RtMaterial* mats, int max_mats, int* n_mats)
{
    MeshIOData data = read_meshio_data(path);
    if (data.positions.empty()) return -1;
    std::vector<Triangle> tris = MeshIOUtils::create_triangles(data, 0, transform16 ? transform_from(transform16) : Identity());
    if (n_mats) *n_mats = data.materials.count();
    if (xyz9) {
        for (size_t i = 0; i < tris.size(); i++) {
            const Triangle& t = tris[i];
            float* p = xyz9 + 9 * i;
            p[0] = t._a.x; p[1] = t._a.y; p[2] = t._a.z;
            p[3] = t._b.x; p[4] = t._b.y; p[5] = t._b.z;
            p[6] = t._c.x; p[7] = t._c.y; p[8] = t._c.z;
            if (uv6) {
                uv6[6 * i + 0] = t._tex_coords_u.x; uv6[6 * i + 1] = t._tex_coords_u.y; uv6[6 * i + 2] = t._tex_coords_u.z;
                uv6[6 * i + 3] = t._tex_coords_v.x; uv6[6 * i + 4] = t._tex_coords_v.y; uv6[6 * i + 5] = t._tex_coords_v.z;
            }
            if (mat) mat[i] = t._materialIndex;
        }
    }
    if (mats) {
        for (int i = 0; i < data.materials.count() && i < max_mats; i++) {
            const Material& m = data.materials.materials[i];
            RtMaterial& o = mats[i];
            o.ambient_coeff[0] = m.ambient_coeff.r; o.ambient_coeff[1] = m.ambient_coeff.g; o.ambient_coeff[2] = m.ambient_coeff.b;
            o.diffuse[0] = m.diffuse.r; o.diffuse[1] = m.diffuse.g; o.diffuse[2] = m.diffuse.b;
            o.specular[0] = m.specular.r; o.specular[1] = m.specular.g; o.specular[2] = m.specular.b;
            o.emission[0] = m.emission.r; o.emission[1] = m.emission.g; o.emission[2] = m.emission.b;
            o.reflection = m.reflection; o.roughness = m.roughness; o.ns = m.ns;
            o.specular_threshold = 0.0f; // uninitialised in the reference until precompute_materials
        }
    }
    return (int)tris.size();
}

// ---- BVH-only oracle: BVH::BVH + BVH::intersect -----------------------------------------------------------
void* ref_bvh_create(const float* xyz9, size_t n, int max_depth, int leaf_max, double* build_ms)
{
    RefBvh* h = new RefBvh();
    h->tris = make_triangles(xyz9, nullptr, nullptr, n);
    auto t0 = std::chrono::steady_clock::now();
    h->bvh = new BVH(&h->tris, max_depth, leaf_max);
    auto t1 = std::chrono::steady_clock::now();
    if (build_ms) *build_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    return h;
}

void ref_bvh_destroy(void* handle) { delete (RefBvh*)handle; }

// out[0..5] = nodes, leaves, empty leaves, interior, max depth reached, max leaf size
void ref_bvh_stats(void* handle, uint64_t* out)
{
    RefBvh* h = (RefBvh*)handle;
    for (int i = 0; i < 6; i++) out[i] = 0;
    tree_stats(h->bvh->_root, 0, out);
}

// Returns elapsed milliseconds of the query loop.
double ref_bvh_intersect(void* handle, const float* o3, const float* d3, size_t n,
                         int32_t* tri_id, float* t, float* u, float* v, int threads)
{
    RefBvh* h = (RefBvh*)handle;
    const Triangle* base = h->tris.data();
    if (threads <= 0) threads = omp_get_max_threads();
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
    for (long long i = 0; i < (long long)n; i++) {
        Ray ray(Point(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), Vector(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]));
        HitInfo hit;
        bool found = h->bvh->intersect(ray, hit);
        if (tri_id) tri_id[i] = found ? (int32_t)(hit.triangle - base) : -1;
        if (t) t[i] = found ? hit.t : -1.0f;
        if (u) u[i] = found ? hit.u : 0.0f;
        if (v) v[i] = found ? hit.v : 0.0f;
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// Triangle::intersect alone (triangle.cpp:25-91) for the KATs of tests.cpp:97-112.
int ref_triangle_intersect(const float* xyz9, const float* o3, const float* d3, float* t, float* u, float* v)
{
    Triangle tri(Point(xyz9[0], xyz9[1], xyz9[2]), Point(xyz9[3], xyz9[4], xyz9[5]), Point(xyz9[6], xyz9[7], xyz9[8]));
    Ray ray(Point(o3[0], o3[1], o3[2]), Vector(d3[0], d3[1], d3[2]));
    float tt = -1, uu = 0, vv = 0;
    bool r = tri.intersect(ray, tt, uu, vv);
    if (t) *t = tt;
    if (u) *u = uu;
    if (v) *v = vv;
    return r ? 1 : 0;
}

// ---- Renderer oracle ---------------------------------------------------------------------------------------
void* ref_renderer_create() { return new RefRenderer(); }
void ref_renderer_destroy(void* handle) { delete (RefRenderer*)handle; }

// Order of calls the harness expects (it reproduces the GUI's state machine):
// settings -> triangles -> materials -> textures -> camera -> light -> render.
void ref_renderer_configure(void* handle, const RtSettings* s, float fov)
{
    RefRenderer* h = (RefRenderer*)handle;
    h->settings = *s;
    h->fov = fov;
    apply_settings(h->renderer.render_settings(), *s);
    h->renderer.change_camera_fov(fov);                               // sets _fov before the aspect ratio exists
    h->renderer.change_render_size(s->image_width, s->image_height);  // init_buffers + aspect ratio
}

void ref_renderer_set_triangles(void* handle, const float* xyz9, const float* uv6, const int32_t* mat, size_t n)
{
    RefRenderer* h = (RefRenderer*)handle;
    h->renderer.set_triangles(make_triangles(xyz9, uv6, mat, n));
}

void ref_renderer_set_materials(void* handle, const RtMaterial* mats, size_t n)
{
    RefRenderer* h = (RefRenderer*)handle;
    Materials& ms = h->renderer.get_materials();
    ms.materials.clear();
    for (size_t i = 0; i < n; i++) {
        Material m;
        m.ambient_coeff = Color(mats[i].ambient_coeff[0], mats[i].ambient_coeff[1], mats[i].ambient_coeff[2]);
        m.diffuse = Color(mats[i].diffuse[0], mats[i].diffuse[1], mats[i].diffuse[2]);
        m.specular = Color(mats[i].specular[0], mats[i].specular[1], mats[i].specular[2]);
        m.emission = Color(mats[i].emission[0], mats[i].emission[1], mats[i].emission[2]);
        m.reflection = mats[i].reflection;
        m.roughness = mats[i].roughness;
        m.ns = mats[i].ns;
        m.specular_threshold = mats[i].specular_threshold;
        ms.materials.push_back(m);
    }
}

void ref_renderer_set_texture_f32(void* handle, int slot, const float* rgba, int w, int hgt)
{
    RefRenderer* h = (RefRenderer*)handle;
    Image img(w, hgt);
    for (size_t i = 0; i < (size_t)w * hgt; i++)
        img(i) = Color(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]);
    switch (slot) {
    case RT_TEX_AO: h->renderer.set_ao_map(img); break;
    case RT_TEX_DIFFUSE: h->renderer.set_diffuse_map(img); break;
    case RT_TEX_NORMAL: h->renderer.set_normal_map(img); break;
    case RT_TEX_ROUGHNESS: h->renderer.set_roughness_map(img); break;
    case RT_TEX_SKYSPHERE: h->renderer.set_skysphere(img); break;
    case RT_TEX_DISPLACEMENT: h->renderer.set_displacement_map(img); break;
    default:
        if (slot >= RT_TEX_SKYBOX_RIGHT && slot <= RT_TEX_SKYBOX_FRONT) {
            h->skybox_faces[slot - RT_TEX_SKYBOX_RIGHT] = img;
            h->renderer.set_skybox(Skybox(h->skybox_faces));
        }
        break;
    }
}

// u8 texels go through the arithmetic of read_image (image_io.cpp:115-121): Color(u8...) / 255.
void ref_renderer_set_texture_u8(void* handle, int slot, const uint8_t* rgba, int w, int hgt)
{
    std::vector<float> f((size_t)w * hgt * 4);
    for (size_t i = 0; i < (size_t)w * hgt; i++) {
        Color pixel = Color(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]) / 255;
        f[4 * i] = pixel.r; f[4 * i + 1] = pixel.g; f[4 * i + 2] = pixel.b; f[4 * i + 3] = pixel.a;
    }
    ref_renderer_set_texture_f32(handle, slot, f.data(), w, hgt);
}

void ref_renderer_set_camera_transform(void* handle, const float m[16])
{
    ((RefRenderer*)handle)->renderer.set_camera_transform(transform_from(m));
}

// Renderer::add_analytic_shape -- renderer.cpp:146
void ref_renderer_add_sphere(void* handle, const float c[3], float radius, int mat)
{
    ((RefRenderer*)handle)->renderer.add_analytic_shape(Sphere(Point(c[0], c[1], c[2]), radius, mat));
}
void ref_renderer_add_plane(void* handle, const float p[3], const float n[3], int mat)
{
    ((RefRenderer*)handle)->renderer.add_analytic_shape(Plane(Point(p[0], p[1], p[2]), Vector(n[0], n[1], n[2]), mat));
}

void ref_renderer_set_light(void* handle, const float p[3])
{
    ((RefRenderer*)handle)->renderer.set_light_position(Point(p[0], p[1], p[2]));
}

// The matrices the reference derives for (fov, aspect): Perspective().inverse() (mat.cpp:307-319,378-447),
// so tests can pin the product's host-side camera code.  Row-major.
void ref_camera_matrices(float fov, float aspect, float znear, float zfar, float* proj16, float* proj_inv16)
{
    Transform p = Perspective(fov, aspect, znear, zfar);
    Transform pi = p.inverse();
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            proj16[4 * r + c] = p.m[r][c];
            proj_inv16[4 * r + c] = pi.m[r][c];
        }
}

void ref_transform_inverse(const float m[16], float* out16)
{
    Transform inv = transform_from(m).inverse();
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out16[4 * r + c] = inv.m[r][c];
}

// Renderer::ray_trace() + post_process(): exactly what render() times (utils/mainUtils.cpp:6-21), minus its
// stdout chatter.  Copies Renderer::get_image() (bottom-up ARGB32) to argb_out.  Returns milliseconds.
double ref_renderer_render(void* handle, uint32_t* argb_out, int threads)
{
    RefRenderer* h = (RefRenderer*)handle;
    if (threads > 0) omp_set_num_threads(threads);
    auto t0 = std::chrono::steady_clock::now();
    h->renderer.ray_trace();
    h->renderer.post_process();
    auto t1 = std::chrono::steady_clock::now();
    if (argb_out) {
        QImage* img = h->renderer.get_image();
        memcpy(argb_out, img->raw(), sizeof(uint32_t) * (size_t)img->width() * img->height());
    }
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// SSAO (enable_ssao in the settings): the GUI's sequence -- prepare_ssao_buffers / clear_z_buffer / clear_normal_buffer
// (QT/mainwindow.cpp:174-185), ray_trace(), post_process() -- with the SSAO pass on ONE thread after srand(seed): its
// generators are seeded from std::rand() and the thread number (renderer.cpp:1254-1266), so this is the reproducible run.
// rand_values (may be NULL) receives the first n_rand values of std::rand() after srand(seed): the caller works out which of
// them became generator seeds (the default-constructed generators of renderer.cpp:1252-1253 and the private copy of the
// parallel region consume some first; the evaluation order of _mm256_set_epi32's arguments decides which lane gets which).
double ref_renderer_render_ssao(void* handle, uint32_t* argb_out, unsigned seed, uint32_t* rand_values, int n_rand)
{
    RefRenderer* h = (RefRenderer*)handle;
    h->renderer.prepare_ssao_buffers();
    h->renderer.clear_z_buffer();
    h->renderer.clear_normal_buffer();
    auto t0 = std::chrono::steady_clock::now();
    h->renderer.ray_trace();
    const int threads = omp_get_max_threads();
    if (rand_values) {
        srand(seed);
        for (int i = 0; i < n_rand; i++) rand_values[i] = (uint32_t)rand();
    }
    srand(seed);
    omp_set_num_threads(1);
    h->renderer.post_process();
    omp_set_num_threads(threads);
    auto t1 = std::chrono::steady_clock::now();
    if (argb_out) {
        QImage* img = h->renderer.get_image();
        memcpy(argb_out, img->raw(), sizeof(uint32_t) * (size_t)img->width() * img->height());
    }
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// Renderer::raster_trace() + post_process() (hybrid_rasterization_tracing in the settings given to ref_renderer_configure,
// so that init_buffers allocated the z-buffer): the GUI's sequence clear_z_buffer / clear_normal_buffer / clear_image
// (QT/mainwindow.cpp:184-190), raster_trace(), post_process() (QT/mainWindowThreads.cpp:46-57).  raster_trace() walks the
// triangles under `omp parallel for` with an unsynchronised z-buffer: ONE thread is the reproducible run (sequential
// triangle order), and with enable_ssao the SSAO pass runs after srand(seed) as in ref_renderer_render_ssao.
double ref_renderer_raster(void* handle, uint32_t* argb_out, unsigned seed, uint32_t* rand_values, int n_rand)
{
    RefRenderer* h = (RefRenderer*)handle;
    const int threads = omp_get_max_threads();
    if (h->settings.enable_ssao) {
        h->renderer.prepare_ssao_buffers();
        h->renderer.clear_normal_buffer();
    }
    h->renderer.clear_z_buffer();
    h->renderer.clear_image();
    omp_set_num_threads(1);
    auto t0 = std::chrono::steady_clock::now();
    h->renderer.raster_trace();
    if (rand_values) {
        srand(seed);
        for (int i = 0; i < n_rand; i++) rand_values[i] = (uint32_t)rand();
    }
    srand(seed);
    h->renderer.post_process();
    auto t1 = std::chrono::steady_clock::now();
    omp_set_num_threads(threads);
    if (argb_out) {
        QImage* img = h->renderer.get_image();
        memcpy(argb_out, img->raw(), sizeof(uint32_t) * (size_t)img->width() * img->height());
    }
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// Seeded variant for rough reflections and for row-subset timing: the harness runs the pixel loop of
// Renderer::ray_trace (renderer.cpp:1082-1115) itself on rows [row_begin,row_end) step row_step of the
// supersampled frame, builds each primary ray with the reference's own Transform/normalize code from the
// same public camera state the renderer was given, reseeds the public per-thread generator
// (Renderer::_xorshift_generators, renderer.h:31) with pixel_seed() and calls the public Renderer::trace_ray
// (renderer.h:144).  Output: quantised ARGB32 of the supersampled frame (rows not visited untouched).
double ref_renderer_trace_rows(void* handle, const float cam_to_world[16], uint32_t* argb_super,
                               int row_begin, int row_end, int row_step, int reseed, int threads)
{
    RefRenderer* h = (RefRenderer*)handle;
    const RtSettings& s = h->settings;
    int rw = s.enable_ssaa ? s.image_width * s.ssaa_factor : s.image_width;
    int rh = s.enable_ssaa ? s.image_height * s.ssaa_factor : s.image_height;
    Camera cam;
    cam._fov = h->fov;
    cam.set_aspect_ratio((float)rw / rh);
    Transform c2w = transform_from(cam_to_world);
    Point cam_pos = c2w(Point(0, 0, 0));
    if (threads <= 0) threads = omp_get_max_threads();
    if ((int)Renderer::_xorshift_generators.size() < threads) Renderer::_xorshift_generators.resize(threads);
    long long hits = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) num_threads(threads) reduction(+ : hits)
    for (int py = row_begin; py < row_end; py += row_step) {
        float y_world = ((float)py + 0.5f) / rh * 2 - 1;
        for (int px = 0; px < rw; px++) {
            float x_world = ((float)px + 0.5f) / rw * 2 - 1;
            Point vs = cam._perspective_proj_mat_inv(Point(x_world, y_world, -1));
            Point ws = c2w(vs);
            Ray ray(cam_pos, normalize(ws - cam_pos));
            if (reseed)
                Renderer::_xorshift_generators[omp_get_thread_num()] =
                    XorShiftGenerator(pixel_seed((uint32_t)(py * rw + px), s.rng_seed));
            bool found = false;
            HitInfo hit;
            Color c = h->renderer.trace_ray(ray, hit, 0, found);
            if (found) hits++;
            if (argb_super) argb_super[(size_t)py * rw + px] = ImageUtils::gkit_color_to_Qt_ARGB32_uint(c);
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    g_last_hits = hits;
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// Primary rays of the last ref_renderer_trace_rows call whose trace_ray reported intersection_found
// (= shadow rays traced when compute_shadows is on and no material reflects).
long long ref_last_hit_count() { return g_last_hits; }

// ImageUtils::downscale_image_qt_ARGB32 (imageUtils.h:98-147) on a caller image.
void ref_downscale(const uint32_t* in, int w, int hgt, int factor, uint32_t* out)
{
    QImage src(w, hgt, QImage::Format_ARGB32), dst;
    for (int y = 0; y < hgt; y++)
        for (int x = 0; x < w; x++) src.setPixel(x, y, in[(size_t)y * w + x]);
    ImageUtils::downscale_image_qt_ARGB32(src, dst, factor);
    memcpy(out, dst.raw(), sizeof(uint32_t) * (size_t)dst.width() * dst.height());
}

int ref_omp_max_threads() { return omp_get_max_threads(); }

} // extern "C"
