// hostsim.cpp -- TEST INFRASTRUCTURE ONLY (tests/hostsim/librtb200_hostsim.so).
//
// Compiles the per-ray source of the CUDA kernels (raytracercpp_b200/csrc/rt_device.h) with g++ and drives it
// through the SAME C ABI as librtb200.so, so that the kernel logic (traversal, shading, reflection fan, queues,
// tiles, resolve) can be checked against the oracle in this GPU-less container before GPU minutes are spent.
// It is NOT a product path: the package never loads it, `raytracercpp_b200.load_library()` only ever opens
// librtb200.so, and rt_create() of the real library fails without a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <omp.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include "../../raytracercpp_b200/csrc/rt_device.h"
#include "../../raytracercpp_b200/csrc/raster_device.h"
#include "../../raytracercpp_b200/csrc/ssao_device.h"
#include "../../raytracercpp_b200/csrc/host_common.h"
#include "../../raytracercpp_b200/csrc/scene_layout.h"

using namespace rtb;

struct HostTexture {
    std::vector<uint8_t> bytes;
    int w = 0, h = 0, format = 0;
};

struct RtContext {
    std::string error;
    std::vector<float> xyz9, uv6;
    std::vector<int32_t> mat;
    bool has_uv = false, has_mat = false;
    int min_mat = 0, max_mat = -1;
    bool bvh_valid = false, camera_set = false;
    FlatScene flat;
    RtBvhInfo info{};
    std::vector<F4> mats;
    int n_mats = 0;
    bool any_reflective = false;
    HostTexture tex[RT_TEX_COUNT];
    M4 proj_inv{}, cam_to_world{};
    V3 cam_pos{0, 0, 0}, light{3, 3, 2};
    M4 proj{};
    float proj_fov = 45.0f, proj_aspect = 1.0f;
    bool proj_set = false;
    int leaf_split = 8;
    std::vector<HostShape> shapes;
};

static int fail(RtContext* c, int code, const std::string& msg)
{
    if (c) c->error = msg;
    return code;
}

static SceneView scene_view(const RtContext* c)
{
    SceneView sc;
    memset(&sc, 0, sizeof(sc));
    sc.recs = c->flat.recs.data();
    sc.tris = c->flat.tris.data();
    sc.shade = c->flat.shade.data();
    sc.mats = c->mats.data();
    sc.orig = c->flat.orig.data();
    sc.n_mats = c->n_mats;
    sc.n_tris = (uint32_t)(c->xyz9.size() / 9);
    sc.shapes = reinterpret_cast<const F4*>(c->shapes.data());
    sc.n_shapes = (int32_t)c->shapes.size();
    for (int i = 0; i < RT_TEX_COUNT; i++) {
        sc.tex[i].data = c->tex[i].bytes.data();
        sc.tex[i].w = c->tex[i].w;
        sc.tex[i].h = c->tex[i].h;
        sc.tex[i].format = c->tex[i].format;
    }
    return sc;
}

static FrameView frame_view(const RtContext* c, const RtSettings* s)
{
    FrameView fr;
    memset(&fr, 0, sizeof(fr));
    fr.proj_inv = c->proj_inv;
    fr.cam_to_world = c->cam_to_world;
    fr.cam_pos = c->cam_pos;
    fr.light = c->light;
    fr.factor = s->enable_ssaa ? s->ssaa_factor : 1;
    fr.rw = s->image_width * fr.factor;
    fr.rh = s->image_height * fr.factor;
    fr.s = *s;
    return fr;
}

extern "C" {

void rt_default_settings(RtSettings* s) { if (s) default_settings(s); }
uint32_t rt_pixel_seed(uint32_t pixel_index, uint32_t rng_seed) { return pixel_seed(pixel_index, rng_seed); }

void rt_perspective_inverse(float fov, float aspect, float znear, float zfar, float proj_inv_out[16])
{
    M4 inv = invert_matrix(perspective_matrix(fov, aspect, znear, zfar));
    memcpy(proj_inv_out, inv.m, sizeof(inv.m));
}

void rt_invert_transform(const float m[16], float out[16])
{
    M4 a;
    memcpy(a.m, m, sizeof(a.m));
    M4 inv = invert_matrix(a);
    memcpy(out, inv.m, sizeof(inv.m));
}

void rt_transform_point(const float m[16], const float p[3], float out[3])
{
    M4 a;
    memcpy(a.m, m, sizeof(a.m));
    V3 q = xform_point(a, v3(p[0], p[1], p[2]));
    out[0] = q.x; out[1] = q.y; out[2] = q.z;
}

int rt_create(int, RtContext** out)
{
    RtContext* c = new RtContext();
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) c->cam_to_world.m[i][j] = c->proj_inv.m[i][j] = (i == j) ? 1.0f : 0.0f;
    *out = c;
    return RT_OK;
}
void rt_destroy(RtContext* c) { delete c; }
const char* rt_last_error(const RtContext* c) { return c ? c->error.c_str() : ""; }

int rt_set_option(RtContext* c, int option, int64_t value)
{
    if (option == RT_OPT_LEAF_SPLIT) { c->leaf_split = (int)value; c->bvh_valid = false; return RT_OK; }
    return (option == RT_OPT_COUNT_WORK || option == RT_OPT_CHUNK_PIXELS || option == RT_OPT_REFILL_PRIMARY || option == RT_OPT_REFILL_SHADE || option == RT_OPT_TRI_BATCH || option == RT_OPT_PACKETS || option == RT_OPT_PACKET_ROUNDS || option == RT_OPT_SCREEN_CULL || option == RT_OPT_LANES || option == RT_OPT_ITEM_ROUNDS || option == RT_OPT_PRIMARY_ROUNDS || option == RT_OPT_FUSED_ITEMS || option == RT_OPT_TOP_TABLE || option == RT_OPT_SHADOW_SORT || option == RT_OPT_DEVICE_BUILD || option == RT_OPT_ITEM_PASSES || option == RT_OPT_RASTER_UNITS || option == RT_OPT_GRAPH || option == RT_OPT_PACKET_CULL || option == RT_OPT_FAN_LANES) ? RT_OK : fail(c, RT_ERR_INVALID, "unknown option");
}

int rt_set_stream(RtContext*, void*) { return RT_OK; }

int rt_set_triangles(RtContext* c, const float* xyz9, const float* uv6, const int32_t* mat, size_t n)
{
    c->xyz9.assign(xyz9, xyz9 + 9 * n);
    c->has_uv = uv6 != nullptr;
    c->has_mat = mat != nullptr;
    if (uv6) c->uv6.assign(uv6, uv6 + 6 * n); else c->uv6.clear();
    if (mat) c->mat.assign(mat, mat + n); else c->mat.clear();
    c->min_mat = n ? (mat ? *std::min_element(mat, mat + n) : -1) : 0;
    c->max_mat = n ? (mat ? *std::max_element(mat, mat + n) : -1) : -1;
    c->bvh_valid = false;
    return RT_OK;
}

int rt_build_bvh(RtContext* c, int max_depth, int leaf_max)
{
    if (max_depth < 0 || max_depth > RT_MAX_TREE_DEPTH) return fail(c, RT_ERR_INVALID, "max_depth");
    if (leaf_max < 0) return fail(c, RT_ERR_INVALID, "leaf_max");
    size_t n = c->xyz9.size() / 9;
    build_flat_scene(c->xyz9.data(), c->has_uv ? c->uv6.data() : nullptr, c->has_mat ? c->mat.data() : nullptr, n, max_depth, leaf_max, c->leaf_split, c->flat);
    RtBvhInfo& bi = c->info;
    memset(&bi, 0, sizeof(bi));
    bi.triangles = n;
    bi.nodes = c->flat.nodes; bi.leaves = c->flat.leaves; bi.empty_leaves = c->flat.empty_leaves; bi.interior = c->flat.interior;
    bi.max_depth_reached = c->flat.max_depth_reached; bi.max_leaf_size = c->flat.max_leaf_size;
    bi.child_records = c->flat.n_records;
    bi.device_bytes = (c->flat.recs.size() + c->flat.tris.size() + c->flat.shade.size()) * sizeof(F4) + c->flat.orig.size() * 4;
    c->bvh_valid = true;
    return RT_OK;
}

int rt_bvh_info(const RtContext* c, RtBvhInfo* out)
{
    if (!c->bvh_valid) return RT_ERR_STATE;
    *out = c->info;
    return RT_OK;
}

int rt_transform_triangles(RtContext* c, const float m[16], int max_depth, int leaf_max)
{
    M4 t;
    memcpy(t.m, m, sizeof(t.m));
    for (size_t i = 0; i < c->xyz9.size() / 3; i++) {
        V3 p = xform_point(t, v3(c->xyz9[3 * i], c->xyz9[3 * i + 1], c->xyz9[3 * i + 2]));
        c->xyz9[3 * i] = p.x; c->xyz9[3 * i + 1] = p.y; c->xyz9[3 * i + 2] = p.z;
    }
    return rt_build_bvh(c, max_depth, leaf_max);
}

int rt_set_materials(RtContext* c, const RtMaterial* mats, size_t n)
{
    c->mats.resize(4 * n);
    if (n) memcpy(c->mats.data(), mats, n * sizeof(RtMaterial));
    c->n_mats = (int)n;
    c->any_reflective = false;
    for (size_t i = 0; i < n; i++) if (mats[i].reflection > 0.0f) c->any_reflective = true;
    return RT_OK;
}

static int set_tex(RtContext* c, int slot, const void* data, int w, int h, int format)
{
    if (slot < 0 || slot >= RT_TEX_COUNT || !data || w <= 0 || h <= 0) return fail(c, RT_ERR_INVALID, "texture");
    size_t bytes = (size_t)w * h * (format == 1 ? 4 : 16);
    c->tex[slot].bytes.resize(bytes + 16);
    // keep the texel array 16-byte aligned for the F4 view
    memcpy(c->tex[slot].bytes.data(), data, bytes);
    c->tex[slot].w = w; c->tex[slot].h = h; c->tex[slot].format = format;
    return RT_OK;
}
int rt_add_sphere(RtContext* c, const float center[3], float radius, int32_t mat)
{
    if (!c || !center) return RT_ERR_INVALID;
    if (mat < 0) return fail(c, RT_ERR_INVALID, "material index");
    c->shapes.push_back(make_sphere(center, radius, mat));
    return RT_OK;
}
int rt_add_plane(RtContext* c, const float point[3], const float normal[3], int32_t mat)
{
    if (!c || !point || !normal) return RT_ERR_INVALID;
    if (mat < 0) return fail(c, RT_ERR_INVALID, "material index");
    c->shapes.push_back(make_plane(point, normal, mat));
    return RT_OK;
}
int rt_clear_analytic_shapes(RtContext* c)
{
    if (!c) return RT_ERR_INVALID;
    c->shapes.clear();
    return RT_OK;
}

int rt_set_texture_f32(RtContext* c, int slot, const float* rgba, int w, int h) { return set_tex(c, slot, rgba, w, h, 2); }
int rt_set_texture_u8(RtContext* c, int slot, const uint8_t* rgba, int w, int h) { return set_tex(c, slot, rgba, w, h, 1); }
int rt_clear_texture(RtContext* c, int slot)
{
    if (slot < 0 || slot >= RT_TEX_COUNT) return fail(c, RT_ERR_INVALID, "texture slot");
    c->tex[slot] = HostTexture();
    return RT_OK;
}

int rt_set_camera(RtContext* c, const float proj_inv[16], const float cam_to_world[16], const float position[3])
{
    memcpy(c->proj_inv.m, proj_inv, 64);
    memcpy(c->cam_to_world.m, cam_to_world, 64);
    c->cam_pos = v3(position[0], position[1], position[2]);
    c->camera_set = true;
    return RT_OK;
}
int rt_set_light(RtContext* c, const float p[3]) { c->light = v3(p[0], p[1], p[2]); return RT_OK; }
int rt_set_projection(RtContext* c, float fov, float aspect, float znear, float zfar)
{
    c->proj = perspective_matrix(fov, aspect, znear, zfar);                                // raster_trace and the SSAO pass read Camera::_perspective_proj_mat
    c->proj_fov = fov; c->proj_aspect = aspect;
    c->proj_set = true;
    return RT_OK;
}

int rt_tile_count(const RtSettings* s, int tile_size, int tile_mod, int tile_rem)
{
    if (!s || tile_size <= 0 || tile_mod <= 0 || tile_rem < 0 || tile_rem >= tile_mod) return RT_ERR_INVALID;
    return (int)owned_tiles(s, tile_size, tile_mod, tile_rem, nullptr).size();
}

// Same stage order as the CUDA driver (k_primary -> k_reflect -> k_shade -> k_resolve), one loop per kernel.
int rt_render_device(RtContext* c, const RtSettings* s, uint32_t* out, int tile_size, int tile_mod, int tile_rem, RtRenderStats* stats)
{
    std::string why;
    if (int r = check_settings(s, why)) return fail(c, r, why);
    const bool ssao = s->enable_ssao != 0;
    if (ssao && tile_mod != 1) return fail(c, RT_ERR_UNSUPPORTED, "enable_ssao needs the whole frame on one device");
    if (ssao && !c->proj_set) return fail(c, RT_ERR_STATE, "enable_ssao: the projection has not been set (rt_set_projection)");
    SceneFacts f;
    f.bvh_valid = c->bvh_valid; f.camera_set = c->camera_set; f.n_tris = (uint32_t)(c->xyz9.size() / 9); f.n_mats = c->n_mats;
    f.min_mat_index = c->min_mat; f.max_mat_index = c->max_mat;
    for (int i = 0; i < RT_TEX_COUNT; i++) f.tex_format[i] = c->tex[i].format;
    f.n_shapes = (int)c->shapes.size();
    for (const HostShape& sh : c->shapes) {
        const int m = (int)(sh.bits & 0x7fffffffu);
        f.shape_min_mat = std::min(f.shape_min_mat, m);
        f.shape_max_mat = std::max(f.shape_max_mat, m);
    }
    if (int r = check_scene_for_render(f, s, why)) return fail(c, r, why);
    if (tile_size <= 0 || tile_mod <= 0 || tile_rem < 0 || tile_rem >= tile_mod) return fail(c, RT_ERR_INVALID, "tile args");

    const FrameView fr = frame_view(c, s);
    const SceneView sc = scene_view(c);
    const bool reflect = c->any_reflective && s->shading_method == RT_SHADING;
    int tiles_x = 0;
    std::vector<uint32_t> tiles = owned_tiles(s, tile_size, tile_mod, tile_rem, &tiles_x);
    const int tile_px = tile_size * fr.factor;
    std::vector<uint32_t> super_store;
    uint32_t* super = out;
    if (fr.factor > 1) {
        super_store.assign((size_t)fr.rw * fr.rh, 0);
        super = super_store.data();
    }
    RtRenderStats rs;
    memset(&rs, 0, sizeof(rs));
    // G-buffers of the SSAO pass (clear_z_buffer / clear_normal_buffer, renderer.cpp:165-173)
    std::vector<float> gz;
    std::vector<V3> gn;
    if (ssao) { gz.assign((size_t)fr.rw * fr.rh, INFINITY); gn.assign((size_t)fr.rw * fr.rh, v3(0, 0, 0)); }


    if (s->hybrid_rasterization_tracing) {
        // Renderer::raster_trace (raster_device.h): the three passes of the CUDA path (cover, emit, shade), sequentially
        if (tile_mod != 1) return fail(c, RT_ERR_UNSUPPORTED, "hybrid_rasterization_tracing renders the whole frame on one device");
        if (!c->proj_set) return fail(c, RT_ERR_STATE, "hybrid_rasterization_tracing: the projection has not been set (rt_set_projection)");
        RasterView rv;
        rv.world_to_cam = invert_matrix(c->cam_to_world);
        rv.proj = c->proj;
        rv.clipping = s->enable_clipping;
        const size_t npx = (size_t)fr.rw * fr.rh;
        std::vector<unsigned long long> keys(npx, RT_RASTER_EMPTY);
        std::vector<float> fx(npx), fy(npx);
        const float sx = raster_scale(fr.rw), sy = raster_scale(fr.rh);
        for (int pass = 0; pass < 2; pass++)
            for (uint32_t tri = 0; tri < sc.n_tris; tri++) {
                const RasterSource src = raster_source(sc, tri);
                Tri4 scratch[RT_CLIP_MAX], clipped[RT_CLIP_MAX];
                const int np = raster_clip(rv, src.a, src.b, src.c, src.tu, src.tv, scratch, clipped);
                for (int pi = 0; pi < np; pi++) {
                    const RasterPiece pc = raster_piece(fr, clipped[pi]);
                    float image_y = raster_start(pc.min_y, sy);
                    for (int py = pc.min_y; py <= pc.max_y; py++, image_y += sy) {
                        float image_x = raster_start(pc.min_x, sx);
                        for (int px = pc.min_x; px <= pc.max_x; px++, image_x += sx) {
                            float ppx, ppy, u, v, w, z;
                            if (!raster_fragment(pc, image_x, image_y, sx, sy, ppx, ppy, u, v, w, z)) continue;
                            unsigned long long key;
                            if (!raster_key(z, (uint32_t)src.orig * 16u + (uint32_t)pi, key)) continue;
                            const size_t pix = (size_t)py * fr.rw + px;
                            if (pass == 0) keys[pix] = std::min(keys[pix], key);
                            else if (keys[pix] == key) { fx[pix] = ppx; fy[pix] = ppy; }
                        }
                    }
                }
            }
        std::vector<int32_t> leaf_of(sc.n_tris);
        for (uint32_t i = 0; i < sc.n_tris; i++) leaf_of[(size_t)sc.orig[i]] = (int32_t)i;
        const uint32_t background = quantise_argb(col(135.0f / 255.0f, 206.0f / 255.0f, 235.0f / 255.0f));   // clear_image, renderer.cpp:175-180
        F4 frag_slot[2];
        SceneView scf = sc;
        scf.frag_shade = frag_slot;
        uint64_t sv2 = 0, st2 = 0;
        for (size_t pix = 0; pix < npx; pix++) {
            if (keys[pix] == RT_RASTER_EMPTY) { super[pix] = background; continue; }
            const uint32_t order = (uint32_t)(keys[pix] & 0xffffffffull);
            TraceCounters tc = zero_counters();
            const RasterShadeOut o = raster_shade<true>(scf, fr, rv, (uint32_t)leaf_of[order >> 4], (int)(order & 15u), (uint32_t)pix, fx[pix], fy[pix], frag_slot, 0u, &tc);
            if (tc.stack_overflow) return fail(c, RT_ERR_STATE, "traversal stack overflow");
            super[pix] = quantise_argb(o.colour);
            if (ssao) { gz[pix] = raster_key_depth(keys[pix]); gn[pix] = o.normal; }     // renderer.cpp:976-979
            rs.primary_rays += o.shaded; rs.primary_hits += o.hit; rs.shadow_rays += o.shadow_ray;
            rs.reflection_rays += tc.refl_rays; rs.reflection_shadow_rays += tc.refl_shadow_rays;
            sv2 += tc.vol_tests; st2 += tc.tri_tests;
        }
        rs.shadow_volume_tests = sv2; rs.shadow_triangle_tests = st2;
    }
    const bool raster = s->hybrid_rasterization_tracing != 0;
    const std::vector<uint32_t> no_tiles;
    // k_primary
    struct QEntry { uint32_t pix; HitRec hr; };
    std::vector<QEntry> queue;
    std::vector<uint32_t> refl_idx;
    for (uint32_t tile : (raster ? no_tiles : tiles)) {
        int tx = (int)(tile % (uint32_t)tiles_x), ty = (int)(tile / (uint32_t)tiles_x);
        for (int ly = 0; ly < tile_px; ly++)
            for (int lx = 0; lx < tile_px; lx++) {
                int px = tx * tile_px + lx, py = ty * tile_px + ly;
                if (px >= fr.rw || py >= fr.rh) continue;
                rs.primary_rays++;
                V3 o, d;
                primary_ray(fr, px, py, o, d);
                HitRec hr;
                TraceCounters tc = zero_counters();
                bool found = trace_closest<true>(sc, o, d, hr, &tc);
                rs.primary_volume_tests += tc.vol_tests; rs.primary_triangle_tests += tc.tri_tests;
                if (tc.stack_overflow) return fail(c, RT_ERR_STATE, "traversal stack overflow");
                // trace_ray's loop over the analytic shapes (renderer.cpp:1029-1037), then the min_t gate (:1039)
                float final_t = found ? hr.t : -1.0f;
                for (int i = 0; i < sc.n_shapes; i++) {
                    float t;
                    V3 sn;
                    int32_t sm;
                    if (shape_intersect(sc, i, o, d, t, sn, sm) && (t < final_t || final_t == -1.0f)) { final_t = t; hr.tri = -2 - i; hr.t = t; found = true; }
                }
                bool hit = found && hr.t > 0.1f;
                if (!hit) { super[(size_t)py * fr.rw + px] = quantise_argb(shade_miss(sc, fr, d)); continue; }
                if (reflect) {
                    const int32_t mat = hr.tri >= 0 ? load_tri_shade(sc, hr.tri).mat : shape_material(sc, -2 - hr.tri);
                    if (load_material(sc, mat).reflection > 0.0f) refl_idx.push_back((uint32_t)queue.size());
                }
                queue.push_back(QEntry{(uint32_t)py * (uint32_t)fr.rw + (uint32_t)px, hr});
            }
    }
    if (!raster) rs.primary_hits = queue.size();
    // k_reflect
    std::vector<Col> refl_rgb(queue.size(), col(0.0f));
    uint64_t refl_rays = 0, refl_shadow = 0, rv = 0, rtt = 0, sv = 0, stt = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : refl_rays, refl_shadow, rv, rtt)
    for (long long r = 0; r < (long long)refl_idx.size(); r++) {
        const QEntry& e = queue[refl_idx[r]];
        V3 o, d;
        primary_ray(fr, (int)(e.pix % (uint32_t)fr.rw), (int)(e.pix / (uint32_t)fr.rw), o, d);
        Hit hit = make_hit(sc, e.hr, o, d);
        V3 p;
        MatView m;
        shade_direct(sc, fr, o, d, hit, p, m);
        XorShift32 rng;
        rng.state = pixel_seed(e.pix, fr.s.rng_seed);
        TraceCounters tc = zero_counters();
        refl_rgb[refl_idx[r]] = compute_reflection<true>(sc, fr, d, p, hit, m, 0, rng, &tc);
        refl_rays += tc.refl_rays;
        refl_shadow += tc.refl_shadow_rays;
        rv += tc.vol_tests; rtt += tc.tri_tests;
    }
    rs.reflection_volume_tests = rv; rs.reflection_triangle_tests = rtt;
    if (!raster) { rs.reflection_rays = refl_rays; rs.reflection_shadow_rays = refl_shadow; }
    // k_shade
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : sv, stt)
    for (long long i = 0; i < (long long)queue.size(); i++) {
        const QEntry& e = queue[i];
        V3 o, d;
        primary_ray(fr, (int)(e.pix % (uint32_t)fr.rw), (int)(e.pix / (uint32_t)fr.rw), o, d);
        Hit hit = make_hit(sc, e.hr, o, d);
        Col cc;
        if (fr.s.shading_method != RT_SHADING) cc = shade_debug(sc, fr, hit);
        else {
            V3 p;
            MatView m;
            Col direct = shade_direct(sc, fr, o, d, hit, p, m);
            bool shadowed = false;
            TraceCounters tc = zero_counters();
            if (fr.s.compute_shadows) shadowed = trace_occluded<true>(sc, p, hit.normal, fr.light, &tc) || (sc.n_shapes > 0 && shapes_occlude(sc, p, hit.normal, fr.light));
            sv += tc.vol_tests; stt += tc.tri_tests;
            Col refl = m.reflection > 0.0f ? refl_rgb[i] : col(0.0f);
            cc = shade_compose(fr, m, direct, shadowed, refl);
        }
        if (ssao) { gz[e.pix] = -(o.z + d.z * e.hr.t); gn[e.pix] = hit.normal; }         // renderer.cpp:1104-1111, after the normal-map update
        super[e.pix] = quantise_argb(cc);
    }
    if (!raster) {
        rs.shadow_rays = (s->shading_method == RT_SHADING && s->compute_shadows) ? rs.primary_hits : 0;
        rs.shadow_volume_tests = sv; rs.shadow_triangle_tests = stt;
    }
    // k_ssao_occlusion, k_ssao_apply (ssao_device.h): Renderer::post_process runs SSAO before the SSAA resolve
    if (ssao) {
        SsaoView sv;
        sv.rw = fr.rw; sv.rh = fr.rh; sv.z = gz.data(); sv.n = gn.data();
        sv.proj = c->proj;
        sv.aspect = c->proj_aspect;
        sv.fov_mult_simd = (float)std::tan(c->proj_fov / 2 / 180 * M_PI);                    // renderer.cpp:1249
        sv.fov_mult_scalar = std::tan(((float)M_PI / 180) * (c->proj_fov / 2));              // radians(), mat.cpp:13-16, std::tan(float)
        sv.samples = s->ssao_sample_count; sv.radius = s->ssao_radius; sv.amount = s->ssao_amount;
        sv.rng_seed = s->rng_seed;
        const size_t npx = (size_t)fr.rw * fr.rh;
        std::vector<int> ao(npx);
#pragma omp parallel for schedule(dynamic, 256)
        for (long long i = 0; i < (long long)npx; i++) ao[(size_t)i] = ssao_count_pixel(sv, (size_t)i);
        for (size_t i = 0; i < npx; i++) ssao_apply_pixel(sv, ao.data(), super, i);
    }
    // k_resolve
    if (fr.factor > 1) {
        const int ff = fr.factor * fr.factor;
        for (uint32_t tile : tiles) {
            int tx = (int)(tile % (uint32_t)tiles_x), ty = (int)(tile / (uint32_t)tiles_x);
            for (int iy = 0; iy < tile_size; iy++)
                for (int ix = 0; ix < tile_size; ix++) {
                    int x = tx * tile_size + ix, y = ty * tile_size + iy;
                    if (x >= s->image_width || y >= s->image_height) continue;
                    int ar = 0, ag = 0, ab = 0;
                    for (int i = 0; i < fr.factor; i++)
                        for (int j = 0; j < fr.factor; j++) {
                            uint32_t cpx = super[(size_t)(y * fr.factor + i) * fr.rw + (size_t)x * fr.factor + j];
                            ar += (cpx >> 16) & 0xff; ag += (cpx >> 8) & 0xff; ab += cpx & 0xff;
                        }
                    ar /= ff; ag /= ff; ab /= ff;
                    out[(size_t)y * s->image_width + x] = 0xff000000u | ((uint32_t)(ar & 0xff) << 16) | ((uint32_t)(ag & 0xff) << 8) | (uint32_t)(ab & 0xff);
                }
        }
    }
    rs.kernel_launches = 0;
    rs.traced_primary_rays = rs.primary_rays;                 // the emulation has no screen cull: every sample gets a ray
    rs.primary_fetched_bytes = 64 * rs.primary_volume_tests + 48 * rs.primary_triangle_tests;   // single rays: one fetch per test
    rs.shadow_fetched_bytes = 64 * rs.shadow_volume_tests + 48 * rs.shadow_triangle_tests;
    rs.reflection_fetched_bytes = 64 * rs.reflection_volume_tests + 48 * rs.reflection_triangle_tests;
    if (stats) *stats = rs;
    return RT_OK;
}

// the host emulation has no stream: _begin renders, _end hands the stats back
static thread_local RtRenderStats g_pending_stats;
int rt_render_device_begin(RtContext* c, const RtSettings* s, uint32_t* out, int tile_size, int tile_mod, int tile_rem)
{
    return rt_render_device(c, s, out, tile_size, tile_mod, tile_rem, &g_pending_stats);
}
int rt_render_device_end(RtContext*, RtRenderStats* stats)
{
    if (stats) *stats = g_pending_stats;
    return RT_OK;
}

int rt_render(RtContext* c, const RtSettings* s, uint32_t* argb_out, RtRenderStats* stats)
{
    return rt_render_device(c, s, argb_out, 64, 1, 0, stats);
}

int rt_set_host_threads(int n)
{
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
    return RT_OK;
}

// Frames shared between the ranks of a box: where the CUDA library hands out CUDA IPC handles, the emulation hands out
// the name of a POSIX shared-memory object, so the world_size-2 gloo test exercises the same "peers store straight into
// rank 0's frame" protocol with real processes.
struct SharedFrame { void* p; size_t bytes; bool owner; char name[RT_FRAME_HANDLE_BYTES]; };
static std::vector<SharedFrame> g_frames;

int rt_frame_alloc(RtContext* c, size_t bytes, void** d_ptr_out, unsigned char handle_out[RT_FRAME_HANDLE_BYTES])
{
    static int serial = 0;
    SharedFrame f;
    memset(&f, 0, sizeof(f));
    f.bytes = std::max<size_t>(bytes, 4);
    f.owner = true;
    snprintf(f.name, sizeof(f.name) - 8, "/rtb200_hostsim_%d_%d", (int)getpid(), serial++);
    int fd = shm_open(f.name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)f.bytes) != 0) return fail(c, RT_ERR_CUDA, "shm_open");
    f.p = mmap(nullptr, f.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (f.p == MAP_FAILED) return fail(c, RT_ERR_CUDA, "mmap");
    memcpy(f.name + RT_FRAME_HANDLE_BYTES - 8, &f.bytes, 8);               // the size travels in the handle's last 8 bytes
    memcpy(handle_out, f.name, RT_FRAME_HANDLE_BYTES);
    g_frames.push_back(f);
    *d_ptr_out = f.p;
    return RT_OK;
}

int rt_frame_open(RtContext* c, const unsigned char handle[RT_FRAME_HANDLE_BYTES], void** d_ptr_out)
{
    SharedFrame f;
    memset(&f, 0, sizeof(f));
    memcpy(f.name, handle, RT_FRAME_HANDLE_BYTES);
    memcpy(&f.bytes, handle + RT_FRAME_HANDLE_BYTES - 8, 8);
    int fd = shm_open(f.name, O_RDWR, 0600);
    if (fd < 0) return fail(c, RT_ERR_CUDA, "shm_open (peer)");
    f.p = mmap(nullptr, f.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (f.p == MAP_FAILED) return fail(c, RT_ERR_CUDA, "mmap (peer)");
    g_frames.push_back(f);
    *d_ptr_out = f.p;
    return RT_OK;
}

static int frame_release(void* p, bool unlink_it)
{
    for (size_t i = 0; i < g_frames.size(); i++)
        if (g_frames[i].p == p) {
            munmap(p, g_frames[i].bytes);
            if (unlink_it && g_frames[i].owner) { char n[RT_FRAME_HANDLE_BYTES]; memcpy(n, g_frames[i].name, sizeof(n)); n[RT_FRAME_HANDLE_BYTES - 8] = 0; shm_unlink(n); }
            g_frames.erase(g_frames.begin() + (long)i);
            return RT_OK;
        }
    return RT_ERR_INVALID;
}
int rt_frame_close(RtContext*, void* p) { return p ? frame_release(p, false) : RT_OK; }
int rt_frame_free(RtContext*, void* p) { return p ? frame_release(p, true) : RT_OK; }

int rt_frame_to_host(RtContext*, const uint32_t* d_frame, uint32_t* host_out, size_t n_pixels)
{
    memcpy(host_out, d_frame, n_pixels * sizeof(uint32_t));
    return RT_OK;
}

static void tile_copy(const RtSettings* s, const uint32_t* frame_in, uint32_t* frame_out, uint32_t* staging, int tile_size, int mod, int rem, int unpack)
{
    int tiles_x = 0;
    std::vector<uint32_t> tiles = owned_tiles(s, tile_size, mod, rem, &tiles_x);
    size_t g = 0;
    for (uint32_t tile : tiles) {
        int tx = (int)(tile % (uint32_t)tiles_x), ty = (int)(tile / (uint32_t)tiles_x);
        for (int iy = 0; iy < tile_size; iy++)
            for (int ix = 0; ix < tile_size; ix++, g++) {
                int x = tx * tile_size + ix, y = ty * tile_size + iy;
                if (x >= s->image_width || y >= s->image_height) { if (!unpack) staging[g] = 0; continue; }
                if (unpack) frame_out[(size_t)y * s->image_width + x] = staging[g];
                else staging[g] = frame_in[(size_t)y * s->image_width + x];
            }
    }
}

int rt_pack_tiles(RtContext*, const RtSettings* s, const uint32_t* frame, uint32_t* staging, int tile_size, int mod, int rem)
{
    tile_copy(s, frame, nullptr, staging, tile_size, mod, rem, 0);
    return RT_OK;
}
int rt_unpack_tiles(RtContext*, const RtSettings* s, uint32_t* frame, const uint32_t* staging, int tile_size, int mod, int rem)
{
    tile_copy(s, nullptr, frame, const_cast<uint32_t*>(staging), tile_size, mod, rem, 1);
    return RT_OK;
}

int rt_unpack_gathered(RtContext*, const RtSettings* s, uint32_t* frame, const uint32_t* gathered, int tile_size, int mod, int self_rem)
{
    size_t longest = 0;
    for (int r = 0; r < mod; r++) longest = std::max(longest, owned_tiles(s, tile_size, mod, r, nullptr).size());
    for (int r = 0; r < mod; r++)
        if (r != self_rem)
            tile_copy(s, nullptr, frame, const_cast<uint32_t*>(gathered) + (size_t)r * longest * tile_size * tile_size, tile_size, mod, r, 1);
    return RT_OK;
}

int rt_intersect(RtContext* c, const float* o3, const float* d3, size_t n, int32_t* tri_id, float* t, float* u, float* v)
{
    if (!c->bvh_valid) return fail(c, RT_ERR_STATE, "rt_build_bvh has not been called for the current triangles");
    const SceneView sc = scene_view(c);
    int overflow = 0;
#pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; i++) {
        HitRec hr;
        TraceCounters tc = zero_counters();
        bool found = trace_closest<false>(sc, v3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), v3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]), hr, &tc);
        if (tc.stack_overflow) overflow = 1;
        if (tri_id) tri_id[i] = found ? c->flat.orig[hr.tri] : -1;
        if (t) t[i] = found ? hr.t : -1.0f;
        if (u) u[i] = found ? hr.u : 0.0f;
        if (v) v[i] = found ? hr.v : 0.0f;
    }
    return overflow ? fail(c, RT_ERR_STATE, "traversal stack overflow") : RT_OK;
}

int rt_occluded(RtContext* c, const float* p3, const float* n3, size_t n, uint8_t* occluded)
{
    if (!c->bvh_valid) return fail(c, RT_ERR_STATE, "rt_build_bvh has not been called for the current triangles");
    const SceneView sc = scene_view(c);
#pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; i++)
        { TraceCounters tcs = zero_counters(); const V3 pp = v3(p3[3 * i], p3[3 * i + 1], p3[3 * i + 2]), nn = v3(n3[3 * i], n3[3 * i + 1], n3[3 * i + 2]); occluded[i] = (trace_occluded<false>(sc, pp, nn, c->light, &tcs) || (sc.n_shapes > 0 && shapes_occlude(sc, pp, nn, c->light))) ? 1 : 0; }
    return RT_OK;
}

int rt_generate_primary_rays(RtContext* c, const RtSettings* s, float* o3, float* d3)
{
    std::string why;
    if (int r = check_settings(s, why)) return fail(c, r, why);
    FrameView fr = frame_view(c, s);
    for (size_t i = 0; i < (size_t)fr.rw * fr.rh; i++) {
        V3 o, d;
        primary_ray(fr, (int)(i % (size_t)fr.rw), (int)(i / (size_t)fr.rw), o, d);
        o3[3 * i] = o.x; o3[3 * i + 1] = o.y; o3[3 * i + 2] = o.z;
        d3[3 * i] = d.x; d3[3 * i + 1] = d.y; d3[3 * i + 2] = d.z;
    }
    return RT_OK;
}

int rt_resolve_ssaa(RtContext* c, const uint32_t* in, int width, int height, int factor, uint32_t* out)
{
    if (factor < 1 || width % factor || height % factor) return fail(c, RT_ERR_INVALID, "image size not divisible by the factor");
    int w = width / factor, h = height / factor, ff = factor * factor;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int ar = 0, ag = 0, ab = 0;
            for (int i = 0; i < factor; i++)
                for (int j = 0; j < factor; j++) {
                    uint32_t cpx = in[(size_t)(y * factor + i) * width + (size_t)x * factor + j];
                    ar += (cpx >> 16) & 0xff; ag += (cpx >> 8) & 0xff; ab += cpx & 0xff;
                }
            ar /= ff; ag /= ff; ab /= ff;
            out[(size_t)y * w + x] = 0xff000000u | ((uint32_t)(ar & 0xff) << 16) | ((uint32_t)(ag & 0xff) << 8) | (uint32_t)(ab & 0xff);
        }
    return RT_OK;
}

} // extern "C"
