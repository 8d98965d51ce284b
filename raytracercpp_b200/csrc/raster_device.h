// raster_device.h -- per-triangle and per-fragment device functions of the hybrid path: Renderer::raster_trace
// (tp2/projets/renderer/renderer.cpp:869-1006) with clip_triangle / clip_triangles_to_plane (:672-853), matrix_transform_z
// (:855-867) and trace_triangle (:619-628).  SURVEY.md section 8(f)4.
//
// What the reference does: every triangle is taken to clip space, clipped against the six planes of the view volume
// (each plane turns one triangle into 0, 1 or 2), every piece is rasterised over its bounding box with edge functions at
// the pixel centres, a fragment whose perspective-correct depth is STRICTLY smaller than the z-buffer's replaces the pixel:
// it is shaded on the spot by intersecting the pixel's ray with the piece (taken back to world space) and running
// shade_ray_inter_point on that hit -- shadow rays and reflection fans through the octree as in ray_trace().  The loop over
// the triangles is an `omp parallel for` over an unsynchronised z-buffer; its deterministic meaning (one thread) is: a
// pixel ends up with the FIRST fragment, in (triangle, piece) order, of the smallest depth.  That is what is built here:
//   pass 1 (cover)  every fragment does atomicMin on a 64-bit key (ordered depth bits << 32 | triangle * 16 + piece);
//   pass 2 (emit)   the same rasterisation once more: the fragment that owns the key leaves its pixel point, which the
//                   reference accumulates by repeated float additions from the piece's bounding box (image_x += increment),
//                   so it depends on the piece and cannot be recomputed from the pixel alone;
//   pass 3 (shade)  one thread per covered pixel re-clips its triangle, takes its piece and shades it.
// Everything that decides coverage, depth or the shaded hit is written in the reference's operation order (this
// translation unit is compiled without FMA contraction) so that frames match the compiled reference bit for bit;
// oracle/oracle.cpp restates the same lines on the CPU and is pinned to the compiled reference.
//
// Plain inline code over views, like rt_device.h: tests/hostsim compiles it with g++ for the CPU-side parity tests.
#pragma once

#include "rt_device.h"

namespace rtb {

struct V4 {
    float x, y, z, w;
};
RT_DEV V4 v4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
RT_DEV V4 operator+(V4 u, V4 v) { return v4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w); }      // vec.cpp:123-126
RT_DEV V4 operator-(V4 u, V4 v) { return v4(u.x - v.x, u.y - v.y, u.z - v.z, u.w - v.w); }      // vec.cpp:128-131
RT_DEV V4 operator*(float t, V4 u) { return v4(u.x * t, u.y * t, u.z * t, u.w * t); }           // vec.cpp:133-141
RT_DEV float comp(const V4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// Transform::operator()(const vec4&) -- mat.cpp:119-132
RT_DEV V4 xform_v4(const M4& t, V4 v)
{
    float xt = t.m[0][0] * v.x + t.m[0][1] * v.y + t.m[0][2] * v.z + t.m[0][3] * v.w;
    float yt = t.m[1][0] * v.x + t.m[1][1] * v.y + t.m[1][2] * v.z + t.m[1][3] * v.w;
    float zt = t.m[2][0] * v.x + t.m[2][1] * v.y + t.m[2][2] * v.z + t.m[2][3] * v.w;
    float wt = t.m[3][0] * v.x + t.m[3][1] * v.y + t.m[3][2] * v.z + t.m[3][3] * v.w;
    return v4(xt, yt, zt, wt);
}

struct Tri4 {               // Triangle4, triangle.h:24-40
    V4 a, b, c;
    V3 tu, tv;
};
#define RT_CLIP_MAX 12      /* std::array<Triangle4, 12>, renderer.cpp:881-882 (unchecked there; pieces beyond 12 are dropped here) */

// The camera state raster_trace reads besides FrameView's (Camera::_world_to_camera_mat, _perspective_proj_mat).
struct RasterView {
    M4 world_to_cam;
    M4 proj;
    int32_t clipping;       // RenderSettings::enable_clipping
};

// is_inside<plane_index, plane_sign> -- renderer.cpp:630-670
RT_DEV bool clip_inside(int plane, int sign, const V4& v) { return sign > 0 ? comp(v, plane) < v.w : comp(v, plane) > -v.w; }

// Renderer::clip_triangles_to_plane<plane_index, plane_sign> -- renderer.cpp:672-835 (CLIPPING_EPSILON = 0, so
// tP - CLIPPING_EPSILON is tP).  in == out is allowed for one triangle (the first stage of clip_triangle).
RT_DEV int clip_to_plane(int plane, int sign, const Tri4* in, int n, Tri4* out)
{
    int added = 0;
    const float fs = (float)sign;
    for (int i = 0; i < n; i++) {
        const Tri4 t = in[i];
        const bool ai = clip_inside(plane, sign, t.a), bi = clip_inside(plane, sign, t.b), ci = clip_inside(plane, sign, t.c);
        const int sum = (int)ai + (int)bi + (int)ci;
        if (sum == 3) {
            if (added < RT_CLIP_MAX) out[added++] = t;
        } else if (sum == 1) {
            V4 iv, o1, o2;
            float u0, u1, u2, w0, w1, w2;                                        // { inside, outside_1, outside_2 }
            if (ai) { iv = t.a; o1 = t.b; o2 = t.c; u0 = t.tu.x; u1 = t.tu.y; u2 = t.tu.z; w0 = t.tv.x; w1 = t.tv.y; w2 = t.tv.z; }
            else if (bi) { iv = t.b; o1 = t.c; o2 = t.a; u0 = t.tu.y; u1 = t.tu.z; u2 = t.tu.x; w0 = t.tv.y; w1 = t.tv.z; w2 = t.tv.x; }
            else { iv = t.c; o1 = t.a; o2 = t.b; u0 = t.tu.z; u1 = t.tu.x; u2 = t.tu.y; w0 = t.tv.z; w1 = t.tv.x; w2 = t.tv.y; }
            const float din = comp(iv, plane) - iv.w * fs;
            const float d1 = comp(o1, plane) - o1.w * fs;
            const float d2 = comp(o2, plane) - o2.w * fs;
            const float tp1 = d1 / (d1 - din);
            const float tp2 = d2 / (d2 - din);
            Tri4 nt;
            nt.a = iv;
            nt.b = o1 + tp1 * (iv - o1);
            nt.c = o2 + tp2 * (iv - o2);
            nt.tu = v3(u0, u1 + tp1 * (u0 - u1), u2 + tp2 * (u0 - u2));
            nt.tv = v3(w0, w1 + tp1 * (w0 - w1), w2 + tp2 * (w0 - w2));
            if (added < RT_CLIP_MAX) out[added++] = nt;
        } else if (sum == 2) {
            V4 i1, i2, ov;
            float u0, u1, u2, w0, w1, w2;                                        // { inside_1, inside_2, outside }
            if (!ai) { ov = t.a; i1 = t.b; i2 = t.c; u0 = t.tu.y; u1 = t.tu.z; u2 = t.tu.x; w0 = t.tv.y; w1 = t.tv.z; w2 = t.tv.x; }
            else if (!bi) { ov = t.b; i1 = t.c; i2 = t.a; u0 = t.tu.z; u1 = t.tu.x; u2 = t.tu.y; w0 = t.tv.z; w1 = t.tv.x; w2 = t.tv.y; }
            else { ov = t.c; i1 = t.a; i2 = t.b; u0 = t.tu.x; u1 = t.tu.y; u2 = t.tu.z; w0 = t.tv.x; w1 = t.tv.y; w2 = t.tv.z; }
            const float d1 = comp(i1, plane) - i1.w * fs;
            const float d2 = comp(i2, plane) - i2.w * fs;
            const float dout = comp(ov, plane) - ov.w * fs;
            const float tp1 = dout / (dout - d1);
            const float tp2 = dout / (dout - d2);
            const V4 p1 = ov + tp1 * (i1 - ov);
            const V4 p2 = ov + tp2 * (i2 - ov);
            Tri4 t1, t2;
            t1.a = i1; t1.b = i2; t1.c = p2;
            t1.tu = v3(u0, u1, u2 + tp2 * (u1 - u2));
            t1.tv = v3(w0, w1, w2 + tp2 * (w1 - w2));
            t2.a = i1; t2.b = p2; t2.c = p1;
            t2.tu = v3(u0, u2 + tp2 * (u1 - u2), u2 + tp1 * (u0 - u2));
            t2.tv = v3(w0, w2 + tp2 * (w1 - w2), w2 + tp1 * (w0 - w2));
            if (added < RT_CLIP_MAX) out[added++] = t1;
            if (added < RT_CLIP_MAX) out[added++] = t2;
        }
    }
    return added;
}

// World-space triangle -> clip space (renderer.cpp:888-894) -> Renderer::clip_triangle (:837-853: right, left, top, bottom,
// far, near, ping-ponging between the two arrays).  The pieces end up in `clipped`; `scratch` is the other array.
RT_DEV int raster_clip(const RasterView& rv, V3 a, V3 b, V3 c, V3 tu, V3 tv, Tri4* scratch, Tri4* clipped)
{
    const V3 ca = xform_point(rv.world_to_cam, a), cb = xform_point(rv.world_to_cam, b), cc = xform_point(rv.world_to_cam, c);
    scratch[0].a = xform_v4(rv.proj, v4(ca.x, ca.y, ca.z, 1.0f));
    scratch[0].b = xform_v4(rv.proj, v4(cb.x, cb.y, cb.z, 1.0f));
    scratch[0].c = xform_v4(rv.proj, v4(cc.x, cc.y, cc.z, 1.0f));
    scratch[0].tu = tu; scratch[0].tv = tv;
    int n = 1;
    if (rv.clipping) {
        n = clip_to_plane(0, 1, scratch, n, scratch);
        n = clip_to_plane(0, -1, scratch, n, clipped);
        n = clip_to_plane(1, 1, clipped, n, scratch);
        n = clip_to_plane(1, -1, scratch, n, clipped);
        n = clip_to_plane(2, 1, clipped, n, scratch);
        n = clip_to_plane(2, -1, scratch, n, clipped);
    } else
        clipped[0] = scratch[0];
    return n;
}

// double -> int as the reference's host does it (cvttsd2si): INT_MIN for NaN and for values outside the int range,
// where CUDA's conversion saturates.  A piece with a vertex at infinity (clipping off, w = 0) then gets max < min: no pixels.
RT_DEV int d2i_x86(double x)
{
#if defined(__CUDACC__)
    return (x > -2147483649.0 && x < 2147483648.0) ? (int)x : (int)0x80000000;
#else
    return (int)x;
#endif
}

// One clipped piece, ready to be rasterised (renderer.cpp:896-924,956-966).
struct RasterPiece {
    V3 a, b, c;                     // clipped_triangle_NDC: Triangle(const Triangle4&), triangle.cpp:12-23
    float inv_area;
    float iza, izb, izc;            // 1 / (z of the piece's vertices back in world space), matrix_transform_z(cam_to_world, proj_inv(v))
    int min_x, min_y, max_x, max_y; // bounding box in pixels, clamped to the frame
};

// Renderer::matrix_transform_z -- renderer.cpp:855-867
RT_DEV float matrix_transform_z(const M4& m, V3 p)
{
    const float zt = m.m[2][0] * p.x + m.m[2][1] * p.y + m.m[2][2] * p.z + m.m[2][3];
    const float wt = m.m[3][0] * p.x + m.m[3][1] * p.y + m.m[3][2] * p.z + m.m[3][3];
    if (wt == 1.0f) return zt;
    return zt / wt;
}

RT_DEV V3 ndc_vertex(const V4& v)
{
    const float iw = 1.0f / v.w;
    return v3(v.x * iw, v.y * iw, v.z * iw);
}

RT_DEV RasterPiece raster_piece(const FrameView& fr, const Tri4& t)
{
    RasterPiece p;
    p.a = ndc_vertex(t.a); p.b = ndc_vertex(t.b); p.c = ndc_vertex(t.c);
    p.inv_area = 1 / ((p.b.x - p.a.x) * (p.c.y - p.a.y) - (p.b.y - p.a.y) * (p.c.x - p.a.x));
    const float bminx = fminf(p.a.x, fminf(p.b.x, p.c.x)), bminy = fminf(p.a.y, fminf(p.b.y, p.c.y));
    const float bmaxx = fmaxf(p.a.x, fmaxf(p.b.x, p.c.x)), bmaxy = fmaxf(p.a.y, fmaxf(p.b.y, p.c.y));
    // std::min / std::max return their first argument when the comparison is false (NaN): a NaN vertex poisons the box in
    // the reference unless it comes first; fminf / fmaxf drop it.  A piece with a NaN vertex covers nothing either way --
    // its edge functions are NaN at every pixel and `u < 0` is false but the depth test `z < zbuf` fails on the NaN depth.
    p.min_x = d2i_x86((double)(bminx + 1) * 0.5 * (double)fr.rw);
    p.min_y = d2i_x86((double)(bminy + 1) * 0.5 * (double)fr.rh);
    p.max_x = d2i_x86((double)(bmaxx + 1) * 0.5 * (double)fr.rw);
    p.max_y = d2i_x86((double)(bmaxy + 1) * 0.5 * (double)fr.rh);
    p.min_x = p.min_x > 0 ? p.min_x : 0;
    p.min_y = p.min_y > 0 ? p.min_y : 0;
    p.max_x = p.max_x < fr.rw - 1 ? p.max_x : fr.rw - 1;
    p.max_y = p.max_y < fr.rh - 1 ? p.max_y : fr.rh - 1;
    p.iza = 1 / matrix_transform_z(fr.cam_to_world, xform_point(fr.proj_inv, p.a));
    p.izb = 1 / matrix_transform_z(fr.cam_to_world, xform_point(fr.proj_inv, p.b));
    p.izc = 1 / matrix_transform_z(fr.cam_to_world, xform_point(fr.proj_inv, p.c));
    return p;
}

// The scale of a pixel in NDC (renderer.cpp:877-878) and the start of a piece's pixel-point sequences (:926,932).
RT_DEV float raster_scale(int extent) { return 1.0f / (float)extent * 2; }
RT_DEV float raster_start(int first_pixel, float scale) { return (float)first_pixel * scale - 1; }

// Triangle::edge_function -- triangle.h:65-68 (z is not read)
RT_DEV float edge_function(float px, float py, V3 a, V3 b) { return (b.x - a.x) * (py - a.y) - (b.y - a.y) * (px - a.x); }

// One pixel of a piece (renderer.cpp:940-972): image_x / image_y are the accumulated coordinates of the pixel's lower left
// corner.  Returns true when the pixel centre is inside; u, v, w are the normalised barycentrics, z the depth.
RT_DEV bool raster_fragment(const RasterPiece& p, float image_x, float image_y, float sx, float sy, float& ppx, float& ppy, float& u, float& v,
                            float& w, float& z)
{
    ppx = image_x + sx * 0.5f;
    ppy = image_y + sy * 0.5f;
    u = edge_function(ppx, ppy, p.c, p.a);
    if (u < 0) return false;
    v = edge_function(ppx, ppy, p.a, p.b);
    if (v < 0) return false;
    w = edge_function(ppx, ppy, p.b, p.c);
    if (w < 0) return false;
    u *= p.inv_area; v *= p.inv_area; w *= p.inv_area;
    z = -1 / (p.iza * w + p.izb * u + p.izc * v);
    return true;
}

// The z-buffer as 64-bit keys: (order-preserving image of the depth) << 32 | triangle * 16 + piece.  `z < zbuf` against a
// buffer cleared to +inf (renderer.cpp:974, :165-168) never holds for NaN or +inf: such fragments have no key.
#define RT_RASTER_EMPTY 0xffffffffffffffffull
RT_DEV bool raster_key(float z, uint32_t order, unsigned long long& key)
{
    if (!(z < INFINITY)) return false;
    if (z == 0.0f) z = 0.0f;                                                      // -0 and +0 compare equal in the reference
    const uint32_t b = f4_bits(z);
    const uint32_t ordered = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    key = ((unsigned long long)ordered << 32) | (unsigned long long)order;
    return true;
}
RT_DEV float raster_key_depth(unsigned long long key)
{
    const uint32_t o = (uint32_t)(key >> 32);
    const uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#if defined(__CUDACC__)
    return __uint_as_float(b);
#else
    float f; __builtin_memcpy(&f, &b, 4); return f;
#endif
}

// The world-space triangle of leaf-order index `tri` as raster_trace sees it: vertices, Triangle::_normal, texture
// coordinates, material, and its index in the caller's array (the order of the reference's loop).
struct RasterSource {
    V3 a, b, c, normal, tu, tv;
    int32_t mat, orig;
};
RT_DEV RasterSource raster_source(const SceneView& sc, uint32_t tri)
{
    const rt_f4* tp = sc.tris + 3 * (size_t)tri;
    const rt_f4 p0 = RT_LDG4(tp), p1 = RT_LDG4(tp + 1), p2 = RT_LDG4(tp + 2);
    const TriShade ts = load_tri_shade(sc, (int32_t)tri);
    RasterSource s;
    s.a = v3(p0.x, p0.y, p0.z); s.b = v3(p1.x, p1.y, p1.z); s.c = v3(p2.x, p2.y, p2.z);
    s.normal = v3(p0.w, p1.w, p2.w);
    s.tu = ts.tu; s.tv = ts.tv; s.mat = ts.mat; s.orig = ts.orig;
    return s;
}

// What the shade pass needs of the winning fragment beyond its key.
struct RasterShadeOut {
    Col colour;
    bool shaded;            // RT_SHADING: the fragment's ray was tested against its piece (counts as a traced primary ray)
    bool hit;               // ... and hit it
    bool shadow_ray;
    V3 normal;              // Renderer::_normal_buffer: the ORIGINAL triangle's un-normalised normal (renderer.cpp:978-979)
};

// Shading of the fragment that won pixel `pix`: renderer.cpp:981-997.  (ppx, ppy) is the pixel point the emit pass left,
// `piece_index` the piece of leaf-order triangle `tri`.  frag_slot: where this thread may park the piece's texture
// coordinates so that the shading functions find them as triangle sc.n_tris + frag_index (load_tri_shade).
template <bool COUNT>
RT_DEV RasterShadeOut raster_shade(const SceneView& sc, const FrameView& fr, const RasterView& rv, uint32_t tri, int piece_index, uint32_t pix,
                                   float ppx, float ppy, rt_f4* frag_slot, uint32_t frag_index, TraceCounters* tc)
{
    RasterShadeOut out;
    out.colour = col(0.0f); out.shaded = false; out.hit = false; out.shadow_ray = false;
    const RasterSource src = raster_source(sc, tri);
    out.normal = src.normal;
    const int mode = fr.s.shading_method;
    if (mode == RT_ABS_NORMALS_SHADING) {                                         // shade_abs_normals(normalize(_normal)), :404-407
        const V3 n = normalize(src.normal);
        out.colour = col(fabsf(n.x), fabsf(n.y), fabsf(n.z));
        return out;
    }
    if (mode == RT_PASTEL_NORMALS_SHADING) {                                      // :409-412
        const V3 n = normalize(src.normal);
        out.colour = (col(n.x, n.y, n.z) + col(1.0f)) * 0.5f;
        return out;
    }
    Tri4 scratch[RT_CLIP_MAX], clipped[RT_CLIP_MAX];
    const int n_pieces = raster_clip(rv, src.a, src.b, src.c, src.tu, src.tv, scratch, clipped);
    if (piece_index >= n_pieces) return out;                                      // cannot happen: the key came from the same code
    const Tri4& t4 = clipped[piece_index];
    const V3 na = ndc_vertex(t4.a), nb = ndc_vertex(t4.b), nc = ndc_vertex(t4.c);
    if (mode == RT_BARYCENTRIC_COORDINATES_SHADING || mode == RT_VISUALIZE_AO) {
        const float inv_area = 1 / ((nb.x - na.x) * (nc.y - na.y) - (nb.y - na.y) * (nc.x - na.x));
        const float u = edge_function(ppx, ppy, nc, na) * inv_area;
        const float v = edge_function(ppx, ppy, na, nb) * inv_area;
        if (mode == RT_BARYCENTRIC_COORDINATES_SHADING)
            out.colour = col(1, 0, 0) * u + col(0, 1, 0) * v + col(0, 0, 1) * (1 - u - v);      // :414-417, not clamped here
        else {
            out.colour = col(0.9f);                                              // shade_visualize_ao(proj_inv(piece), u, v), :419-434
            if (fr.s.enable_ao_mapping) {
                TriShade ts;
                ts.tu = t4.tu; ts.tv = t4.tv; ts.mat = src.mat; ts.orig = src.orig;
                float tu, tv;
                tri_texcoords(ts, u, v, tu, tv);
                out.colour = out.colour * col(tex_floor(sc.tex[RT_TEX_AO], tu, tv).r);
            }
        }
        return out;
    }
    // RT_SHADING: trace_triangle(Ray(camera, pixel), cam_to_world(proj_inv(piece)), 0) -- renderer.cpp:983-987,619-628
    const V3 o = fr.cam_pos;
    const V3 d = normalize(xform_point(fr.cam_to_world, xform_point(fr.proj_inv, v3(ppx, ppy, -1))) - fr.cam_pos);
    const V3 wa = xform_point(fr.cam_to_world, xform_point(fr.proj_inv, na));
    const V3 wb = xform_point(fr.cam_to_world, xform_point(fr.proj_inv, nb));
    const V3 wc = xform_point(fr.cam_to_world, xform_point(fr.proj_inv, nc));
    const V3 wn = cross(wb - wa, wc - wa);                                        // Triangle(a, b, c, ...), triangle.cpp:9-10
    out.shaded = true;
    rt_f4 p0, p1, p2;
    p0.x = wa.x; p0.y = wa.y; p0.z = wa.z; p0.w = wn.x;
    p1.x = wb.x; p1.y = wb.y; p1.z = wb.z; p1.w = wn.y;
    p2.x = wc.x; p2.y = wc.y; p2.z = wc.z; p2.w = wn.z;
    float t, u, v;
    if (!tri_test(p0, p1, p2, o, -d, t, u, v)) return out;                        // Color(0, 0, 0): a black pixel
    out.hit = true;
    // the piece as the triangle HitInfo::triangle points at: its texture coordinates are the clipped ones
    frag_slot[0].x = t4.tu.x; frag_slot[0].y = t4.tu.y; frag_slot[0].z = t4.tu.z; frag_slot[0].w = t4.tv.x;
    frag_slot[1].x = t4.tv.y; frag_slot[1].y = t4.tv.z;
#if defined(__CUDACC__)
    frag_slot[1].z = __int_as_float(src.mat); frag_slot[1].w = __int_as_float(src.orig);
#else
    { float fm, fo; __builtin_memcpy(&fm, &src.mat, 4); __builtin_memcpy(&fo, &src.orig, 4); frag_slot[1].z = fm; frag_slot[1].w = fo; }
#endif
    Hit hit;
    hit.tri = (int32_t)(sc.n_tris + frag_index);
    hit.t = t; hit.u = u; hit.v = v;
    hit.mat = src.mat;
    hit.normal = normalize(wn);
    {
        const V3 ab = wb - wa, ac = wc - wa;                                      // Triangle::get_tangent, triangle.cpp:134-153
        const float du1 = t4.tu.y - t4.tu.x, dv1 = t4.tv.y - t4.tv.x;
        const float du2 = t4.tu.z - t4.tu.x, dv2 = t4.tv.z - t4.tv.x;
        const float f = 1.0f / (du1 * dv2 - du2 * dv1);
        hit.tangent = v3(f * (dv2 * ab.x - dv1 * ac.x), f * (dv2 * ab.y - dv1 * ac.y), f * (dv2 * ab.z - dv1 * ac.z));
    }
    V3 p;
    MatView m;
    const Col direct = shade_direct(sc, fr, o, d, hit, p, m);                     // shade_ray_inter_point, renderer.cpp:556-617
    bool shadowed = false;
    if (fr.s.compute_shadows) {
        out.shadow_ray = true;
        shadowed = trace_occluded<COUNT>(sc, p, hit.normal, fr.light, tc) || (sc.n_shapes > 0 && shapes_occlude(sc, p, hit.normal, fr.light));
    }
    Col refl = col(0.0f);
    if (m.reflection > 0.0f) {
        XorShift32 rng;
        rng.state = pixel_seed(pix, fr.s.rng_seed);
        refl = compute_reflection<COUNT>(sc, fr, d, p, hit, m, 0, rng, tc);
    }
    out.colour = shade_compose(fr, m, direct, shadowed, refl);
    return out;
}

} // namespace rtb
