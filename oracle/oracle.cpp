// oracle.cpp -- TEST INFRASTRUCTURE ONLY (oracle/liboracle.so).
//
// A from-scratch CPU restatement of the ray-tracing hot path of TomClabault/RayTracerCPP, written to be
// arithmetically identical (operation for operation, no FMA contraction: built with -ffp-contract=off) to the
// reference compiled with -ffp-contract=off (oracle/_ref/libref_strict.so).  Every function names the reference
// lines it follows.  It exists so that (1) the CUDA path can be checked on machines where /root/reference is
// absent, and (2) the traversal work (volume tests V, triangle tests T) of the REFERENCE algorithm can be counted,
// which the unmodified reference cannot do.  PINNED: tests/test_oracle_vs_reference.py compares it with the
// compiled reference (bit-exact on hit ids/t/u/v and on images) and with the reference's own KATs
// (tp2/projets/tests.cpp:97-112); tests/golden/ holds vectors cut from the compiled reference.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <queue>
#include <vector>
#include <omp.h>

#ifndef M_PI
#define M_PI 3.141592653589793
#endif

namespace {

// ---------------------------------------------------------------------------------------------------------------
// vec.h / vec.cpp : Point and Vector share one layout here.
struct V3 { float x, y, z; };

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }           // vec.cpp:58-61,68-71,88-91
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }           // vec.cpp:37-40,93-96
inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }                                // vec.cpp:63-66
inline V3 scale(float k, V3 a) { return v3(k * a.x, k * a.y, k * a.z); }            // vec.cpp:42-45,98-101
inline float dot(V3 u, V3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }          // vec.cpp:164-167
inline V3 cross(V3 u, V3 v)                                                         // vec.cpp:156-162
{
    return v3((u.y * v.z) - (u.z * v.y), (u.z * v.x) - (u.x * v.z), (u.x * v.y) - (u.y * v.x));
}
inline float length2(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }            // vec.cpp:174-177
inline float length(V3 v) { return std::sqrt(length2(v)); }                         // vec.cpp:169-172
inline V3 normalize(V3 v) { float kk = 1 / length(v); return scale(kk, v); }        // vec.cpp:150-154
inline V3 div(V3 a, float k) { float kk = 1.f / k; return scale(kk, a); }           // vec.cpp:52-56
inline V3 vmin(V3 a, V3 b) { return v3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); } // vec.cpp:26-29
inline V3 vmax(V3 a, V3 b) { return v3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); } // vec.cpp:31-34

// color.h / color.cpp : alpha never reaches the output of the path, rgb only.
struct Col { float r, g, b; };
inline Col col(float v) { return Col{v, v, v}; }
inline Col cadd(Col a, Col b) { return Col{a.r + b.r, a.g + b.g, a.b + b.b}; }      // color.cpp:47-50
inline Col cmul(Col a, Col b) { return Col{a.r * b.r, a.g * b.g, a.b * b.b}; }      // color.cpp:62-65
inline Col cscale(Col c, float k) { return Col{c.r * k, c.g * k, c.b * k}; }        // color.cpp:67-75
inline Col cdiv(Col a, Col b) { return Col{a.r / b.r, a.g / b.g, a.b / b.b}; }      // color.cpp:77-80

// mat.h / mat.cpp : row-major 4x4.
struct M4 {
    float m[4][4];
};

M4 identity()
{
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.m[i][j] = i == j ? 1.f : 0.f;
    return r;
}

M4 from16(const float* p)
{
    M4 r;
    memcpy(r.m, p, sizeof(r.m));
    return r;
}

// Transform::operator()(const Point&) -- mat.cpp:83-100
V3 xform_point(const M4& t, V3 p)
{
    float x = p.x, y = p.y, z = p.z;
    float xt = t.m[0][0] * x + t.m[0][1] * y + t.m[0][2] * z + t.m[0][3];
    float yt = t.m[1][0] * x + t.m[1][1] * y + t.m[1][2] * z + t.m[1][3];
    float zt = t.m[2][0] * x + t.m[2][1] * y + t.m[2][2] * z + t.m[2][3];
    float wt = t.m[3][0] * x + t.m[3][1] * y + t.m[3][2] * z + t.m[3][3];
    float w = 1.f / wt;
    if (wt == 1.f) return v3(xt, yt, zt);
    return v3(xt * w, yt * w, zt * w);
}

// Perspective() -- mat.cpp:307-319 ; radians() -- mat.cpp:13-16
M4 perspective(float fov, float aspect, float znear, float zfar)
{
    float rad = ((float)M_PI / 180) * fov;
    float itan = 1 / tanf(rad * 0.5f);
    float id = 1 / (znear - zfar);
    M4 r;
    memset(&r, 0, sizeof(r));
    r.m[0][0] = itan / aspect;
    r.m[1][1] = itan;
    r.m[2][2] = (zfar + znear) * id;
    r.m[2][3] = 2.f * zfar * znear * id;
    r.m[3][2] = -1;
    return r;
}

// Transform::inverse() -- mat.cpp:378-447 (Gauss-Jordan with full pivoting).
M4 inverse(const M4& src)
{
    M4 a = src;
    int col_of[4], row_of[4];
    int used[4] = {0, 0, 0, 0};
    for (int step = 0; step < 4; step++) {
        int prow = -1, pcol = -1;
        float big = 0.f;
        for (int j = 0; j < 4; j++) {
            if (used[j] == 1) continue;
            for (int k = 0; k < 4; k++) {
                if (used[k] == 0 && fabsf(a.m[j][k]) >= big) {
                    big = std::abs(a.m[j][k]);
                    prow = j;
                    pcol = k;
                }
            }
        }
        ++used[pcol];
        if (prow != pcol)
            for (int k = 0; k < 4; k++) std::swap(a.m[prow][k], a.m[pcol][k]);
        row_of[step] = prow;
        col_of[step] = pcol;
        float pivinv = 1.f / a.m[pcol][pcol];
        a.m[pcol][pcol] = 1.f;
        for (int j = 0; j < 4; j++) a.m[pcol][j] *= pivinv;
        for (int j = 0; j < 4; j++) {
            if (j == pcol) continue;
            float save = a.m[j][pcol];
            a.m[j][pcol] = 0;
            for (int k = 0; k < 4; k++) a.m[j][k] -= a.m[pcol][k] * save;
        }
    }
    for (int j = 3; j >= 0; j--)
        if (row_of[j] != col_of[j])
            for (int k = 0; k < 4; k++) std::swap(a.m[k][row_of[j]], a.m[k][col_of[j]]);
    return a;
}

// ---------------------------------------------------------------------------------------------------------------
// triangle.h / triangle.cpp, hitInfo.h
struct Tri {
    V3 a, b, c;
    V3 normal;   // cross(b - a, c - a), NOT normalised -- triangle.cpp:9-10
    int mat;
    V3 tu, tv;   // per-vertex texture coordinates (u_a,u_b,u_c) / (v_a,v_b,v_c) -- triangle.h:99-103
};

struct Hit {       // hitInfo.h:8-29
    int tri = -1;  // index instead of the reference's pointer
    float t = -1;
    float u = 1.0f, v = 0.0f;
    int mat = -1;
    V3 normal = {0, 0, 0};
    V3 tangent = {0, 0, 0};
    const struct Tri* frag = nullptr;  // HitInfo::triangle when it points at raster_trace's temporary (renderer.cpp:619-628), not into _triangles
};

struct RayT { V3 o, d; };

// Triangle::get_tangent -- triangle.cpp:134-153
V3 tri_tangent(const Tri& tr, V3 ab, V3 ac)
{
    float u1 = tr.tu.x, v1 = tr.tv.x;
    float u2 = tr.tu.y, v2 = tr.tv.y;
    float u3 = tr.tu.z, v3_ = tr.tv.z;
    float du1 = u2 - u1, dv1 = v2 - v1, du2 = u3 - u1, dv2 = v3_ - v1;
    float f = 1.0f / (du1 * dv2 - du2 * dv1);
    return v3(f * (dv2 * ab.x - dv1 * ac.x), f * (dv2 * ab.y - dv1 * ac.y), f * (dv2 * ab.z - dv1 * ac.z));
}

// Triangle::intersect (MOLLER_TRUMBORE 1, BACKFACE_CULLING 1) -- triangle.cpp:25-91
bool tri_intersect(const Tri& tr, int index, const RayT& ray, Hit& hit)
{
    V3 ab = sub(tr.b, tr.a);
    V3 ac = sub(tr.c, tr.a);
    V3 oa = sub(ray.o, tr.a);
    V3 md = neg(ray.d);
    V3 mdxoa = cross(md, oa);
    float det = dot(tr.normal, md);
    if (det <= 0) return false;
    det = 1 / det;
    float u = dot(mdxoa, ac) * det;
    if (u < 0 || u > 1) return false;
    float v = dot(mdxoa, neg(ab)) * det;
    if (v < 0 || u + v > 1) return false;
    float t = dot(tr.normal, oa) * det;
    if (t < 0) return false;
    hit.t = t;
    hit.u = u;
    hit.v = v;
    hit.tangent = tri_tangent(tr, ab, ac);
    hit.mat = tr.mat;
    hit.normal = normalize(tr.normal);
    hit.tri = index;
    return true;
}

// Triangle::interpolate_texcoords -- triangle.cpp:155-160
void tri_texcoords(const Tri& tr, float u, float v, float& tex_u, float& tex_v)
{
    tex_u = (1 - u - v) * tr.tu.x + u * tr.tu.y + v * tr.tu.z;
    tex_v = (1 - u - v) * tr.tv.x + u * tr.tv.y + v * tr.tv.z;
}

// Triangle::bbox_centroid -- triangle.cpp:162-165
V3 tri_bbox_centroid(const Tri& tr)
{
    return div(add(vmin(tr.a, vmin(tr.b, tr.c)), vmax(tr.a, vmax(tr.b, tr.c))), 2);
}

std::vector<Tri> make_tris(const float* xyz9, const float* uv6, const int32_t* mat, size_t n)
{
    std::vector<Tri> out(n);
    for (size_t i = 0; i < n; i++) {
        const float* p = xyz9 + 9 * i;
        Tri& t = out[i];
        t.a = v3(p[0], p[1], p[2]);
        t.b = v3(p[3], p[4], p[5]);
        t.c = v3(p[6], p[7], p[8]);
        t.normal = cross(sub(t.b, t.a), sub(t.c, t.a));
        t.mat = mat ? mat[i] : -1;
        t.tu = uv6 ? v3(uv6[6 * i], uv6[6 * i + 1], uv6[6 * i + 2]) : v3(-1, -1, -1);
        t.tv = uv6 ? v3(uv6[6 * i + 3], uv6[6 * i + 4], uv6[6 * i + 5]) : v3(-1, -1, -1);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// bvh.h / bvh.cpp
constexpr int PLANES = 7;                                                           // bvh.h:17

struct PlaneSet {
    V3 n[PLANES];
    PlaneSet()                                                                      // bvh.cpp:8-16
    {
        float s = std::sqrt(3.0f) / 3;
        n[0] = v3(1, 0, 0);
        n[1] = v3(0, 1, 0);
        n[2] = v3(0, 0, 1);
        n[3] = v3(s, s, s);
        n[4] = v3(-s, s, s);
        n[5] = v3(-s, -s, s);
        n[6] = v3(s, -s, s);
    }
};
const PlaneSet PLANE_NORMALS;

struct Counters {
    uint64_t volume_tests = 0;
    uint64_t triangle_tests = 0;
};

struct Volume {                                                                     // bvh.h:15-106
    float d_near[PLANES], d_far[PLANES];
    Volume()
    {
        for (int i = 0; i < PLANES; i++) { d_near[i] = INFINITY; d_far[i] = -INFINITY; }
    }
    void extend(const float* nr, const float* fr)                                   // bvh.h:48-55
    {
        for (int i = 0; i < PLANES; i++) {
            d_near[i] = std::min(d_near[i], nr[i]);
            d_far[i] = std::max(d_far[i], fr[i]);
        }
    }
    void extend(const Tri& t)                                                       // bvh.h:33-46,62-76
    {
        float nr[PLANES], fr[PLANES];
        const V3* verts = &t.a;
        for (int i = 0; i < PLANES; i++) {
            nr[i] = INFINITY;
            fr[i] = -INFINITY;
            for (int j = 0; j < 3; j++) {
                float dist = dot(PLANE_NORMALS.n[i], verts[j]);
                nr[i] = std::min(nr[i], dist);
                fr[i] = std::max(fr[i], dist);
            }
        }
        extend(nr, fr);
    }
    // BoundingVolume::intersect -- bvh.h:79-105
    bool intersect(float& t_near, float& t_far, const float* denoms, const float* numers) const
    {
        t_near = -INFINITY;
        t_far = INFINITY;
        for (int i = 0; i < PLANES; i++) {
            float denom = denoms[i];
            if (denom == 0.0) continue;
            float dn = (d_near[i] - numers[i]) / denom;
            float df = (d_far[i] - numers[i]) / denom;
            if (denom < 0) std::swap(dn, df);
            t_near = std::max(t_near, dn);
            t_far = std::min(t_far, df);
            if (t_far < t_near) return false;
        }
        return true;
    }
};

struct Node {                                                                       // bvh.h:108-289
    bool leaf = true;
    std::vector<int> tris;      // indices in insertion order (the reference keeps Triangle*)
    Node* child[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    V3 lo, hi;
    Volume vol;

    Node(V3 l, V3 h) : lo(l), hi(h) {}
    ~Node()
    {
        if (!leaf)
            for (int i = 0; i < 8; i++) delete child[i];
    }
};

struct QueueElement {                                                               // bvh.h:110-124
    const Node* node;
    float t_near;
};
struct QueueGreater {
    bool operator()(const QueueElement& a, const QueueElement& b) const { return a.t_near > b.t_near; }
};

struct Octree {
    const std::vector<Tri>* tris = nullptr;
    Node* root = nullptr;
    int max_depth = 10, leaf_max = 8;

    ~Octree() { delete root; }

    // OctreeNode::create_children -- bvh.h:153-167.  The "+" in children 2, 4 and 6 is the reference's: their
    // lower corner is lo + (0, mid_y, 0) etc., not (lo.x, mid_y, lo.z); the cells that follow inherit it.
    static void create_children(Node* n)
    {
        float mx = (n->lo.x + n->hi.x) / 2;
        float my = (n->lo.y + n->hi.y) / 2;
        float mz = (n->lo.z + n->hi.z) / 2;
        V3 lo = n->lo, hi = n->hi;
        n->child[0] = new Node(lo, v3(mx, my, mz));
        n->child[1] = new Node(v3(mx, lo.y, lo.z), v3(hi.x, my, mz));
        n->child[2] = new Node(add(lo, v3(0, my, 0)), v3(mx, hi.y, mz));
        n->child[3] = new Node(v3(mx, my, lo.z), v3(hi.x, hi.y, mz));
        n->child[4] = new Node(add(lo, v3(0, 0, mz)), v3(mx, my, hi.z));
        n->child[5] = new Node(v3(mx, lo.y, mz), v3(hi.x, my, hi.z));
        n->child[6] = new Node(add(lo, v3(0, my, mz)), v3(mx, hi.y, hi.z));
        n->child[7] = new Node(v3(mx, my, mz), v3(hi.x, hi.y, hi.z));
    }

    // OctreeNode::insert_to_children -- bvh.h:195-210
    void insert_to_children(Node* n, int tri, int depth)
    {
        V3 c = tri_bbox_centroid((*tris)[tri]);
        float mx = (n->lo.x + n->hi.x) / 2;
        float my = (n->lo.y + n->hi.y) / 2;
        float mz = (n->lo.z + n->hi.z) / 2;
        int oct = 0;
        if (c.x > mx) oct += 1;
        if (c.y > my) oct += 2;
        if (c.z > mz) oct += 4;
        insert(n->child[oct], tri, depth + 1);
    }

    // OctreeNode::insert -- bvh.h:169-193
    void insert(Node* n, int tri, int depth)
    {
        bool depth_exceeded = depth == max_depth;
        if (n->leaf || depth_exceeded) {
            n->tris.push_back(tri);
            if (n->tris.size() > (size_t)leaf_max && !depth_exceeded) {
                n->leaf = false;
                create_children(n);
                for (int t : n->tris) insert_to_children(n, t, depth);
                n->tris.clear();
                n->tris.shrink_to_fit();
            }
        } else
            insert_to_children(n, tri, depth);
    }

    // OctreeNode::compute_volume -- bvh.h:141-151
    Volume compute_volume(Node* n)
    {
        if (n->leaf)
            for (int t : n->tris) n->vol.extend((*tris)[t]);
        else
            for (int i = 0; i < 8; i++) {
                Volume cv = compute_volume(n->child[i]);
                n->vol.extend(cv.d_near, cv.d_far);
            }
        return n->vol;
    }

    // BVH::BVH + build_bvh -- bvh.cpp:19-43,58-66
    void build(const std::vector<Tri>* triangles, int depth_limit, int leaf_limit)
    {
        tris = triangles;
        max_depth = depth_limit;
        leaf_max = leaf_limit;
        V3 lo = v3(INFINITY, INFINITY, INFINITY), hi = v3(-INFINITY, -INFINITY, -INFINITY);
        for (const Tri& t : *tris) {
            lo = vmin(lo, vmin(t.a, vmin(t.b, t.c)));   // min(min(min(lo,a),b),c) == this for non-NaN input
            hi = vmax(hi, vmax(t.a, vmax(t.b, t.c)));
        }
        root = new Node(lo, hi);
        for (size_t i = 0; i < tris->size(); i++) insert(root, (int)i, 0);
        compute_volume(root);
    }

    // OctreeNode::intersect(ray, hit, t_near, denoms, numers) -- bvh.h:228-287
    bool intersect_node(const Node* n, const RayT& ray, Hit& hit, float& t_near, const float* denoms,
                        const float* numers, Counters* cnt) const
    {
        float t_far, trash;
        if (cnt) cnt->volume_tests++;
        if (!n->vol.intersect(trash, t_far, denoms, numers)) return false;

        if (n->leaf) {
            for (int ti : n->tris) {
                Hit local;
                if (cnt) cnt->triangle_tests++;
                if (tri_intersect((*tris)[ti], ti, ray, local))
                    if (local.t < hit.t || hit.t == -1) hit = local;
            }
            t_near = hit.t;
            return t_near > 0;
        }

        std::priority_queue<QueueElement, std::vector<QueueElement>, QueueGreater> queue;
        for (int i = 0; i < 8; i++) {
            float d;
            if (cnt) cnt->volume_tests++;
            if (n->child[i]->vol.intersect(d, t_far, denoms, numers)) queue.emplace(QueueElement{n->child[i], d});
        }

        float closest = INFINITY, inter = INFINITY;
        while (!queue.empty()) {
            QueueElement top = queue.top();
            queue.pop();
            if (intersect_node(top.node, ray, hit, inter, denoms, numers, cnt)) {
                closest = std::min(closest, inter);
                if (queue.empty() || closest < queue.top().t_near) {
                    t_near = closest;
                    return true;
                }
            }
        }
        if (closest == INFINITY) return false;
        t_near = closest;
        return true;
    }

    // OctreeNode::intersect(ray, hit) + BVH::intersect -- bvh.h:212-226, bvh.cpp:68-71
    bool intersect(const RayT& ray, Hit& hit, Counters* cnt = nullptr) const
    {
        float trash;
        float denoms[PLANES], numers[PLANES];
        for (int i = 0; i < PLANES; i++) {
            denoms[i] = dot(PLANE_NORMALS.n[i], ray.d);
            numers[i] = dot(PLANE_NORMALS.n[i], ray.o);
        }
        return intersect_node(root, ray, hit, trash, denoms, numers, cnt);
    }
};

void tree_stats(const Node* n, int depth, uint64_t* out)
{
    out[0]++;
    if (n->leaf) {
        out[1]++;
        if (n->tris.empty()) out[2]++;
        if ((uint64_t)depth > out[4]) out[4] = depth;
        if (n->tris.size() > out[5]) out[5] = n->tris.size();
        return;
    }
    out[3]++;
    for (int i = 0; i < 8; i++) tree_stats(n->child[i], depth + 1, out);
}

// ---------------------------------------------------------------------------------------------------------------
// image.h : float RGBA texels, nearest sampling with clamp addressing.
struct Tex {
    std::vector<float> px; // rgba
    int w = 0, h = 0;
    Col texel(int x, int y) const                                                   // image.h:122-134 (offset)
    {
        int cx = x;
        if (cx < 0) cx = 0;
        if (cx > w - 1) cx = w - 1;
        int cy = y;
        if (cy < 0) cy = 0;
        if (cy > h - 1) cy = h - 1;
        const float* p = &px[4 * ((size_t)cy * w + cx)];
        return Col{p[0], p[1], p[2]};
    }
    Col texture_floor(float x, float y) const                                       // image.h:79-86,94-97
    {
        float u = std::floor(x * w);
        float v = std::floor(y * h);
        return texel((int)u, (int)v);
    }
    Col texture_bilinear(float xn, float yn) const                                  // image.h:66-77,89-92
    {
        float x = xn * w, y = yn * h;
        float u = x - std::floor(x);
        float v = y - std::floor(y);
        int ix = x;
        int iy = y;
        return cadd(cadd(cadd(cscale(texel(ix, iy), (1 - u) * (1 - v)), cscale(texel(ix + 1, iy), u * (1 - v))),
                         cscale(texel(ix, iy + 1), (1 - u) * v)), cscale(texel(ix + 1, iy + 1), u * v));
    }
};

// xorshift.h:37-65
struct XorShift {
    uint32_t state = 1;
    uint32_t next()
    {
        uint32_t x = state;
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        return state = x;
    }
    float bilateral() { return next() / (float)std::numeric_limits<uint32_t>::max() * 2 - 1; }
};

uint32_t pixel_seed(uint32_t pixel_index, uint32_t rng_seed)
{
    uint32_t x = pixel_index ^ rng_seed;
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x | 1u;
}

struct RenderCounters {
    uint64_t primary_rays = 0, shadow_rays = 0, reflection_rays = 0, reflection_shadow_rays = 0, primary_hits = 0;
    Counters primary, shadow, reflection;
    void add(const RenderCounters& o)
    {
        primary_rays += o.primary_rays; shadow_rays += o.shadow_rays; reflection_rays += o.reflection_rays;
        reflection_shadow_rays += o.reflection_shadow_rays; primary_hits += o.primary_hits;
        primary.volume_tests += o.primary.volume_tests; primary.triangle_tests += o.primary.triangle_tests;
        shadow.volume_tests += o.shadow.volume_tests; shadow.triangle_tests += o.shadow.triangle_tests;
        reflection.volume_tests += o.reflection.volume_tests; reflection.triangle_tests += o.reflection.triangle_tests;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// renderer.h / renderer.cpp : the ray-tracing part.
struct Renderer {
    RtSettings s;
    float fov = 45.f;
    std::vector<Tri> tris;
    Octree* bvh = nullptr;
    std::vector<RtMaterial> mats;
    Tex tex[RT_TEX_COUNT];
    V3 light = {3, 3, 2};                                                           // scene/light.h:8
    V3 cam_pos = {0, 0, 0};
    M4 cam_to_world = identity();

    // Renderer::_analytic_shapes (renderer.h:331), in the order of add_analytic_shape
    struct Shape {
        int kind;            // 0 = Sphere, 1 = Plane
        V3 a;                // centre / point
        V3 n;                // plane normal
        float radius2;       // Sphere::_radius2 = radius * radius (analyticShape.cpp:7)
        int mat;
    };
    std::vector<Shape> shapes;

    ~Renderer() { delete bvh; }

    // Sphere::intersect / Plane::intersect -- analyticShape.cpp:9-76.  hit is the record the caller hands in (the reference
    // writes t into it even when it then returns false); only t, normal, mat of a successful test are read afterwards.
    static bool shape_intersect(const Shape& sh, const RayT& ray, Hit& hit)
    {
        if (sh.kind == 0) {
            V3 L = sub(ray.o, sh.a);
            constexpr float a = 1;
            float b = 2 * dot(ray.d, L);
            float c = dot(L, L) - sh.radius2;
            float delta = b * b - 4 * a * c;
            if (delta < 0) return false;
            constexpr float a2 = 2 * a;
            if (delta == 0.0) hit.t = -b / a2;
            else {
                float sqrt_delta = std::sqrt(delta);
                float t1 = (-b - sqrt_delta) / a2;
                float t2 = (-b + sqrt_delta) / a2;
                if (t1 < t2) {
                    hit.t = t1;
                    if (hit.t < 0) hit.t = t2;
                }
            }
            if (hit.t < 0) return false;
            hit.normal = normalize(sub(add(ray.o, scale(hit.t, ray.d)), sh.a));
            hit.mat = sh.mat;
            return true;
        }
        float t = dot(sub(sh.a, ray.o), sh.n) / dot(ray.d, sh.n);
        if (t < 0) return false;
        hit.t = t;
        hit.mat = sh.mat;
        hit.normal = sh.n;
        return true;
    }

    static constexpr float EPSILON = 1.0e-4f;                                       // renderer.h:23
    static constexpr float SHADOW_INTENSITY = 0.5f;                                 // renderer.h:24

    void tex_coords(const Hit& h, float u, float v, float& tu, float& tv) const     // renderer.cpp:436-445
    {
        if (h.frag) tri_texcoords(*h.frag, u, v, tu, tv);
        else if (h.tri >= 0) tri_texcoords(tris[h.tri], u, v, tu, tv);
        else { tu = u; tv = v; }
    }

    // Renderer::normal_mapping -- renderer.cpp:464-478 ; Transform(x,y,z,w) columns -- mat.cpp:69-75
    V3 normal_mapping(const Hit& h, float u, float v) const
    {
        float tu, tv;
        tex_coords(h, u, v, tu, tv);
        V3 tangent = h.tangent;
        V3 bitangent = cross(tangent, h.normal);
        Col nc = tex[RT_TEX_NORMAL].texture_floor(tu, tv);
        V3 nm = sub(scale(2, v3(nc.r, nc.g, nc.b)), v3(1, 1, 1));
        V3 q = normalize(nm);
        V3 p = v3(tangent.x * q.x + bitangent.x * q.y + h.normal.x * q.z,
                  tangent.y * q.x + bitangent.y * q.y + h.normal.y * q.z,
                  tangent.z * q.x + bitangent.z * q.y + h.normal.z * q.z);
        return normalize(p);
    }

    // Renderer::is_shadowed -- renderer.cpp:340-402 (BVH branch, then the analytic shapes)
    bool is_shadowed(V3 p, V3 n, RenderCounters* rc, bool secondary) const
    {
        if (!s.compute_shadows) return false;
        RayT ray{add(p, scale(EPSILON, n)), normalize(sub(light, p))};
        Hit h;
        Counters* c = nullptr;
        if (rc) {
            if (secondary) { rc->reflection_shadow_rays++; c = &rc->reflection; }
            else { rc->shadow_rays++; c = &rc->shadow; }
        }
        if (bvh->intersect(ray, h, c)) {
            V3 q = add(ray.o, scale(h.t, ray.d));
            if (length2(sub(p, q)) < length2(sub(p, light))) return true;
        }
        for (const Shape& sh : shapes)                                               // :376-397, the same HitInfo all along
            if (shape_intersect(sh, ray, h)) {
                V3 q = add(ray.o, scale(h.t, ray.d));
                if (length2(sub(p, q)) < length2(sub(p, light))) return true;
            }
        return false;
    }

    // Renderer::compute_specular -- renderer.cpp:270-280
    Col specular(const RtMaterial& m, V3 ray_dir, V3 n, V3 to_light) const
    {
        V3 half = normalize(sub(to_light, ray_dir));
        float angle = dot(half, n);
        if (angle <= m.specular_threshold) return Col{0, 0, 0};
        float p = std::pow(std::max(0.0f, angle), m.ns);
        return Col{m.specular[0] * p, m.specular[1] * p, m.specular[2] * p};
    }

    // Renderer::compute_reflection -- renderer.cpp:283-338.  `rh` lives across the samples exactly as the
    // reference's reflection_hit_info does (declared outside the loop, :286).
    Col reflection(const RayT& ray, V3 p, const Hit& hit, int depth, XorShift& rng, RenderCounters* rc) const
    {
        bool found = false;
        Hit rh;
        const RtMaterial& m = mats[hit.mat];
        V3 n = hit.normal;
        V3 origin = add(p, scale(0.01f, n));
        V3 mirror = sub(ray.d, scale(2 * dot(ray.d, n), n));
        int samples = 0;
        Col total = col(0.0f);
        for (int i = 0; i < s.rough_reflections_sample_count; i++) {
            float roughness;
            if (s.enable_roughness_mapping) {
                float tu, tv;
                tex_coords(hit, hit.u, hit.v, tu, tv);
                roughness = tex[RT_TEX_ROUGHNESS].texture_floor(tu, tv).r;
            } else
                roughness = m.roughness;
            if (roughness > 0) {
                // Vector(rand(), rand(), rand()) at renderer.cpp:313: g++ evaluates the three arguments
                // right to left (pinned against the compiled reference by tests/test_oracle_vs_reference.py).
                float rz = rng.bilateral();
                float ry = rng.bilateral();
                float rx = rng.bilateral();
                V3 rd = normalize(v3(rx, ry, rz));
                if (dot(rd, hit.normal) < 0) rd = neg(rd);
                V3 lerped = add(scale(roughness, rd), scale(1 - roughness, mirror));
                total = cadd(total, trace_ray(RayT{origin, lerped}, rh, depth + 1, found, rng, rc, true));
                samples++;
            } else {
                total = cadd(total, trace_ray(RayT{origin, mirror}, rh, depth + 1, found, rng, rc, true));
                samples = 1;
                break;
            }
        }
        return cmul(cdiv(total, col((float)samples)), col(m.reflection));
    }

    // Renderer::parallax_occlusion_mapping -- renderer.cpp:518-554
    void parallax_occlusion_mapping(const Hit& hit, float u, float v, V3 view_dir, float& new_u, float& new_v) const
    {
        const Tex& dm = tex[RT_TEX_DISPLACEMENT];
        float tex_coord_u, tex_coord_v;
        tex_coords(hit, u, v, tex_coord_u, tex_coord_v);
        float current_depth;
        float depth_step = 1.0f / s.parallax_mapping_steps;
        float sampled_depth = dm.texture_floor(tex_coord_u, tex_coord_v).r;
        V3 search_direction = scale(s.displacement_mapping_strength, neg(view_dir));
        float delta_u = search_direction.x / s.parallax_mapping_steps;
        float delta_v = search_direction.y / s.parallax_mapping_steps;
        current_depth = 0.0f;
        new_u = tex_coord_u;
        new_v = tex_coord_v;
        while (current_depth < sampled_depth) {
            new_u += delta_u;
            new_v += delta_v;
            sampled_depth = dm.texture_floor(new_u, new_v).r;
            current_depth += depth_step;
        }
        float previous_u_coord = new_u - delta_u;
        float previous_v_coord = new_v - delta_v;
        float after_depth = sampled_depth - current_depth;
        float before_depth = dm.texture_floor(previous_u_coord, previous_v_coord).r - (current_depth - depth_step);
        float interpolation_weight = after_depth / (after_depth - before_depth);
        new_u = (1 - interpolation_weight) * new_u + interpolation_weight * previous_u_coord;
        new_v = (1 - interpolation_weight) * new_v + interpolation_weight * previous_v_coord;
    }

    // Renderer::shade_ray_inter_point -- renderer.cpp:556-617
    Col shade(const RayT& ray, Hit& hit, int depth, XorShift& rng, RenderCounters* rc, bool secondary) const
    {
        Col c = Col{0, 0, 0};
        if (s.shading_method == RT_SHADING) {
            float u = hit.u, v = hit.v;
            V3 p = add(ray.o, scale(hit.t, ray.d));
            if (s.enable_displacement_mapping) parallax_occlusion_mapping(hit, hit.u, hit.v, normalize(sub(cam_pos, p)), u, v);   // :567-568
            V3 to_light = normalize(sub(light, p));
            if (s.enable_normal_mapping) hit.normal = normal_mapping(hit, u, v);
            RtMaterial m = mats[hit.mat];
            float ao = 1.0f;
            if (s.enable_ao_mapping) {
                float tu, tv;
                tex_coords(hit, u, v, tu, tv);
                ao = tex[RT_TEX_AO].texture_floor(tu, tv).r;
            }
            Col diffuse;
            if (s.enable_diffuse_mapping) {
                float tu, tv;
                tex_coords(hit, u, v, tu, tv);
                diffuse = tex[RT_TEX_DIFFUSE].texture_floor(tu, tv);
                diffuse = cmul(diffuse, col(std::max(0.5f, dot(hit.normal, normalize(sub(cam_pos, p))))));
            } else {
                float k = std::max(0.0f, dot(hit.normal, to_light));                 // compute_diffuse :263-266
                diffuse = Col{m.diffuse[0] * k, m.diffuse[1] * k, m.diffuse[2] * k};
            }
            c = cadd(c, cscale(cscale(diffuse, ao), s.enable_diffuse ? 1.0f : 0.0f));
            c = cadd(c, cscale(specular(m, ray.d, hit.normal, to_light), s.enable_specular ? 1.0f : 0.0f));
            if (is_shadowed(p, hit.normal, rc, secondary)) c = cmul(c, col(SHADOW_INTENSITY));
            c = cadd(c, cscale(Col{m.emission[0], m.emission[1], m.emission[2]}, s.enable_emissive ? 1.0f : 0.0f));
            if (m.reflection > 0.0f) c = cadd(c, cscale(reflection(ray, p, hit, depth, rng, rc), m.reflection));
            Col amb = cmul(col(0.1f), Col{m.ambient_coeff[0], m.ambient_coeff[1], m.ambient_coeff[2]}); // AMBIENT_COLOR :18
            c = cadd(c, cscale(cscale(amb, 1 - m.reflection), s.enable_ambient ? 1.0f : 0.0f));
        } else if (s.shading_method == RT_ABS_NORMALS_SHADING) {                     // :404-407
            c = Col{std::abs(hit.normal.x), std::abs(hit.normal.y), std::abs(hit.normal.z)};
        } else if (s.shading_method == RT_PASTEL_NORMALS_SHADING) {                  // :409-412 (Color * double 0.5)
            c = cscale(cadd(Col{hit.normal.x, hit.normal.y, hit.normal.z}, col(1.0f)), 0.5f);
        } else if (s.shading_method == RT_BARYCENTRIC_COORDINATES_SHADING) {         // :414-417
            c = cadd(cadd(cscale(Col{1, 0, 0}, hit.u), cscale(Col{0, 1, 0}, hit.v)), cscale(Col{0, 0, 1}, 1 - hit.u - hit.v));
        } else if (s.shading_method == RT_VISUALIZE_AO) {                            // :419-434
            c = Col{0.9f, 0.9f, 0.9f};
            if (s.enable_ao_mapping) {
                float tu, tv;
                tri_texcoords(hit.frag ? *hit.frag : tris[hit.tri], hit.u, hit.v, tu, tv);
                c = cmul(c, col(tex[RT_TEX_AO].texture_floor(tu, tv).r));
            }
        }
        c.r = std::clamp(c.r, 0.0f, 1.0f);
        c.g = std::clamp(c.g, 0.0f, 1.0f);
        c.b = std::clamp(c.b, 0.0f, 1.0f);
        return c;
    }

    // Skybox::sample -- renderer/skybox.cpp:12-51 (faces: right, left, top, bottom, back, front)
    Col skybox_sample(V3 direction) const
    {
        V3 direction2 = v3(direction.x, direction.y, -direction.z);
        V3 dir_abs = v3(std::abs(direction2.x), std::abs(direction2.y), std::abs(direction2.z));
        int face_index;
        float norm_factor;
        float u, v;
        if (dir_abs.z >= dir_abs.x && dir_abs.z >= dir_abs.y) {
            face_index = direction2.z < 0.0 ? 4.0 : 5.0;
            norm_factor = 0.5 / dir_abs.z;
            u = direction2.z < 0.0 ? -direction2.x : direction2.x;
            v = -direction2.y;
        } else if (dir_abs.y >= dir_abs.x) {
            face_index = direction2.y < 0.0 ? 3.0 : 2.0;
            norm_factor = 0.5 / dir_abs.y;
            u = direction2.x;
            v = direction2.y < 0.0 ? -direction2.z : direction2.z;
        } else {
            face_index = direction2.x < 0.0 ? 1.0 : 0.0;
            norm_factor = 0.5 / dir_abs.x;
            u = direction2.x < 0.0 ? direction2.z : -direction2.z;
            v = -direction2.y;
        }
        u = u * norm_factor + 0.5;
        v = v * norm_factor + 0.5;
        return tex[RT_TEX_SKYBOX_RIGHT + face_index].texture_bilinear(u, v);
    }

    // Renderer::trace_ray -- renderer.cpp:1008-1066 (BVH branch, then the analytic shapes)
    Col trace_ray(const RayT& ray, Hit& final_hit, int depth, bool& found, XorShift& rng, RenderCounters* rc, bool secondary) const
    {
        Hit local;
        if (depth > s.max_recursion_depth) return col(0.0f);
        Counters* c = nullptr;
        if (rc) {
            if (secondary) { rc->reflection_rays++; c = &rc->reflection; }
            else { rc->primary_rays++; c = &rc->primary; }
        }
        if (bvh->intersect(ray, local, c))
            if (local.t < final_hit.t || final_hit.t == -1) final_hit = local;
        for (const Shape& sh : shapes)                                               // :1029-1037
            if (shape_intersect(sh, ray, local))
                if (local.t < final_hit.t || final_hit.t == -1) final_hit = local;
        float min_t = 0.1;
        if (final_hit.t > min_t) {
            found = true;
            if (rc && !secondary) rc->primary_hits++;
            Col out = shade(ray, final_hit, depth, rng, rc, secondary);
            out.r = std::clamp(out.r, 0.0f, 1.0f);
            out.g = std::clamp(out.g, 0.0f, 1.0f);
            out.b = std::clamp(out.b, 0.0f, 1.0f);
            return out;
        }
        if (s.enable_skysphere) {                                                    // :1054-1060
            float u = 0.5 + std::atan2(-ray.d.z, -ray.d.x) / (2 * M_PI);
            float v = 0.5 + std::asin(-ray.d.y) / M_PI;
            return tex[RT_TEX_SKYSPHERE].texture_floor(u, v);
        }
        if (s.enable_skybox) return skybox_sample(ray.d);                            // :1061-1062
        return Col{135.0f / 255.0f, 206.0f / 255.0f, 235.0f / 255.0f};               // BACKGROUND_COLOR :19
    }

    void super_dims(int& rw, int& rh) const                                         // renderer.cpp:116-120
    {
        rw = s.enable_ssaa ? s.image_width * s.ssaa_factor : s.image_width;
        rh = s.enable_ssaa ? s.image_height * s.ssaa_factor : s.image_height;
    }
};

// ImageUtils::gkit_color_to_Qt_ARGB32_uint + qRgb -- imageUtils.h:149-152
inline uint32_t quantise(Col c)
{
    int r = (int)(c.r * 255), g = (int)(c.g * 255), b = (int)(c.b * 255);
    return 0xff000000u | (((uint32_t)r & 0xffu) << 16) | (((uint32_t)g & 0xffu) << 8) | ((uint32_t)b & 0xffu);
}

// Pixel loop of Renderer::ray_trace -- renderer.cpp:1082-1115
double trace_rows(const Renderer& r, const M4& c2w, uint32_t* argb_super, int row_begin, int row_end, int row_step,
                  int threads, RenderCounters* total, float* zbuf = nullptr, V3* nbuf = nullptr)
{
    int rw, rh;
    r.super_dims(rw, rh);
    M4 proj_inv = inverse(perspective(r.fov, (float)rw / rh, 0.1f, 1000.0f));      // scene/camera.cpp:5-11, camera.h:11
    V3 cam_pos = xform_point(c2w, v3(0, 0, 0));                                     // renderer.cpp:232
    if (threads <= 0) threads = omp_get_max_threads();
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(threads)
    {
        RenderCounters local;
#pragma omp for schedule(dynamic)
        for (int py = row_begin; py < row_end; py += row_step) {
            float yw = ((float)py + 0.5f) / rh * 2 - 1;
            for (int px = 0; px < rw; px++) {
                float xw = ((float)px + 0.5f) / rw * 2 - 1;
                V3 vs = xform_point(proj_inv, v3(xw, yw, -1));
                V3 ws = xform_point(c2w, vs);
                RayT ray{cam_pos, normalize(sub(ws, cam_pos))};
                XorShift rng;
                rng.state = pixel_seed((uint32_t)(py * rw + px), r.s.rng_seed);
                bool found = false;
                Hit hit;
                Col c = r.trace_ray(ray, hit, 0, found, rng, total ? &local : nullptr, false);
                if (found && zbuf) {                                                 // renderer.cpp:1104-1111
                    zbuf[(size_t)py * rw + px] = -(ray.o.z + ray.d.z * hit.t);
                    nbuf[(size_t)py * rw + px] = hit.normal;
                }
                if (argb_super) argb_super[(size_t)py * rw + px] = quantise(c);
            }
        }
        if (total) {
#pragma omp critical
            total->add(local);
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// ImageUtils::downscale_image_qt_ARGB32 -- imageUtils.h:98-147
void downscale(const uint32_t* in, int w, int h, int factor, uint32_t* out)
{
    int dw = w / factor, dh = h / factor;
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            int ar = 0, ag = 0, ab = 0;
            for (int i = 0; i < factor; i++)
                for (int j = 0; j < factor; j++) {
                    uint32_t c = in[(size_t)(y * factor + i) * w + (x * factor + j)];
                    ar += (c >> 16) & 0xff;
                    ag += (c >> 8) & 0xff;
                    ab += c & 0xff;
                }
            ar = ar / (factor * factor);
            ag = ag / (factor * factor);
            ab = ab / (factor * factor);
            out[(size_t)y * dw + x] = 0xff000000u | ((uint32_t)(ar & 0xff) << 16) | ((uint32_t)(ag & 0xff) << 8) | (uint32_t)(ab & 0xff);
        }
}


// ---------------------------------------------------------------------------------------------------------------------
// Renderer::post_process_ssao_SIMD -- renderer.cpp:1229-1434, with its helpers SIMD/m256Vector.cpp:81-110,
// SIMD/m256Point.cpp:3-31 and renderer/xorshift.h:8-65.  The reference processes 8 pixels per AVX2 register; every
// lane's arithmetic is restated here as scalar float operations in the same order (one explicit fmaf per _mm256_fmadd_ps),
// so one pixel costs one call of ssao_simd_pixel.  The columns the last partial group of a row leaves over go through
// the reference's scalar loop (renderer.cpp:1358-1407), which uses other formulas (double arithmetic and truncation for
// the sampled pixel, unsigned random numbers): ssao_scalar_pixel.
//
// Random numbers.  The reference seeds one 8-lane generator and one scalar generator per OpenMP thread from std::rand()
// and the thread number (renderer.cpp:1254-1266) and draws from them in processing order, which is not reproducible.
// Two disciplines are offered here:
//   SSAO_RNG_REFERENCE_ORDER  one 8-lane generator + one scalar generator with caller-given seeds, rows and 8-pixel groups
//                             in the order of a single-threaded run, groups without geometry skipped (renderer.cpp:1277)
//                             -- what the reference does with OMP_NUM_THREADS=1 after srand(); pins the arithmetic against
//                             the compiled reference (tests/test_oracle_vs_reference.py);
//   SSAO_RNG_PER_PIXEL        every pixel owns a generator seeded pixel_seed(py * W' + px, rng_seed + kSsaoSeedOffset):
//                             the shared stream of the CUDA path (same arithmetic, order-independent).
constexpr uint32_t kSsaoSeedOffset = 0x9e3779b9u;
enum { SSAO_RNG_PER_PIXEL = 0, SSAO_RNG_REFERENCE_ORDER = 1 };

inline uint32_t xs_next(uint32_t& st)                                                 // xorshift.h:13-22,43-52
{
    uint32_t x = st;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    return st = x;
}
// __m256_XorShiftGenerator::get_rand_bilateral / get_rand_lateral -- xorshift.h:24-32 (_mm256_cvtepi32_ps is SIGNED)
inline float simd_bilateral(uint32_t& st) { return (float)(int32_t)xs_next(st) / (float)std::numeric_limits<int32_t>::max(); }
inline float simd_lateral(uint32_t& st) { return ((float)(int32_t)xs_next(st) / (float)std::numeric_limits<int32_t>::max() + 1.0f) * 0.5f; }
// XorShiftGenerator::get_rand_bilateral / get_rand_lateral -- xorshift.h:54-62
inline float scalar_bilateral(uint32_t& st) { return xs_next(st) / (float)std::numeric_limits<uint32_t>::max() * 2 - 1; }
inline float scalar_lateral(uint32_t& st) { return xs_next(st) / (float)std::numeric_limits<uint32_t>::max(); }

// _mm256_cvtps_epi32: round to nearest even; NaN and out-of-range give the "integer indefinite" 0x80000000
inline int32_t cvtps_epi32(float f)
{
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return std::numeric_limits<int32_t>::min();
    return (int32_t)std::nearbyintf(f);
}
// (int) of a double as x86 does it (cvttsd2si): truncation, indefinite when out of range
inline int32_t cvttsd_epi32(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return std::numeric_limits<int32_t>::min();
    return (int32_t)v;
}

struct SsaoFrame {
    int rw, rh;
    const float* z;
    const V3* n;
    M4 proj;                     // Camera::_perspective_proj_mat
    float aspect;
    float fov_mult_simd;         // (float)std::tan(fov / 2 / 180 * M_PI), renderer.cpp:1249
    float fov_mult_scalar;       // std::tan(radians(fov / 2)) in float, renderer.cpp:1369
    int samples;
    float radius;
};

// One lane of the AVX2 loop body, renderer.cpp:1283-1355.  Returns the lane's pixel_occlusion.
int ssao_simd_pixel(const SsaoFrame& f, int x, int y, uint32_t& rng)
{
    const float view_z = f.z[(size_t)y * f.rw + x];
    const bool has_geometry = view_z != INFINITY;                                   // infinity_mask (_CMP_NEQ_OQ)
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
    const float n_len = std::sqrt(nb.x * nb.x + (nb.y * nb.y + nb.z * nb.z));      // _mm256_length: x + (y + z)
    const float n_inv = 1.0f / n_len;
    const V3 normal = v3(nb.x * n_inv, nb.y * n_inv, nb.z * n_inv);
    int occlusion = 0;
    for (int i = 0; i < f.samples; i++) {
        const float rx = simd_bilateral(rng), ry = simd_bilateral(rng), rz = simd_bilateral(rng);
        const float r_inv = 1.0f / std::sqrt(rx * rx + (ry * ry + rz * rz));
        V3 rs = v3(rx * r_inv, ry * r_inv, rz * r_inv);
        const float k = simd_lateral(rng) + 0.0001f;
        rs = v3(rs.x * k, rs.y * k, rs.z * k);
        rs = v3(rs.x * f.radius, rs.y * f.radius, rs.z * f.radius);
        rs = v3(rs.x + P.x, rs.y + P.y, rs.z + P.z);
        const V3 vd = v3(rs.x - P.x, rs.y - P.y, rs.z - P.z);
        const float d = vd.x * normal.x + (vd.y * normal.y + vd.z * normal.z);       // _mm256_dot_product: x + (y + z)
        const float flip = d < 0.0f ? 1.0f : 0.0f;                                   // _CMP_LT_OQ & ones
        const V3 back = v3((P.x - rs.x) * 2.0f, (P.y - rs.y) * 2.0f, (P.z - rs.z) * 2.0f);
        rs = v3(rs.x + back.x * flip, rs.y + back.y * flip, rs.z + back.z * flip);
        const float (*m)[4] = f.proj.m;                                              // __m256Point::transform
        const float xt = std::fmaf(m[0][0], rs.x, std::fmaf(m[0][1], rs.y, std::fmaf(m[0][2], rs.z, m[0][3])));
        const float yt = std::fmaf(m[1][0], rs.x, std::fmaf(m[1][1], rs.y, std::fmaf(m[1][2], rs.z, m[1][3])));
        const float wt = std::fmaf(m[3][0], rs.x, std::fmaf(m[3][1], rs.y, std::fmaf(m[3][2], rs.z, m[3][3])));
        const float w = 1.0f / wt;
        const float ndc_x = xt * w, ndc_y = yt * w;
        int32_t px = cvtps_epi32(((ndc_x + 1.0f) * 0.5f) * (float)f.rw);
        int32_t py = cvtps_epi32(((ndc_y + 1.0f) * 0.5f) * (float)f.rh);
        px = std::max(std::min(px, cvtps_epi32((float)f.rw - 1.0f)), 0);
        py = std::max(std::min(py, cvtps_epi32((float)f.rh - 1.0f)), 0);
        const float sample_geometry_depth = -1.0f * f.z[(size_t)px + (size_t)py * f.rw];
        const bool in_range = std::fabs(sample_geometry_depth - P.z) <= f.radius;   // _CMP_LE_OQ
        const bool behind = rs.z < sample_geometry_depth;                            // _CMP_LT_OQ
        if (in_range && behind && has_geometry) occlusion++;
    }
    return occlusion;
}

// The scalar loop for the left-over columns, renderer.cpp:1358-1407 (called for pixels with geometry only).
int ssao_scalar_pixel(const SsaoFrame& f, int x, int y, uint32_t& rng)
{
    float x_ndc = (float)x / f.rw * 2 - 1;
    float y_ndc = (float)y / f.rh * 2 - 1;
    float view_z = f.z[(size_t)y * f.rw + x];
    float view_ray_x = x_ndc * f.aspect * f.fov_mult_scalar;
    float view_ray_y = y_ndc * f.fov_mult_scalar;
    V3 P = v3(view_ray_x * view_z, view_ray_y * view_z, -view_z);
    V3 normal = normalize(f.n[(size_t)y * f.rw + x]);
    int occlusion = 0;
    for (int i = 0; i < f.samples; i++) {
        float rx = scalar_bilateral(rng);
        float ry = scalar_bilateral(rng);
        float rz = scalar_bilateral(rng);
        V3 rs = normalize(v3(rx, ry, rz));
        rs = scale(scalar_lateral(rng) + 0.0001f, rs);
        rs = scale(f.radius, rs);
        rs = add(rs, P);
        if (dot(sub(rs, P), normal) < 0) rs = add(rs, scale(2, sub(P, rs)));
        V3 ndc = xform_point(f.proj, rs);
        int px = cvttsd_epi32((ndc.x + 1) * 0.5 * f.rw);
        int py = cvttsd_epi32((ndc.y + 1) * 0.5 * f.rh);
        px = std::min(std::max(0, px), f.rw - 1);
        py = std::min(std::max(0, py), f.rh - 1);
        float sample_geometry_depth = -f.z[(size_t)py * f.rw + px];
        if (std::abs(sample_geometry_depth - P.z) > f.radius) continue;
        if (rs.z < sample_geometry_depth) occlusion++;
    }
    return occlusion;
}

// ao[] for the whole frame, then the 7x7 blur applied to the image (renderer.cpp:1411-1431).
void ssao_post_process(const SsaoFrame& f, uint32_t* argb_super, uint32_t rng_seed, float amount, int rng_mode, const uint32_t* ref_seeds9)
{
    std::vector<int> ao((size_t)f.rw * f.rh, 0);
    const int leftover = f.rw % 8;
    if (rng_mode == SSAO_RNG_REFERENCE_ORDER) {
        uint32_t lanes[8], scalar = ref_seeds9[8];
        for (int l = 0; l < 8; l++) lanes[l] = ref_seeds9[l];
        for (int y = 0; y < f.rh; y++) {
            for (int x = 0; x + 8 <= f.rw; x += 8) {                                 // (the reference's partial group reads past the row)
                bool any = false;
                for (int l = 0; l < 8; l++) any |= f.z[(size_t)y * f.rw + x + l] != INFINITY;
                if (!any) continue;                                                  // renderer.cpp:1276-1278
                // the 8 lanes draw in lockstep, each from its own state: lane by lane is the same sequence per lane
                for (int l = 0; l < 8; l++) ao[(size_t)y * f.rw + x + l] = ssao_simd_pixel(f, x + l, y, lanes[l]);
            }
            for (int x = f.rw - leftover; x < f.rw; x++) {
                if (f.z[(size_t)y * f.rw + x] == INFINITY) continue;
                ao[(size_t)y * f.rw + x] = ssao_scalar_pixel(f, x, y, scalar);
            }
        }
    } else {
#pragma omp parallel for schedule(dynamic)
        for (int y = 0; y < f.rh; y++)
            for (int x = 0; x < f.rw; x++) {
                if (f.z[(size_t)y * f.rw + x] == INFINITY) continue;
                uint32_t st = pixel_seed((uint32_t)(y * f.rw + x), rng_seed + kSsaoSeedOffset);
                ao[(size_t)y * f.rw + x] = x < f.rw - leftover ? ssao_simd_pixel(f, x, y, st) : ssao_scalar_pixel(f, x, y, st);
            }
    }
    const int blur_size = 7, half = blur_size / 2;
    for (int y = half; y < f.rh - half; y++)
        for (int x = half; x < f.rw - half; x++) {
            if (f.z[(size_t)y * f.rw + x] == INFINITY) continue;
            int sum = 0;
            for (int oy = -half; oy <= half; oy++)
                for (int ox = -half; ox <= half; ox++) sum += ao[(size_t)(y + oy) * f.rw + x + ox];
            float mult = 1 - ((float)sum / (float)(blur_size * blur_size) / (float)f.samples * (float)amount);
            uint32_t c = argb_super[(size_t)y * f.rw + x];
            int r = (c >> 16) & 0xff, g = (c >> 8) & 0xff, b = c & 0xff;
            // QColor(int, int, int) from float expressions: truncation (renderer.cpp:1429)
            int nr = (int)(r * mult), ng = (int)(g * mult), nb = (int)(b * mult);
            argb_super[(size_t)y * f.rw + x] = 0xff000000u | (((uint32_t)nr & 0xffu) << 16) | (((uint32_t)ng & 0xffu) << 8) | ((uint32_t)nb & 0xffu);
        }
}

// ray_trace() with the G-buffers, post_process_ssao_SIMD(), apply_ssaa() -- renderer.cpp:1068-1135
void render_with_ssao(const Renderer& r, uint32_t* argb_out, int threads, int rng_mode, const uint32_t* ref_seeds9)
{
    int rw, rh;
    r.super_dims(rw, rh);
    std::vector<uint32_t> super((size_t)rw * rh);
    std::vector<float> z((size_t)rw * rh, INFINITY);                                 // clear_z_buffer, renderer.cpp:165-168
    std::vector<V3> n((size_t)rw * rh, v3(0, 0, 0));                                 // clear_normal_buffer, :170-173
    trace_rows(r, r.cam_to_world, super.data(), 0, rh, 1, threads, nullptr, z.data(), n.data());
    SsaoFrame f;
    f.rw = rw; f.rh = rh; f.z = z.data(); f.n = n.data();
    f.aspect = (float)rw / rh;                                                       // renderer.cpp:250-261
    f.proj = perspective(r.fov, f.aspect, 0.1f, 1000.0f);
    f.fov_mult_simd = (float)std::tan(r.fov / 2 / 180 * M_PI);
    f.fov_mult_scalar = std::tan(((float)M_PI / 180) * (r.fov / 2));                 // radians(), mat.cpp:13-16; std::tan(float)
    f.samples = r.s.ssao_sample_count; f.radius = r.s.ssao_radius;
    ssao_post_process(f, super.data(), r.s.rng_seed, r.s.ssao_amount, rng_mode, ref_seeds9);
    if (r.s.enable_ssaa) downscale(super.data(), rw, rh, r.s.ssaa_factor, argb_out);
    else memcpy(argb_out, super.data(), super.size() * sizeof(uint32_t));
}

// ---------------------------------------------------------------------------------------------------------------------
// Renderer::raster_trace -- renderer.cpp:869-1006, with clip_triangle / clip_triangles_to_plane (:630-853),
// matrix_transform_z (:855-867) and trace_triangle (:619-628).  The reference walks the triangles under
// `omp parallel for` with an unsynchronised z-buffer; the restatement is the sequential order (what one thread does):
// a fragment replaces the pixel iff its depth is STRICTLY smaller, so the pixel ends up with the first fragment, in
// (triangle, clipped piece) order, of the smallest depth.
struct V4 { float x, y, z, w; };
inline V4 v4(float x, float y, float z, float w) { return V4{x, y, z, w}; }
inline V4 add4(V4 u, V4 v) { return v4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w); }  // vec.cpp:123-126
inline V4 sub4(V4 u, V4 v) { return v4(u.x - v.x, u.y - v.y, u.z - v.z, u.w - v.w); }  // vec.cpp:128-131
inline V4 scale4(float t, V4 u) { return v4(u.x * t, u.y * t, u.z * t, u.w * t); }     // vec.cpp:133-141
inline float comp4(const V4& v, int i) { return (&v.x)[i]; }

// Transform::operator()(const vec4&) -- mat.cpp:119-132
V4 xform_v4(const M4& t, V4 v)
{
    float x = v.x, y = v.y, z = v.z, w = v.w;
    float xt = t.m[0][0] * x + t.m[0][1] * y + t.m[0][2] * z + t.m[0][3] * w;
    float yt = t.m[1][0] * x + t.m[1][1] * y + t.m[1][2] * z + t.m[1][3] * w;
    float zt = t.m[2][0] * x + t.m[2][1] * y + t.m[2][2] * z + t.m[2][3] * w;
    float wt = t.m[3][0] * x + t.m[3][1] * y + t.m[3][2] * z + t.m[3][3] * w;
    return v4(xt, yt, zt, wt);
}

struct Tri4 {                                                                       // Triangle4, triangle.h:24-40
    V4 a, b, c;
    V3 tu, tv;
};
constexpr int kClipMax = 12;                                                        // std::array<Triangle4, 12>, renderer.cpp:881-882

// is_inside<plane_index, plane_sign> -- renderer.cpp:630-670
inline bool is_inside(int plane_index, int plane_sign, const V4& v)
{
    return plane_sign > 0 ? comp4(v, plane_index) < v.w : comp4(v, plane_index) > -v.w;
}

// Renderer::clip_triangles_to_plane<plane_index, plane_sign> -- renderer.cpp:672-835 (CLIPPING_EPSILON = 0).  `in` and
// `out` may be the same array (the first stage of clip_triangle, one triangle): every triangle is read before it is written.
int clip_to_plane(int plane_index, int plane_sign, const Tri4* in, int n, Tri4* out)
{
    int added = 0;
    for (int i = 0; i < n; i++) {
        const Tri4 t = in[i];
        const bool ai = is_inside(plane_index, plane_sign, t.a), bi = is_inside(plane_index, plane_sign, t.b), ci = is_inside(plane_index, plane_sign, t.c);
        const int sum = ai + bi + ci;
        const float sign = (float)plane_sign;
        if (sum == 3) {
            if (added < kClipMax) out[added++] = t;
        } else if (sum == 1) {
            V4 in_v, o1, o2;
            float u[3], v[3];                                                       // { inside, outside_1, outside_2 }
            if (ai) { in_v = t.a; o1 = t.b; o2 = t.c; u[0] = t.tu.x; u[1] = t.tu.y; u[2] = t.tu.z; v[0] = t.tv.x; v[1] = t.tv.y; v[2] = t.tv.z; }
            else if (bi) { in_v = t.b; o1 = t.c; o2 = t.a; u[0] = t.tu.y; u[1] = t.tu.z; u[2] = t.tu.x; v[0] = t.tv.y; v[1] = t.tv.z; v[2] = t.tv.x; }
            else { in_v = t.c; o1 = t.a; o2 = t.b; u[0] = t.tu.z; u[1] = t.tu.x; u[2] = t.tu.y; v[0] = t.tv.z; v[1] = t.tv.x; v[2] = t.tv.y; }
            float din = comp4(in_v, plane_index) - in_v.w * sign;
            float d1 = comp4(o1, plane_index) - o1.w * sign;
            float d2 = comp4(o2, plane_index) - o2.w * sign;
            float tp1 = d1 / (d1 - din);
            float tp2 = d2 / (d2 - din);
            V4 p1 = add4(o1, scale4(tp1 - 0.0f, sub4(in_v, o1)));
            V4 p2 = add4(o2, scale4(tp2 - 0.0f, sub4(in_v, o2)));
            Tri4 nt;
            nt.a = in_v; nt.b = p1; nt.c = p2;
            nt.tu = v3(u[0], u[1] + (tp1 - 0.0f) * (u[0] - u[1]), u[2] + (tp2 - 0.0f) * (u[0] - u[2]));
            nt.tv = v3(v[0], v[1] + (tp1 - 0.0f) * (v[0] - v[1]), v[2] + (tp2 - 0.0f) * (v[0] - v[2]));
            if (added < kClipMax) out[added++] = nt;
        } else if (sum == 2) {
            V4 i1, i2, ov;
            float u[3], v[3];                                                       // { inside_1, inside_2, outside }
            if (!ai) { ov = t.a; i1 = t.b; i2 = t.c; u[0] = t.tu.y; u[1] = t.tu.z; u[2] = t.tu.x; v[0] = t.tv.y; v[1] = t.tv.z; v[2] = t.tv.x; }
            else if (!bi) { ov = t.b; i1 = t.c; i2 = t.a; u[0] = t.tu.z; u[1] = t.tu.x; u[2] = t.tu.y; v[0] = t.tv.z; v[1] = t.tv.x; v[2] = t.tv.y; }
            else { ov = t.c; i1 = t.a; i2 = t.b; u[0] = t.tu.x; u[1] = t.tu.y; u[2] = t.tu.z; v[0] = t.tv.x; v[1] = t.tv.y; v[2] = t.tv.z; }
            float d1 = comp4(i1, plane_index) - i1.w * sign;
            float d2 = comp4(i2, plane_index) - i2.w * sign;
            float dout = comp4(ov, plane_index) - ov.w * sign;
            float tp1 = dout / (dout - d1);
            float tp2 = dout / (dout - d2);
            V4 p1 = add4(ov, scale4(tp1 - 0.0f, sub4(i1, ov)));
            V4 p2 = add4(ov, scale4(tp2 - 0.0f, sub4(i2, ov)));
            Tri4 t1, t2;
            t1.a = i1; t1.b = i2; t1.c = p2;
            t1.tu = v3(u[0], u[1], u[2] + (tp2 - 0.0f) * (u[1] - u[2]));
            t1.tv = v3(v[0], v[1], v[2] + (tp2 - 0.0f) * (v[1] - v[2]));
            t2.a = i1; t2.b = p2; t2.c = p1;
            t2.tu = v3(u[0], u[2] + (tp2 - 0.0f) * (u[1] - u[2]), u[2] + (tp1 - 0.0f) * (u[0] - u[2]));
            t2.tv = v3(v[0], v[2] + (tp2 - 0.0f) * (v[1] - v[2]), v[2] + (tp1 - 0.0f) * (v[0] - v[2]));
            if (added < kClipMax) out[added++] = t1;                                // (the reference's arrays hold 12 and are not checked)
            if (added < kClipMax) out[added++] = t2;
        }
    }
    return added;
}

// Renderer::clip_triangle -- renderer.cpp:837-853: right, left, top, bottom, far, near, ping-ponging between the two arrays.
int clip_triangle(bool enable_clipping, Tri4* to_clip, Tri4* clipped)
{
    int n = 1;
    if (enable_clipping) {
        n = clip_to_plane(0, 1, to_clip, n, to_clip);
        n = clip_to_plane(0, -1, to_clip, n, clipped);
        n = clip_to_plane(1, 1, clipped, n, to_clip);
        n = clip_to_plane(1, -1, to_clip, n, clipped);
        n = clip_to_plane(2, 1, clipped, n, to_clip);
        n = clip_to_plane(2, -1, to_clip, n, clipped);
    } else
        clipped[0] = to_clip[0];
    return n;
}

// Renderer::matrix_transform_z -- renderer.cpp:855-867
inline float matrix_transform_z(const M4& m, V3 p)
{
    float zt = m.m[2][0] * p.x + m.m[2][1] * p.y + m.m[2][2] * p.z + m.m[2][3];
    float wt = m.m[3][0] * p.x + m.m[3][1] * p.y + m.m[3][2] * p.z + m.m[3][3];
    if (wt == 1.0f) return zt;
    return zt / wt;
}

// Triangle::edge_function -- triangle.h:65-68
inline float edge_function(V3 p, V3 a, V3 b) { return (b.x - a.x) * (p.y - a.y) - (b.y - a.y) * (p.x - a.x); }

// Transform::operator()(const Triangle&) -- mat.cpp:133-140: the three points, then Triangle(a, b, c, ...) recomputes the normal
inline Tri xform_tri(const M4& m, const Tri& t)
{
    Tri r;
    r.a = xform_point(m, t.a); r.b = xform_point(m, t.b); r.c = xform_point(m, t.c);
    r.normal = cross(sub(r.b, r.a), sub(r.c, r.a));
    r.mat = t.mat; r.tu = t.tu; r.tv = t.tv;
    return r;
}

// argb_super: the supersampled image as clear_image() left it; zbuf: +inf; nbuf (may be null): the SSAO normal buffer.
void raster_trace(const Renderer& r, uint32_t* argb_super, float* zbuf, V3* nbuf, RenderCounters* rc)
{
    int rw, rh;
    r.super_dims(rw, rh);
    const M4 proj = perspective(r.fov, (float)rw / rh, 0.1f, 1000.0f);              // scene/camera.cpp:5-11
    const M4 proj_inv = inverse(proj);
    const M4 world_to_cam = inverse(r.cam_to_world);                                // renderer.cpp:229
    const float height_scaling = 1.0f / rh * 2;
    const float width_scaling = 1.0f / rw * 2;
    Tri4 to_clip[kClipMax], clipped[kClipMax];
    for (size_t ti = 0; ti < r.tris.size(); ti++) {
        const Tri& orig = r.tris[ti];
        const Tri cam = xform_tri(world_to_cam, orig);
        to_clip[0].a = xform_v4(proj, v4(cam.a.x, cam.a.y, cam.a.z, 1));
        to_clip[0].b = xform_v4(proj, v4(cam.b.x, cam.b.y, cam.b.z, 1));
        to_clip[0].c = xform_v4(proj, v4(cam.c.x, cam.c.y, cam.c.z, 1));
        to_clip[0].tu = cam.tu; to_clip[0].tv = cam.tv;
        const int nb = clip_triangle(r.s.enable_clipping != 0, to_clip, clipped);
        for (int ci = 0; ci < nb; ci++) {
            const Tri4& c4 = clipped[ci];
            Tri ndc;                                                                // Triangle(const Triangle4&, ...), triangle.cpp:12-23
            {
                float iaw = 1.0f / c4.a.w, ibw = 1.0f / c4.b.w, icw = 1.0f / c4.c.w;
                ndc.a = v3(c4.a.x * iaw, c4.a.y * iaw, c4.a.z * iaw);
                ndc.b = v3(c4.b.x * ibw, c4.b.y * ibw, c4.b.z * ibw);
                ndc.c = v3(c4.c.x * icw, c4.c.y * icw, c4.c.z * icw);
                ndc.normal = v3(0, 0, 0);                                           // left uninitialised by the reference, never read
                ndc.mat = orig.mat; ndc.tu = c4.tu; ndc.tv = c4.tv;
            }
            const Tri cam_space = xform_tri(proj_inv, ndc);
            const V3 a = ndc.a, b = ndc.b, c = ndc.c;
            float inv_area = 1 / ((b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x));
            float bminx = std::min(a.x, std::min(b.x, c.x)), bminy = std::min(a.y, std::min(b.y, c.y));
            float bmaxx = std::max(a.x, std::max(b.x, c.x)), bmaxy = std::max(a.y, std::max(b.y, c.y));
            int min_x = (int)((bminx + 1) * 0.5 * rw), min_y = (int)((bminy + 1) * 0.5 * rh);
            int max_x = (int)((bmaxx + 1) * 0.5 * rw), max_y = (int)((bmaxy + 1) * 0.5 * rh);
            min_x = std::max(min_x, 0); min_y = std::max(min_y, 0);
            max_x = std::min(rw - 1, max_x); max_y = std::min(rh - 1, max_y);
            float image_y = min_y * height_scaling - 1;
            for (int py = min_y; py <= max_y; py++, image_y += height_scaling) {
                float image_x = min_x * width_scaling - 1;
                for (int px = min_x; px <= max_x; px++, image_x += width_scaling) {
                    V3 pp = v3(image_x + width_scaling * 0.5f, image_y + height_scaling * 0.5f, -1);
                    float u = edge_function(pp, c, a);
                    if (u < 0) continue;
                    float v = edge_function(pp, a, b);
                    if (v < 0) continue;
                    float w = edge_function(pp, b, c);
                    if (w < 0) continue;
                    u *= inv_area; v *= inv_area; w *= inv_area;
                    float za = matrix_transform_z(r.cam_to_world, xform_point(proj_inv, a));
                    float zb = matrix_transform_z(r.cam_to_world, xform_point(proj_inv, b));
                    float zc = matrix_transform_z(r.cam_to_world, xform_point(proj_inv, c));
                    float z_tri = -1 / (1 / za * w + 1 / zb * u + 1 / zc * v);
                    const size_t pix = (size_t)py * rw + px;
                    if (!(z_tri < zbuf[pix])) continue;
                    zbuf[pix] = z_tri;
                    if (nbuf) nbuf[pix] = orig.normal;
                    Col color = col(0.0f);
                    if (r.s.shading_method == RT_SHADING) {
                        RayT ray{r.cam_pos, normalize(sub(xform_point(r.cam_to_world, xform_point(proj_inv, pp)), r.cam_pos))};
                        const Tri world = xform_tri(r.cam_to_world, cam_space);
                        Hit hit;                                                    // trace_triangle, renderer.cpp:619-628
                        if (rc) rc->primary_rays++;
                        if (tri_intersect(world, (int)ti, ray, hit)) {
                            hit.frag = &world;
                            XorShift rng;
                            rng.state = pixel_seed((uint32_t)(py * rw + px), r.s.rng_seed);
                            if (rc) rc->primary_hits++;
                            color = r.shade(ray, hit, 0, rng, rc, false);
                        }
                    } else if (r.s.shading_method == RT_ABS_NORMALS_SHADING) {
                        V3 n = normalize(orig.normal);
                        color = Col{std::abs(n.x), std::abs(n.y), std::abs(n.z)};
                    } else if (r.s.shading_method == RT_PASTEL_NORMALS_SHADING) {
                        V3 n = normalize(orig.normal);
                        color = cscale(cadd(Col{n.x, n.y, n.z}, col(1.0f)), 0.5f);
                    } else if (r.s.shading_method == RT_BARYCENTRIC_COORDINATES_SHADING) {
                        color = cadd(cadd(cscale(Col{1, 0, 0}, u), cscale(Col{0, 1, 0}, v)), cscale(Col{0, 0, 1}, 1 - u - v));
                    } else if (r.s.shading_method == RT_VISUALIZE_AO) {             // shade_visualize_ao(proj_inv(ndc), u, v), :419-434
                        color = Col{0.9f, 0.9f, 0.9f};
                        if (r.s.enable_ao_mapping) {
                            float tu, tv;
                            tri_texcoords(cam_space, u, v, tu, tv);
                            color = cmul(color, col(r.tex[RT_TEX_AO].texture_floor(tu, tv).r));
                        }
                    }
                    argb_super[pix] = quantise(color);
                }
            }
        }
    }
}

// clear_z_buffer / clear_normal_buffer / clear_image (QT/mainwindow.cpp:184-190), raster_trace(), post_process()
void render_raster(const Renderer& r, uint32_t* argb_out, int rng_mode, const uint32_t* ref_seeds9, RenderCounters* rc)
{
    int rw, rh;
    r.super_dims(rw, rh);
    const uint32_t background = quantise(Col{135.0f / 255.0f, 206.0f / 255.0f, 235.0f / 255.0f});   // renderer.cpp:19,175-180
    std::vector<uint32_t> super((size_t)rw * rh, background);
    std::vector<float> z((size_t)rw * rh, INFINITY);
    std::vector<V3> n;
    if (r.s.enable_ssao) n.assign((size_t)rw * rh, v3(0, 0, 0));
    raster_trace(r, super.data(), z.data(), r.s.enable_ssao ? n.data() : nullptr, rc);
    if (r.s.enable_ssao) {
        SsaoFrame f;
        f.rw = rw; f.rh = rh; f.z = z.data(); f.n = n.data();
        f.aspect = (float)rw / rh;
        f.proj = perspective(r.fov, f.aspect, 0.1f, 1000.0f);
        f.fov_mult_simd = (float)std::tan(r.fov / 2 / 180 * M_PI);
        f.fov_mult_scalar = std::tan(((float)M_PI / 180) * (r.fov / 2));
        f.samples = r.s.ssao_sample_count; f.radius = r.s.ssao_radius;
        ssao_post_process(f, super.data(), r.s.rng_seed, r.s.ssao_amount, rng_mode, ref_seeds9);
    }
    if (r.s.enable_ssaa) downscale(super.data(), rw, rh, r.s.ssaa_factor, argb_out);
    else memcpy(argb_out, super.data(), super.size() * sizeof(uint32_t));
}

struct BvhHandle {
    std::vector<Tri> tris;
    Octree tree;
};

void default_settings(RtSettings* s);

} // namespace

extern "C" {

void* orc_bvh_create(const float* xyz9, size_t n, int max_depth, int leaf_max, double* build_ms)
{
    BvhHandle* h = new BvhHandle();
    h->tris = make_tris(xyz9, nullptr, nullptr, n);
    auto t0 = std::chrono::steady_clock::now();
    h->tree.build(&h->tris, max_depth, leaf_max);
    auto t1 = std::chrono::steady_clock::now();
    if (build_ms) *build_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    return h;
}

void orc_bvh_destroy(void* h) { delete (BvhHandle*)h; }

void orc_bvh_stats(void* handle, uint64_t* out)
{
    for (int i = 0; i < 6; i++) out[i] = 0;
    tree_stats(((BvhHandle*)handle)->tree.root, 0, out);
}

double orc_bvh_intersect(void* handle, const float* o3, const float* d3, size_t n, int32_t* tri_id, float* t, float* u,
                         float* v, int threads)
{
    BvhHandle* h = (BvhHandle*)handle;
    if (threads <= 0) threads = omp_get_max_threads();
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
    for (long long i = 0; i < (long long)n; i++) {
        RayT ray{v3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), v3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2])};
        Hit hit;
        bool found = h->tree.intersect(ray, hit);
        if (tri_id) tri_id[i] = found ? hit.tri : -1;
        if (t) t[i] = found ? hit.t : -1.0f;
        if (u) u[i] = found ? hit.u : 0.0f;
        if (v) v[i] = found ? hit.v : 0.0f;
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

double orc_bvh_count(void* handle, const float* o3, const float* d3, size_t n, uint64_t* out2, int threads)
{
    BvhHandle* h = (BvhHandle*)handle;
    if (threads <= 0) threads = omp_get_max_threads();
    uint64_t vt = 0, tt = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads) reduction(+ : vt, tt)
    for (long long i = 0; i < (long long)n; i++) {
        RayT ray{v3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), v3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2])};
        Hit hit;
        Counters c;
        h->tree.intersect(ray, hit, &c);
        vt += c.volume_tests;
        tt += c.triangle_tests;
    }
    auto t1 = std::chrono::steady_clock::now();
    out2[0] = vt;
    out2[1] = tt;
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

int orc_triangle_intersect(const float* xyz9, const float* o3, const float* d3, float* t, float* u, float* v)
{
    std::vector<Tri> tr = make_tris(xyz9, nullptr, nullptr, 1);
    RayT ray{v3(o3[0], o3[1], o3[2]), v3(d3[0], d3[1], d3[2])};
    Hit hit;
    bool r = tri_intersect(tr[0], 0, ray, hit);
    if (t) *t = hit.t;
    if (u) *u = hit.u;
    if (v) *v = hit.v;
    return r ? 1 : 0;
}

void* orc_renderer_create(void)
{
    Renderer* r = new Renderer();
    default_settings(&r->s);
    return r;
}

void orc_renderer_destroy(void* h) { delete (Renderer*)h; }

void orc_renderer_configure(void* h, const RtSettings* s, float fov)
{
    Renderer* r = (Renderer*)h;
    r->s = *s;
    r->fov = fov;
}

void orc_renderer_set_triangles(void* h, const float* xyz9, const float* uv6, const int32_t* mat, size_t n)
{
    Renderer* r = (Renderer*)h;                                                     // renderer.cpp:137-144
    r->tris = make_tris(xyz9, uv6, mat, n);
    delete r->bvh;
    r->bvh = new Octree();
    r->bvh->build(&r->tris, r->s.bvh_max_depth, r->s.bvh_leaf_object_count);
}

void orc_renderer_set_materials(void* h, const RtMaterial* mats, size_t n)
{
    Renderer* r = (Renderer*)h;
    r->mats.assign(mats, mats + n);
}

void orc_renderer_set_texture_f32(void* h, int slot, const float* rgba, int w, int hgt)
{
    Renderer* r = (Renderer*)h;
    if (slot < 0 || slot >= RT_TEX_COUNT) return;
    r->tex[slot].w = w;
    r->tex[slot].h = hgt;
    r->tex[slot].px.assign(rgba, rgba + (size_t)w * hgt * 4);
}

void orc_renderer_set_texture_u8(void* h, int slot, const uint8_t* rgba, int w, int hgt)
{
    std::vector<float> f((size_t)w * hgt * 4);                                      // image_io.cpp:115-121: Color(u8) / 255
    float kk = 1 / 255.0f;                                                          // color.cpp:87-91: kk = 1 / k; kk * c
    for (size_t i = 0; i < f.size(); i++) f[i] = (float)rgba[i] * kk;
    orc_renderer_set_texture_f32(h, slot, f.data(), w, hgt);
}

void orc_renderer_set_camera_transform(void* h, const float m[16])
{
    Renderer* r = (Renderer*)h;                                                     // renderer.cpp:226-233
    r->cam_to_world = from16(m);
    r->cam_pos = xform_point(r->cam_to_world, v3(0, 0, 0));
}

void orc_renderer_set_light(void* h, const float p[3]) { ((Renderer*)h)->light = v3(p[0], p[1], p[2]); }

// Renderer::add_analytic_shape(Sphere(center, radius, mat)) / (Plane(point, normal, mat)) -- renderer.cpp:146
void orc_renderer_add_sphere(void* h, const float c[3], float radius, int mat)
{
    ((Renderer*)h)->shapes.push_back(Renderer::Shape{0, v3(c[0], c[1], c[2]), v3(0, 0, 0), radius * radius, mat});
}
void orc_renderer_add_plane(void* h, const float p[3], const float n[3], int mat)
{
    ((Renderer*)h)->shapes.push_back(Renderer::Shape{1, v3(p[0], p[1], p[2]), v3(n[0], n[1], n[2]), 0.0f, mat});
}

void orc_camera_matrices(float fov, float aspect, float znear, float zfar, float* proj16, float* proj_inv16)
{
    M4 p = perspective(fov, aspect, znear, zfar);
    M4 pi = inverse(p);
    memcpy(proj16, p.m, sizeof(p.m));
    memcpy(proj_inv16, pi.m, sizeof(pi.m));
}

void orc_transform_inverse(const float m[16], float* out16)
{
    M4 inv = inverse(from16(m));
    memcpy(out16, inv.m, sizeof(inv.m));
}

double orc_renderer_trace_rows(void* h, const float cam_to_world[16], uint32_t* argb_super, int row_begin, int row_end,
                               int row_step, int reseed, int threads)
{
    (void)reseed; // the restatement always uses the per-pixel stream
    Renderer* r = (Renderer*)h;
    M4 c2w = from16(cam_to_world);
    r->cam_to_world = c2w;
    r->cam_pos = xform_point(c2w, v3(0, 0, 0));
    return trace_rows(*r, c2w, argb_super, row_begin, row_end, row_step, threads, nullptr);
}

void orc_renderer_count_rows(void* h, const float cam_to_world[16], int row_begin, int row_end, int row_step,
                             uint64_t* out, int threads)
{
    Renderer* r = (Renderer*)h;
    M4 c2w = from16(cam_to_world);
    r->cam_to_world = c2w;
    r->cam_pos = xform_point(c2w, v3(0, 0, 0));
    RenderCounters rc;
    trace_rows(*r, c2w, nullptr, row_begin, row_end, row_step, threads, &rc);
    out[0] = rc.primary_rays; out[1] = rc.shadow_rays; out[2] = rc.reflection_rays; out[3] = rc.reflection_shadow_rays;
    out[4] = rc.primary_hits;
    out[5] = rc.primary.volume_tests; out[6] = rc.primary.triangle_tests;
    out[7] = rc.shadow.volume_tests; out[8] = rc.shadow.triangle_tests;
    out[9] = rc.reflection.volume_tests; out[10] = rc.reflection.triangle_tests;
    for (int i = 11; i < 16; i++) out[i] = 0;
}

// Renderer::ray_trace + post_process(SSAA) -- renderer.cpp:1068-1135
double orc_renderer_render(void* h, uint32_t* argb_out, int threads)
{
    Renderer* r = (Renderer*)h;
    int rw, rh;
    r->super_dims(rw, rh);
    std::vector<uint32_t> super((size_t)rw * rh);
    auto t0 = std::chrono::steady_clock::now();
    trace_rows(*r, r->cam_to_world, super.data(), 0, rh, 1, threads, nullptr);
    if (r->s.enable_ssaa) downscale(super.data(), rw, rh, r->s.ssaa_factor, argb_out);
    else memcpy(argb_out, super.data(), super.size() * sizeof(uint32_t));
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// The same with enable_ssao: rng_mode 0 = per-pixel stream (rng_seed), 1 = reference order with the 9 given generator seeds
void orc_renderer_render_ssao(void* h, uint32_t* argb_out, int threads, int rng_mode, const uint32_t* ref_seeds9)
{
    render_with_ssao(*(Renderer*)h, argb_out, threads, rng_mode, ref_seeds9);
}

// Renderer::raster_trace + post_process (hybrid_rasterization_tracing) -- renderer.cpp:869-1006,1118-1124, in sequential
// triangle order.  out5 (may be NULL): fragments shaded with RT_SHADING, of which hit their triangle, shadow rays,
// reflection rays, reflection shadow rays.
void orc_renderer_raster(void* h, uint32_t* argb_out, int rng_mode, const uint32_t* ref_seeds9, uint64_t* out5)
{
    RenderCounters rc;
    render_raster(*(Renderer*)h, argb_out, rng_mode, ref_seeds9, out5 ? &rc : nullptr);
    if (out5) { out5[0] = rc.primary_rays; out5[1] = rc.primary_hits; out5[2] = rc.shadow_rays; out5[3] = rc.reflection_rays; out5[4] = rc.reflection_shadow_rays; }
}

void orc_downscale(const uint32_t* in, int w, int hgt, int factor, uint32_t* out) { downscale(in, w, hgt, factor, out); }

int orc_omp_max_threads(void) { return omp_get_max_threads(); }

} // extern "C"

// The oracle library carries its own copy of the settings defaults so that it never links the product.
namespace {
void default_settings(RtSettings* s)
{
    memset(s, 0, sizeof(*s));
    s->image_width = 1024; s->image_height = 1024; s->ssaa_factor = 2; s->max_recursion_depth = 5;
    s->enable_bvh = 1; s->bvh_max_depth = 12; s->bvh_leaf_object_count = 40;
    s->enable_ambient = s->enable_diffuse = s->enable_specular = s->enable_emissive = 1;
    s->rough_reflections_sample_count = 3;
    s->enable_clipping = 1;                                                          // rendererSettings.h:40
}
} // namespace
