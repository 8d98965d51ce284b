// octree_build.cpp -- host-side construction of the flattened octree (see scene_layout.h).
//
// The reference builds its octree by inserting triangles one at a time (BVH::build_bvh, bvh.cpp:58-66 ->
// OctreeNode::insert, bvh.h:169-193).  The resulting tree does not depend on insertion order: a cell is split iff
// more than `leaf_max` triangles are ever routed to it and it is shallower than `max_depth`, and a split
// re-distributes everything by comparing the triangle's bbox centroid with the cell midpoint (bvh.h:195-210).
// Only the ORDER of the triangles inside a leaf follows the input array.  So the same tree is produced here
// top-down with a stable 8-way counting partition of an index array (parallel for large cells), which is what
// lets a 10M-triangle scene build in about a second instead of the reference's pointer-chasing insertion.
//
// Faithfully kept quirks: child cells 2, 4 and 6 get the lower corner `lo + (0,my,0)` etc. (bvh.h:161,163,165),
// all 8 children exist after a split (empty ones are counted, then dropped from the flat form), slab extents are
// the min/max over the 7 plane normals of bvh.cpp:8-16 of the vertex dot products (bvh.h:33-46).
#include "scene_layout.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>

namespace rtb {
namespace {

constexpr int PLANES = 7;

struct Planes {
    V3 n[PLANES];
    Planes()
    {
        float s = std::sqrt(3.0f) / 3; // bvh.cpp:12-15
        n[0] = v3(1, 0, 0);
        n[1] = v3(0, 1, 0);
        n[2] = v3(0, 0, 1);
        n[3] = v3(s, s, s);
        n[4] = v3(-s, s, s);
        n[5] = v3(-s, -s, s);
        n[6] = v3(s, -s, s);
    }
};
const Planes kPlanes;

struct Cell {
    int32_t child[8];     // index into cells, -1 = empty cell
    uint32_t begin, count; // leaf: range in the permutation array
    bool leaf;
    int32_t job;          // >= 0: the subtree below this cell is built and flattened by that job (see build_flat_scene)
    float nr[PLANES], fr[PLANES];
};

// A subtree handed to one thread: it builds the subtree's cells and flattens them into buffers of its own.  Local record
// 0 is the subtree's root record, 1 is padding (so that local and final record indices have the same parity), the
// children blocks start at 2.
struct Job {
    V3 lo, hi;
    int depth;
    uint32_t b, e;
    FlatScene flat;
    float nr[PLANES], fr[PLANES];
    size_t rec_base = 0, tri_base = 0;     // where the job's records [2..) and triangles go in the final arrays
};

struct Builder {
    const float* xyz9;
    size_t n;
    int max_depth, leaf_max;
    int leaf_split = 0;          // > 0: reference leaves with more triangles than this are refined (emit_group)
    // shared by the top-level builder and the jobs; a job only touches its own range [b, e) of perm / scratch / oct
    const V3* centroid = nullptr;
    uint32_t* perm = nullptr;
    uint32_t* scratch = nullptr;
    uint8_t* oct = nullptr;
    std::vector<Cell> cells;
    FlatScene* out;
    // top-level builder only: subtrees of at most job_size triangles become jobs; rec slots waiting for a job's root record
    uint32_t job_size = 0;
    bool in_parallel = false;    // a job runs inside the parallel loop: its partitions are single-threaded
    std::vector<Job>* jobs = nullptr;
    std::vector<std::pair<uint32_t, int32_t>> job_slots;

    V3 vert(uint32_t tri, int k) const
    {
        const float* p = xyz9 + 9 * (size_t)tri + 3 * k;
        return v3(p[0], p[1], p[2]);
    }

    // Stable partition of perm[b, e) into 8 octants by centroid > midpoint; returns bucket offsets.
    void partition(uint32_t b, uint32_t e, V3 mid, uint32_t offs[9])
    {
        const uint32_t len = e - b;
        const int T = (!in_parallel && len > (1u << 16)) ? omp_get_max_threads() : 1;
        std::vector<uint32_t> counts((size_t)T * 8, 0);
        const uint32_t chunk = (len + T - 1) / T;
#pragma omp parallel num_threads(T) if (T > 1)
        {
            int t = T > 1 ? omp_get_thread_num() : 0;
            uint32_t cb = b + (uint32_t)t * chunk, ce = std::min(e, cb + chunk);
            uint32_t* cnt = &counts[(size_t)t * 8];
            for (uint32_t i = cb; i < ce; i++) {
                V3 c = centroid[perm[i]];
                int o = (c.x > mid.x ? 1 : 0) + (c.y > mid.y ? 2 : 0) + (c.z > mid.z ? 4 : 0); // bvh.h:203-207
                oct[i] = (uint8_t)o;
                cnt[o]++;
            }
        }
        uint32_t total[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int t = 0; t < T; t++)
            for (int o = 0; o < 8; o++) total[o] += counts[(size_t)t * 8 + o];
        offs[0] = b;
        for (int o = 0; o < 8; o++) offs[o + 1] = offs[o] + total[o];
        // start of (bucket o, chunk t) = offs[o] + sum of earlier chunks' counts: keeps the input order
        std::vector<uint32_t> start((size_t)T * 8);
        for (int o = 0; o < 8; o++) {
            uint32_t acc = offs[o];
            for (int t = 0; t < T; t++) {
                start[(size_t)t * 8 + o] = acc;
                acc += counts[(size_t)t * 8 + o];
            }
        }
#pragma omp parallel num_threads(T) if (T > 1)
        {
            int t = T > 1 ? omp_get_thread_num() : 0;
            uint32_t cb = b + (uint32_t)t * chunk, ce = std::min(e, cb + chunk);
            uint32_t* st = &start[(size_t)t * 8];
            for (uint32_t i = cb; i < ce; i++) scratch[st[oct[i]]++] = perm[i];
        }
        if (T > 1) {
#pragma omp parallel for num_threads(T)
            for (long long i = b; i < (long long)e; i++) perm[i] = scratch[i];
        } else
            memcpy(&perm[b], &scratch[b], sizeof(uint32_t) * len);
    }

    int32_t build(V3 lo, V3 hi, int depth, uint32_t b, uint32_t e)
    {
        int32_t self = (int32_t)cells.size();
        cells.emplace_back();
        {
            Cell& c = cells[self];
            for (int i = 0; i < 8; i++) c.child[i] = -1;
            for (int i = 0; i < PLANES; i++) { c.nr[i] = INFINITY; c.fr[i] = -INFINITY; }
            c.begin = b;
            c.count = e - b;
            c.leaf = true;
            c.job = -1;
        }
        if (jobs && e - b <= job_size) {                 // small enough: another thread builds and flattens this subtree
            cells[self].job = (int32_t)jobs->size();
            jobs->emplace_back();
            Job& j = jobs->back();
            j.lo = lo; j.hi = hi; j.depth = depth; j.b = b; j.e = e;
            return self;
        }
        out->nodes++;
        const bool split = (size_t)(e - b) > (size_t)leaf_max && depth != max_depth; // bvh.h:171-177
        if (!split) {
            out->leaves++;
            if (e == b) out->empty_leaves++;
            out->max_depth_reached = std::max(out->max_depth_reached, (uint32_t)depth);
            out->max_leaf_size = std::max(out->max_leaf_size, e - b);
            Cell& c = cells[self];
            for (uint32_t i = b; i < e; i++) {                                      // bvh.h:33-46,143-145
                uint32_t t = perm[i];
                for (int k = 0; k < 3; k++) {
                    V3 p = vert(t, k);
                    for (int pl = 0; pl < PLANES; pl++) {
                        float d = dot(kPlanes.n[pl], p);
                        c.nr[pl] = std::min(c.nr[pl], d);
                        c.fr[pl] = std::max(c.fr[pl], d);
                    }
                }
            }
            return self;
        }
        out->interior++;
        // bvh.h:155-157 (division by the int 2 is the exact float halving)
        V3 mid = v3((lo.x + hi.x) / 2, (lo.y + hi.y) / 2, (lo.z + hi.z) / 2);
        uint32_t offs[9];
        partition(b, e, mid, offs);
        // OctreeNode::create_children, bvh.h:159-166 -- children 2, 4, 6 use lo + (..) as written there.
        V3 clo[8], chi[8];
        clo[0] = lo;                                   chi[0] = v3(mid.x, mid.y, mid.z);
        clo[1] = v3(mid.x, lo.y, lo.z);                chi[1] = v3(hi.x, mid.y, mid.z);
        clo[2] = lo + v3(0, mid.y, 0);                 chi[2] = v3(mid.x, hi.y, mid.z);
        clo[3] = v3(mid.x, mid.y, lo.z);               chi[3] = v3(hi.x, hi.y, mid.z);
        clo[4] = lo + v3(0, 0, mid.z);                 chi[4] = v3(mid.x, mid.y, hi.z);
        clo[5] = v3(mid.x, lo.y, mid.z);               chi[5] = v3(hi.x, mid.y, hi.z);
        clo[6] = lo + v3(0, mid.y, mid.z);             chi[6] = v3(mid.x, hi.y, hi.z);
        clo[7] = v3(mid.x, mid.y, mid.z);              chi[7] = v3(hi.x, hi.y, hi.z);
        for (int o = 0; o < 8; o++) {
            if (offs[o + 1] == offs[o]) {               // an empty child cell still exists in the reference tree
                out->nodes++;
                out->leaves++;
                out->empty_leaves++;
                out->max_depth_reached = std::max(out->max_depth_reached, (uint32_t)(depth + 1));
                continue;
            }
            int32_t ci = build(clo[o], chi[o], depth + 1, offs[o], offs[o + 1]);
            Cell& c = cells[self];
            c.child[o] = ci;
            for (int pl = 0; pl < PLANES; pl++) {                                    // bvh.h:147-149
                c.nr[pl] = std::min(c.nr[pl], cells[ci].nr[pl]);
                c.fr[pl] = std::max(c.fr[pl], cells[ci].fr[pl]);
            }
        }
        cells[self].leaf = false;
        return self;
    }

    // f moved by `ulps` representable floats towards -inf / +inf: nextafterf applied `ulps` times, done on the bits
    // (floats ordered as sign-magnitude integers); 56 calls per record made the libm loop the slowest part of flattening.
    static float pad(float f, int ulps)
    {
        if (f != f) return f;
        int32_t i;
        memcpy(&i, &f, 4);
        int64_t ord = i >= 0 ? (int64_t)i : -(int64_t)(i & 0x7fffffff);     // one zero, as nextafterf walks: -min, 0, +min
        ord += ulps;
        const int64_t inf = 0x7f800000;
        if (ord >= inf) return INFINITY;
        if (ord <= -inf) return -INFINITY;
        uint32_t o = ord > 0 ? (uint32_t)ord : ord < 0 ? (0x80000000u | (uint32_t)(-ord)) : ((uint32_t)i & 0x80000000u);   // zero keeps the sign it came from
        float r;
        memcpy(&r, &o, 4);
        return r;
    }
    static float pad_down(float f) { return pad(f, -RT_SLAB_PAD_ULPS); }
    static float pad_up(float f) { return pad(f, RT_SLAB_PAD_ULPS); }

    uint32_t alloc_block(uint32_t k)
    {
        size_t recs = out->recs.size() / 4;
        if (recs & 1) recs++; // 128-byte alignment of the block
        out->recs.resize((recs + k) * 4, F4{0, 0, 0, 0});
        return (uint32_t)recs;
    }

    void write_record(uint32_t rec, const float* nr_in, const float* fr_in, uint32_t link, uint32_t meta)
    {
        float nr[PLANES], fr[PLANES];
        for (int i = 0; i < PLANES; i++) { nr[i] = pad_down(nr_in[i]); fr[i] = pad_up(fr_in[i]); }
        F4* q = &out->recs[(size_t)rec * 4];                                     // (near, far) pairs, slab by slab: scene_layout.h
        q[0] = F4{nr[0], fr[0], nr[1], fr[1]};
        q[1] = F4{nr[2], fr[2], nr[3], fr[3]};
        q[2] = F4{nr[4], fr[4], nr[5], fr[5]};
        q[3] = F4{nr[6], fr[6], u2f(link), u2f(meta)};
    }

    // Appends triangles `ids` (sorted ascending = the reference's order inside a leaf) as one leaf; returns link.
    uint32_t append_leaf(const uint32_t* ids, uint32_t count, const float* uv6, const int32_t* mat)
    {
        const uint32_t link = (uint32_t)(out->tris.size() / 3);
        for (uint32_t i = 0; i < count; i++) {
            uint32_t t = ids[i];
            V3 a = vert(t, 0), b = vert(t, 1), cc = vert(t, 2);
            V3 nrm = cross(b - a, cc - a);                                           // triangle.cpp:9-10
            out->tris.push_back(F4{a.x, a.y, a.z, nrm.x});
            out->tris.push_back(F4{b.x, b.y, b.z, nrm.y});
            out->tris.push_back(F4{cc.x, cc.y, cc.z, nrm.z});
            float u[6] = {-1, -1, -1, -1, -1, -1};                                   // triangle.h:48
            if (uv6) memcpy(u, uv6 + 6 * (size_t)t, sizeof(u));
            int32_t m = mat ? mat[t] : -1;
            out->shade.push_back(F4{u[0], u[1], u[2], u[3]});
            out->shade.push_back(F4{u[4], u[5], u2f((uint32_t)m), u2f(t)});
            out->orig.push_back((int32_t)t);
        }
        return link;
    }

    void group_volume(const uint32_t* ids, uint32_t count, float* nr, float* fr) const
    {
        for (int i = 0; i < PLANES; i++) { nr[i] = INFINITY; fr[i] = -INFINITY; }
        for (uint32_t i = 0; i < count; i++)
            for (int k = 0; k < 3; k++) {
                V3 p = vert(ids[i], k);
                for (int pl = 0; pl < PLANES; pl++) {
                    float d = dot(kPlanes.n[pl], p);
                    nr[pl] = std::min(nr[pl], d);
                    fr[pl] = std::max(fr[pl], d);
                }
            }
    }

    // B200-side refinement of an oversized reference leaf (not part of the reference tree): the triangles of the
    // leaf are split into up to 8 groups by three rounds of median cuts of their bbox centroids (each cut along the
    // axis of largest centroid extent of the group being cut), recursively, until a group holds <= leaf_split
    // triangles.  Every group gets its own 7-slab volume.  Closest-hit results cannot change: the groups partition the
    // leaf, volumes bound their triangles, and ties on t are resolved by original index in the kernel
    // (the reference keeps the first triangle of the leaf in array order, bvh.h:241).
    void emit_group(uint32_t* ids, uint32_t count, uint32_t rec, const float* nr, const float* fr, const float* uv6, const int32_t* mat)
    {
        if (count <= (uint32_t)leaf_split) {
            std::sort(ids, ids + count);
            write_record(rec, nr, fr, append_leaf(ids, count, uv6, mat), RT_META_LEAF | count);
            return;
        }
        struct Range { uint32_t b, e; };
        Range groups[8];
        int ng = 1;
        groups[0] = Range{0, count};
        for (int round = 0; round < 3; round++) {
            Range next[8];
            int nn = 0;
            for (int g = 0; g < ng; g++) {
                Range r = groups[g];
                if (r.e - r.b <= (uint32_t)leaf_split) { next[nn++] = r; continue; }
                V3 lo = v3(INFINITY, INFINITY, INFINITY), hi = v3(-INFINITY, -INFINITY, -INFINITY);
                for (uint32_t i = r.b; i < r.e; i++) {
                    V3 c = centroid[ids[i]];
                    lo = v3(std::min(lo.x, c.x), std::min(lo.y, c.y), std::min(lo.z, c.z));
                    hi = v3(std::max(hi.x, c.x), std::max(hi.y, c.y), std::max(hi.z, c.z));
                }
                V3 ext = hi - lo;
                int axis = (ext.x >= ext.y && ext.x >= ext.z) ? 0 : (ext.y >= ext.z ? 1 : 2);
                uint32_t mid = r.b + (r.e - r.b) / 2;
                std::nth_element(ids + r.b, ids + mid, ids + r.e, [&](uint32_t a, uint32_t b) {
                    float ca = (&centroid[a].x)[axis], cb = (&centroid[b].x)[axis];
                    return ca < cb || (ca == cb && a < b);
                });
                next[nn++] = Range{r.b, mid};
                next[nn++] = Range{mid, r.e};
            }
            ng = nn;
            for (int g = 0; g < ng; g++) groups[g] = next[g];
        }
        const uint32_t first = alloc_block((uint32_t)ng);
        write_record(rec, nr, fr, first, (uint32_t)ng);
        for (int g = 0; g < ng; g++) {
            float gn[PLANES], gf[PLANES];
            group_volume(ids + groups[g].b, groups[g].e - groups[g].b, gn, gf);
            emit_group(ids + groups[g].b, groups[g].e - groups[g].b, first + (uint32_t)g, gn, gf, uv6, mat);
        }
    }

    void emit(int32_t cell, uint32_t rec, const float* uv6, const int32_t* mat)
    {
        const Cell c = cells[cell];
        if (c.job >= 0) {                              // the job's root record goes here once the job's place is known
            job_slots.emplace_back(rec, c.job);
            return;
        }
        if (c.leaf) {
            if (leaf_split > 0 && c.count > (uint32_t)leaf_split) {
                std::vector<uint32_t> ids(perm + c.begin, perm + c.begin + c.count);
                emit_group(ids.data(), c.count, rec, c.nr, c.fr, uv6, mat);
            } else
                write_record(rec, c.nr, c.fr, append_leaf(&perm[c.begin], c.count, uv6, mat), RT_META_LEAF | c.count);
            return;
        }
        int kids[8], k = 0;
        for (int o = 0; o < 8; o++)
            if (c.child[o] >= 0) kids[k++] = c.child[o];
        if (k == 1 && leaf_split > 0) {       // a cell with one non-empty child has that child's volume: skip the level
            emit(kids[0], rec, uv6, mat);
            return;
        }
        const uint32_t first = alloc_block((uint32_t)k);
        write_record(rec, c.nr, c.fr, first, (uint32_t)k);
        for (int j = 0; j < k; j++) emit(kids[j], first + (uint32_t)j, uv6, mat);
    }

    // Top-level builder, after the jobs have run: the slab extents of the cells above them (bvh.h:147-149).
    void refit(int32_t cell)
    {
        Cell& c = cells[cell];
        if (c.job >= 0) {
            const Job& j = (*jobs)[c.job];
            for (int pl = 0; pl < PLANES; pl++) { c.nr[pl] = j.nr[pl]; c.fr[pl] = j.fr[pl]; }
            return;
        }
        if (c.leaf) return;
        for (int pl = 0; pl < PLANES; pl++) { c.nr[pl] = INFINITY; c.fr[pl] = -INFINITY; }
        for (int o = 0; o < 8; o++) {
            if (c.child[o] < 0) continue;
            refit(c.child[o]);
            const Cell& k = cells[c.child[o]];
            for (int pl = 0; pl < PLANES; pl++) { c.nr[pl] = std::min(c.nr[pl], k.nr[pl]); c.fr[pl] = std::max(c.fr[pl], k.fr[pl]); }
        }
    }
};

} // namespace

// The top of the tree as one small contiguous table (scene_layout.h): breadth-first from the root cell, whole child
// blocks, while they fit RT_TOP_RECORDS records.  The parent record of a block that made it gets RT_META_TOP and the
// block's offset in the table written into its meta word -- in recs[] and, when the parent is itself in the table, in its
// copy there.
static void build_top_table(FlatScene& out)
{
    out.top.clear();
    if (out.recs.size() < 4) return;
    struct Ref { size_t rec; long top_at; };                  // a record of recs[] and where its copy sits in the table (-1: the root, not copied)
    std::vector<Ref> queue;
    queue.push_back(Ref{0, -1});
    for (size_t qi = 0; qi < queue.size(); qi++) {
        const Ref r = queue[qi];
        F4* q3 = &out.recs[r.rec * 4 + 3];
        const uint32_t link = f2u(q3->z), meta = f2u(q3->w);
        if ((meta & RT_META_LEAF) || meta == 0) continue;
        const uint32_t count = meta & RT_META_COUNT_MASK;
        const size_t at = out.top.size() / 4;
        if (at + count > RT_TOP_RECORDS) continue;
        for (uint32_t k = 0; k < count; k++) {
            for (int j = 0; j < 4; j++) out.top.push_back(out.recs[((size_t)link + k) * 4 + j]);
            queue.push_back(Ref{(size_t)link + k, (long)(at + k)});
        }
        const uint32_t tagged = meta | RT_META_TOP | ((uint32_t)at << RT_META_TOP_SHIFT);
        q3->w = u2f(tagged);
        if (r.top_at >= 0) out.top[(size_t)r.top_at * 4 + 3].w = u2f(tagged);
    }
}

void build_flat_scene(const float* xyz9, const float* uv6, const int32_t* mat, size_t n, int max_depth,
                      int leaf_max, int leaf_split, FlatScene& out)
{
    out = FlatScene();
    std::vector<V3> centroid(n);
    std::vector<uint32_t> perm(n), scratch(n);
    std::vector<uint8_t> oct(n);
    Builder b;
    b.xyz9 = xyz9;
    b.n = n;
    b.max_depth = max_depth;
    b.leaf_max = leaf_max;
    b.leaf_split = leaf_split;
    b.out = &out;
    b.centroid = centroid.data();
    b.perm = perm.data();
    b.scratch = scratch.data();
    b.oct = oct.data();
    // root cell = bounds of all vertices (bvh.cpp:25-36); centroids as Triangle::bbox_centroid (triangle.cpp:162-165)
    V3 lo = v3(INFINITY, INFINITY, INFINITY), hi = v3(-INFINITY, -INFINITY, -INFINITY);
#pragma omp parallel
    {
        V3 tlo = lo, thi = hi;
#pragma omp for nowait
        for (long long i = 0; i < (long long)n; i++) {
            V3 a = b.vert((uint32_t)i, 0), bb = b.vert((uint32_t)i, 1), c = b.vert((uint32_t)i, 2);
            V3 mn = v3(std::min(a.x, std::min(bb.x, c.x)), std::min(a.y, std::min(bb.y, c.y)), std::min(a.z, std::min(bb.z, c.z)));
            V3 mx = v3(std::max(a.x, std::max(bb.x, c.x)), std::max(a.y, std::max(bb.y, c.y)), std::max(a.z, std::max(bb.z, c.z)));
            float kk = 1.f / 2;                                                     // Point operator/ (vec.cpp:52-56)
            centroid[i] = kk * (mn + mx);
            perm[i] = (uint32_t)i;
            tlo = v3(std::min(tlo.x, mn.x), std::min(tlo.y, mn.y), std::min(tlo.z, mn.z));
            thi = v3(std::max(thi.x, mx.x), std::max(thi.y, mx.y), std::max(thi.z, mx.z));
        }
#pragma omp critical
        {
            lo = v3(std::min(lo.x, tlo.x), std::min(lo.y, tlo.y), std::min(lo.z, tlo.z));
            hi = v3(std::max(hi.x, thi.x), std::max(hi.y, thi.y), std::max(hi.z, thi.z));
        }
    }
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();

    // Stage A (this thread; the 8-way partitions of big cells are parallel inside): the top of the tree, down to subtrees
    // of at most job_size triangles.  Stage B: those subtrees, one thread each, cells AND flattening, into buffers of
    // their own.  Stage C: the top is flattened, the jobs' buffers are appended behind it and their links relocated.
    // The tree is the same as a sequential build's; only the order of the record blocks in memory depends on job_size.
    std::vector<Job> jobs;
    const int threads = std::max(1, omp_get_max_threads());
    b.jobs = &jobs;
    b.job_size = (uint32_t)std::max<size_t>(4096, n / ((size_t)threads * 8));
    int32_t root = b.build(lo, hi, 0, 0, (uint32_t)n);
    const double t1 = now();

#pragma omp parallel for schedule(dynamic, 1)
    for (long long ji = 0; ji < (long long)jobs.size(); ji++) {
        Job& j = jobs[ji];
        Builder jb;
        jb.xyz9 = xyz9; jb.n = n; jb.max_depth = max_depth; jb.leaf_max = leaf_max; jb.leaf_split = leaf_split;
        jb.centroid = centroid.data(); jb.perm = perm.data(); jb.scratch = scratch.data(); jb.oct = oct.data();
        jb.out = &j.flat;
        jb.in_parallel = true;
        const size_t cnt = j.e - j.b;
        jb.cells.reserve(cnt / 8 + 16);
        j.flat.tris.reserve(cnt * 3); j.flat.shade.reserve(cnt * 2); j.flat.orig.reserve(cnt);
        int32_t jr = jb.build(j.lo, j.hi, j.depth, j.b, j.e);
        for (int pl = 0; pl < PLANES; pl++) { j.nr[pl] = jb.cells[jr].nr[pl]; j.fr[pl] = jb.cells[jr].fr[pl]; }
        j.flat.recs.assign(8, F4{0, 0, 0, 0});                                      // record 0 = the subtree's root, record 1 = padding
        jb.emit(jr, 0, uv6, mat);
    }
    const double t2 = now();

    // statistics of the whole (reference-shaped) tree
    for (const Job& j : jobs) {
        out.nodes += j.flat.nodes; out.leaves += j.flat.leaves; out.empty_leaves += j.flat.empty_leaves; out.interior += j.flat.interior;
        out.max_depth_reached = std::max(out.max_depth_reached, j.flat.max_depth_reached);
        out.max_leaf_size = std::max(out.max_leaf_size, j.flat.max_leaf_size);
    }
    b.refit(root);
    out.recs.assign(4, F4{0, 0, 0, 0}); // record 0 = the root cell
    b.emit(root, 0, uv6, mat);
    // places of the jobs behind the top-level records (block starts stay 128-byte aligned: even record indices)
    size_t n_recs = out.recs.size() / 4, n_tris = out.tris.size() / 3;
    for (Job& j : jobs) {
        if (n_recs & 1) n_recs++;
        j.rec_base = n_recs;
        j.tri_base = n_tris;
        n_recs += j.flat.recs.size() / 4 - 2;
        n_tris += j.flat.tris.size() / 3;
    }
    out.recs.resize(n_recs * 4, F4{0, 0, 0, 0});
    out.tris.resize(n_tris * 3);
    out.shade.resize(n_tris * 2);
    out.orig.resize(n_tris);
    auto relocate = [](F4 q3, const Job& j) {           // link of a record: first triangle (leaf) or first child record
        uint32_t link = f2u(q3.z);
        const uint32_t meta = f2u(q3.w);
        link = (meta & RT_META_LEAF) ? link + (uint32_t)j.tri_base : (uint32_t)(j.rec_base + (link - 2));
        q3.z = u2f(link);
        return q3;
    };
#pragma omp parallel for schedule(dynamic, 1)
    for (long long ji = 0; ji < (long long)jobs.size(); ji++) {
        const Job& j = jobs[ji];
        const size_t nr = j.flat.recs.size() / 4 - 2, nt = j.flat.tris.size() / 3;
        for (size_t r = 0; r < nr; r++) {
            const F4* src = &j.flat.recs[(r + 2) * 4];
            F4* dst = &out.recs[(j.rec_base + r) * 4];
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = relocate(src[3], j);
        }
        if (nt) {
            memcpy(&out.tris[j.tri_base * 3], j.flat.tris.data(), nt * 3 * sizeof(F4));
            memcpy(&out.shade[j.tri_base * 2], j.flat.shade.data(), nt * 2 * sizeof(F4));
            memcpy(&out.orig[j.tri_base], j.flat.orig.data(), nt * sizeof(int32_t));
        }
    }
    for (const auto& slot : b.job_slots) {              // the jobs' root records, in the blocks of the cells above them
        const Job& j = jobs[slot.second];
        const F4* src = &j.flat.recs[0];
        F4* dst = &out.recs[(size_t)slot.first * 4];
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = relocate(src[3], j);
    }
    out.n_records = out.recs.size() / 4;
    build_top_table(out);
    if (getenv("RTB200_TRACE"))
        fprintf(stderr, "[rtb200] octree: top %.0f ms, %zu jobs %.0f ms, assemble %.0f ms\n", t1 - t0, jobs.size(), t2 - t1, now() - t2);
}

} // namespace rtb
