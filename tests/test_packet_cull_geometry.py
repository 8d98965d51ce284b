"""The geometry of the packet kernels' bounding-pyramid cull (csrc/kernels.cuh: packet_set_cull and the two cull passes of
packet_trace), restated in numpy float32 with the kernel's formulas and margins, against exact (float64) ray / box and
ray / triangle tests: a box or triangle that ANY ray of the packet touches must never be dropped -- for primary packets
(rays from the camera through a pixel block) and for shadow packets (rays that start EPSILON off the hit points and aim
from the hit points at the light, renderer.cpp:344, so they miss the apex by up to 1e-4).  The CUDA code itself is
exercised by the GPU parity tests (frames bit-identical with the cull on and off); this file pins the construction."""
import numpy as np
import pytest

F = np.float32


def normalize(v):
    return (v / np.sqrt((v * v).sum(-1, keepdims=True))).astype(F)


def packet_planes(apex, points, slack):
    """packet_set_cull: four planes (n, w), inside = n.x - w >= 0; None when the packet has no pyramid."""
    apex, points = apex.astype(F), points.astype(F)
    v = (points - apex).astype(F)
    ez = normalize(v[0])
    helper = np.array([1, 0, 0], F) if abs(ez[0]) < 0.5 else np.array([0, 1, 0], F)
    ex = normalize(np.cross(helper, ez).astype(F))
    ey = np.cross(ez, ex).astype(F)
    vz, vx, vy = (v @ ez).astype(F), (v @ ex).astype(F), (v @ ey).astype(F)
    if not np.all((vz > 0) & (vz * vz > F(0.01) * (v * v).sum(-1))):
        return None
    tx, ty = (vx / vz).astype(F), (vy / vz).astype(F)
    x0, x1, y0, y1 = tx.min(), tx.max(), ty.min(), ty.max()
    wx = F(1.0e-5) * (F(1) + max(abs(x0), abs(x1)))
    wy = F(1.0e-5) * (F(1) + max(abs(y0), abs(y1)))
    x0, x1, y0, y1 = F(x0 - wx), F(x1 + wx), F(y0 - wy), F(y1 + wy)
    planes = []
    for n in ((ex - x0 * ez), (x1 * ez - ex), (ey - y0 * ez), (y1 * ez - ey)):
        n = normalize(n.astype(F))
        planes.append((n, F(F(n @ apex) - F(slack))))
    return planes


def box_dropped(planes, lo, hi):
    for n, w in planes:
        c = np.where(n > 0, hi, lo).astype(F)
        a = (n * c).astype(F)
        if F(a.sum(dtype=F) - w) < F(-4.0e-6) * F(np.abs(a).sum(dtype=F) + abs(w)):
            return True
    return False


def tri_dropped(planes, tri):
    for n, w in planes:
        d = ((tri.astype(F) * n).sum(-1, dtype=F) - w).astype(F)
        mag = F(abs(w) + np.abs(tri).max(0).sum(dtype=F))
        if d.max() < F(-4.0e-6) * mag:
            return True
    return False


def segment_hits_box(o, e, lo, hi):
    """Does the segment o -> e touch the box (float64 slabs)?"""
    o, d = o.astype(np.float64), (e - o).astype(np.float64)
    t0, t1 = 0.0, 1.0
    for k in range(3):
        if abs(d[k]) < 1e-300:
            if o[k] < lo[k] or o[k] > hi[k]:
                return False
            continue
        a, b = (lo[k] - o[k]) / d[k], (hi[k] - o[k]) / d[k]
        t0, t1 = max(t0, min(a, b)), min(t1, max(a, b))
    return t0 <= t1


def segment_hits_triangle(o, e, tri):
    o, d = o.astype(np.float64), (e - o).astype(np.float64)
    a, b, c = tri.astype(np.float64)
    n = np.cross(b - a, c - a)
    den = n @ d
    if abs(den) < 1e-300:
        return False
    t = (n @ (a - o)) / den
    if t < 0 or t > 1:
        return False
    p = o + t * d
    return all(np.cross(q - r, p - r) @ n >= 0 for r, q in ((a, b), (b, c), (c, a)))


@pytest.mark.parametrize("seed", range(6))
def test_primary_pyramid_never_drops_a_touched_box_or_triangle(seed):
    rng = np.random.default_rng(seed)
    dropped = 0
    for _ in range(60):
        cam = rng.uniform(-2, 2, 3).astype(F)
        axis = normalize(rng.normal(size=3).astype(F))
        u = normalize(np.cross(axis, rng.normal(size=3)).astype(F))
        w = np.cross(axis, u).astype(F)
        spread = 10.0 ** rng.uniform(-4, -1)                      # an 8 x 4 pixel block: from 16 spp at 4K to a coarse frame
        ij = np.stack(np.meshgrid(np.arange(8), np.arange(4)), -1).reshape(-1, 2)
        dirs = normalize(axis + spread * ((ij[:, :1] - 3.5) * u + (ij[:, 1:] - 1.5) * w))
        planes = packet_planes(cam, (cam + dirs).astype(F), 0.0)
        assert planes is not None
        far = (cam + F(50) * dirs).astype(F)
        for _ in range(40):
            # boxes around points ON the rays (touched) and near them (mostly not): sizes from a deep octree cell to the scene
            k, t = rng.integers(0, 32), rng.uniform(0.05, 20)
            size = 10.0 ** rng.uniform(-4, 0.5)
            centre = cam + t * dirs[k] + rng.normal(size=3) * size * rng.choice([0.0, 0.6, 3.0])
            lo, hi = (centre - size * rng.uniform(0.1, 1, 3)).astype(F), (centre + size * rng.uniform(0.1, 1, 3)).astype(F)
            touched = any(segment_hits_box(cam, far[r], lo, hi) for r in range(32))
            if box_dropped(planes, lo, hi):
                dropped += 1
                assert not touched
            tri = (centre + rng.normal(size=(3, 3)) * size).astype(F)
            if tri_dropped(planes, tri):
                dropped += 1
                assert not any(segment_hits_triangle(cam, far[r], tri) for r in range(32))
    assert dropped > 300                                        # the cull does drop things


@pytest.mark.parametrize("seed", range(6))
def test_shadow_pyramid_never_drops_a_touched_box_or_triangle(seed):
    rng = np.random.default_rng(100 + seed)
    dropped = 0
    for _ in range(60):
        light = rng.uniform(-4, 4, 3).astype(F)
        base = (light + normalize(rng.normal(size=3).astype(F)) * F(rng.uniform(1.5, 9))).astype(F)
        spread = 10.0 ** rng.uniform(-3.5, -0.5)                  # the hit points of 32 neighbouring pixels
        p = (base + rng.normal(size=(32, 3)) * spread).astype(F)
        nrm = normalize(rng.normal(size=(32, 3)).astype(F))
        so = (p + F(1.0e-4) * nrm).astype(F)                      # Renderer::EPSILON, renderer.h:23
        sd = normalize((light - p).astype(F))                     # ... but the direction is taken from the hit point
        dist = np.sqrt(((p - light) ** 2).sum(-1)).astype(F)
        end = (so + ((dist + F(1.0e-4)) * sd.T).T).astype(F)      # as far as a hit can still satisfy |p - q| < |p - light|
        planes = packet_planes(light, so, 2.0e-4)
        if planes is None:
            continue
        for _ in range(40):
            k, t = rng.integers(0, 32), rng.uniform(0, 1)
            size = 10.0 ** rng.uniform(-4, 0.3)
            centre = so[k] + t * (end[k] - so[k]) + rng.normal(size=3) * size * rng.choice([0.0, 0.6, 3.0])
            lo, hi = (centre - size * rng.uniform(0.1, 1, 3)).astype(F), (centre + size * rng.uniform(0.1, 1, 3)).astype(F)
            if box_dropped(planes, lo, hi):
                dropped += 1
                assert not any(segment_hits_box(so[r], end[r], lo, hi) for r in range(32))
            tri = (centre + rng.normal(size=(3, 3)) * size).astype(F)
            if tri_dropped(planes, tri):
                dropped += 1
                assert not any(segment_hits_triangle(so[r], end[r], tri) for r in range(32))
    assert dropped > 300
