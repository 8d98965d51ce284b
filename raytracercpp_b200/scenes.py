"""Synthetic inputs of the benchmark configurations (BASELINE.json `configs`).

Everything is seeded numpy: procedural meshes (displaced subdivided sphere, hair strands), procedural textures
(none ship with the reference: tp2/data has no diffuse/normal/roughness map and the AO map / skysphere blobs are
stripped), and the GUI's default placement of the robot.  The same arrays feed the CUDA path, the oracle and the
compiled reference.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# value noise / fbm on points of R^3 (float32, vectorised)
def _hash3(ix, iy, iz, seed):
    h = (ix.astype(np.uint32) * np.uint32(0x8da6b343)) ^ (iy.astype(np.uint32) * np.uint32(0xd8163841)) ^ \
        (iz.astype(np.uint32) * np.uint32(0xcb1ab31f)) ^ np.uint32(seed * 0x9e3779b9 & 0xffffffff)
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85ebca6b)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xc2b2ae35)
    h ^= h >> np.uint32(16)
    return h.astype(np.float32) * np.float32(1.0 / 4294967296.0)


def value_noise3(p, seed):
    p = np.asarray(p, np.float32)
    f = np.floor(p)
    t = p - f
    t = t * t * (np.float32(3) - np.float32(2) * t)
    i = f.astype(np.int64)
    ix, iy, iz = i[..., 0], i[..., 1], i[..., 2]
    out = np.zeros(p.shape[:-1], np.float32)
    for dx in (0, 1):
        wx = t[..., 0] if dx else 1 - t[..., 0]
        for dy in (0, 1):
            wy = t[..., 1] if dy else 1 - t[..., 1]
            for dz in (0, 1):
                wz = t[..., 2] if dz else 1 - t[..., 2]
                out += wx * wy * wz * _hash3(ix + dx, iy + dy, iz + dz, seed)
    return out


def fbm3(p, seed, octaves=4):
    p = np.asarray(p, np.float32)
    amp, total, norm = np.float32(1.0), np.zeros(p.shape[:-1], np.float32), np.float32(0)
    for o in range(octaves):
        total += amp * value_noise3(p * np.float32(2 ** o), seed + 101 * o)
        norm += amp
        amp *= np.float32(0.5)
    return total / norm


# ---------------------------------------------------------------------------------------------------------------
def displaced_sphere(n_lat: int, n_lon: int, radius=1.0, center=(0.0, 0.0, -3.0), displacement=0.05, seed=7):
    """Lat-long grid sphere displaced along the normal by `displacement * fbm(p * 8)`: 2 * n_lat * n_lon triangles,
    outward winding (the reference culls back faces, triangle.cpp:37-40), per-vertex UVs = (lon, lat) in [0,1]."""
    lat = np.linspace(0.0, np.pi, n_lat + 1, dtype=np.float64)
    lon = np.linspace(0.0, 2 * np.pi, n_lon + 1, dtype=np.float64)
    st, ct = np.sin(lat)[:, None], np.cos(lat)[:, None]
    unit = np.stack([st * np.cos(lon)[None, :], ct * np.ones_like(lon)[None, :], st * np.sin(lon)[None, :]], -1).astype(np.float32)
    # wrap the seam so both copies of a seam vertex get the same displacement
    r = np.float32(radius) + np.float32(displacement) * (fbm3(unit * np.float32(8.0), seed) - np.float32(0.5))
    r[:, -1] = r[:, 0]
    r[0, :] = r[0, 0]
    r[-1, :] = r[-1, 0]
    verts = unit * r[..., None] + np.asarray(center, np.float32)
    uu = np.broadcast_to((lon / (2 * np.pi)).astype(np.float32)[None, :], r.shape)
    vv = np.broadcast_to((1.0 - lat / np.pi).astype(np.float32)[:, None], r.shape)

    def quad(a):
        return a[:-1, :-1], a[1:, :-1], a[1:, 1:], a[:-1, 1:]
    p00, p10, p11, p01 = quad(verts)
    u00, u10, u11, u01 = quad(uu)
    v00, v10, v11, v01 = quad(vv)
    n = n_lat * n_lon
    xyz9 = np.empty((2 * n, 9), np.float32)
    uv6 = np.empty((2 * n, 6), np.float32)
    # winding: (p00, p11, p10) and (p00, p01, p11) face outward for this parametrisation
    xyz9[0::2] = np.concatenate([p00, p11, p10], -1).reshape(n, 9)
    xyz9[1::2] = np.concatenate([p00, p01, p11], -1).reshape(n, 9)
    uv6[0::2] = np.stack([u00, u11, u10, v00, v11, v10], -1).reshape(n, 6)
    uv6[1::2] = np.stack([u00, u01, u11, v00, v01, v11], -1).reshape(n, 6)
    # make the winding outward wherever the displaced quad flipped it
    a, b, c = xyz9[:, 0:3], xyz9[:, 3:6], xyz9[:, 6:9]
    nrm = np.cross(b - a, c - a)
    out = np.einsum("ij,ij->i", nrm, (a + b + c) / 3 - np.asarray(center, np.float32))
    flip = out < 0
    if flip.any():
        xyz9[flip] = xyz9[flip][:, [0, 1, 2, 6, 7, 8, 3, 4, 5]]
        uv6[flip] = uv6[flip][:, [0, 2, 1, 3, 5, 4]]
    mat = np.zeros(2 * n, np.int32)
    return xyz9, uv6, mat


def sphere_grid_for(n_triangles: int):
    """(n_lat, n_lon) with n_lon = 2 * n_lat and 2 * n_lat * n_lon >= n_triangles (10 M -> 1582 x 3164)."""
    n_lat = int(np.ceil(np.sqrt(n_triangles / 4.0)))
    return n_lat, 2 * n_lat


def hair_ball(n_strands=31250, segments=16, width=1.0e-3, root_radius=0.8, length=0.6, center=(0.0, 0.0, -3.0), seed=11):
    """~1 M thin triangles: cubic Bezier strands (de Casteljau, as tp3Courbes/main.cpp:21-43 authored the reference's
    hair) rooted on a sphere, each segment a ribbon quad of `width`, duplicated with flipped winding because the
    reference culls back faces.  n_strands * segments * 4 triangles."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n_strands, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    root = d * root_radius
    jitter = lambda s: rng.normal(scale=s, size=(n_strands, 3))
    p0 = root
    p1 = root + d * (length / 3) + jitter(0.05)
    p2 = root + d * (2 * length / 3) + jitter(0.10) + np.array([0, -0.10, 0])
    p3 = root + d * length + jitter(0.15) + np.array([0, -0.25, 0])
    ts = np.linspace(0, 1, segments + 1)[None, :, None]
    a, b, c = p0[:, None] * (1 - ts) + p1[:, None] * ts, p1[:, None] * (1 - ts) + p2[:, None] * ts, p2[:, None] * (1 - ts) + p3[:, None] * ts
    ab, bc = a * (1 - ts) + b * ts, b * (1 - ts) + c * ts
    curve = ab * (1 - ts) + bc * ts                                     # [strand, segments + 1, 3]
    tangent = np.gradient(curve, axis=1)
    side = np.cross(tangent, d[:, None, :])
    side /= np.maximum(np.linalg.norm(side, axis=2, keepdims=True), 1e-12)
    left, right = curve - side * (width / 2), curve + side * (width / 2)
    l0, l1, r0, r1 = left[:, :-1], left[:, 1:], right[:, :-1], right[:, 1:]
    front = np.concatenate([np.concatenate([l0, r0, r1], -1), np.concatenate([l0, r1, l1], -1)], 1).reshape(-1, 9)
    back = front[:, [0, 1, 2, 6, 7, 8, 3, 4, 5]]
    xyz9 = (np.concatenate([front, back], 0) + np.tile(np.asarray(center), 3)).astype(np.float32)
    uv6 = None
    mat = np.zeros(len(xyz9), np.int32)
    return xyz9, uv6, mat


# ---------------------------------------------------------------------------------------------------------------
def noise_texture(size, seed, channels="grey"):
    """u8 RGBA value-noise texture (size = (h, w))."""
    h, w = size
    y, x = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    p = np.stack([x * np.float32(16.0 / w), y * np.float32(16.0 / h), np.zeros_like(x)], -1)
    if channels == "grey":
        g = fbm3(p, seed)
        rgb = np.stack([g, g, g], -1)
    else:
        rgb = np.stack([fbm3(p, seed + 1000 * k) for k in range(3)], -1)
    out = np.empty((h, w, 4), np.uint8)
    out[..., :3] = np.clip(rgb * 255.0, 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def normal_map_texture(size, seed, strength=2.0):
    """Tangent-space normal map from the gradient of a value-noise height field, u8 RGBA."""
    h, w = size
    y, x = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    hf = fbm3(np.stack([x * np.float32(16.0 / w), y * np.float32(16.0 / h), np.zeros_like(x)], -1), seed)
    gy, gx = np.gradient(hf)
    n = np.stack([-gx * strength * w / 16, -gy * strength * h / 16, np.ones_like(hf)], -1)
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    out = np.empty((h, w, 4), np.uint8)
    out[..., :3] = np.clip((n * 0.5 + 0.5) * 255.0, 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def sky_texture(size=(2048, 4096), sun_dir=(0.3, 0.6, -0.7)):
    """Equirectangular gradient sky with a sun disc, u8 RGBA (the reference's skysphere.jpg blob is stripped)."""
    h, w = size
    v, u = np.meshgrid((np.arange(h, dtype=np.float32) + 0.5) / h, (np.arange(w, dtype=np.float32) + 0.5) / w, indexing="ij")
    theta = (u - 0.5) * 2 * np.pi
    phi = (v - 0.5) * np.pi
    # inverse of the reference's lookup u = 0.5 + atan2(-dz, -dx)/(2 pi), v = 0.5 + asin(-dy)/pi (renderer.cpp:1056-1057)
    dy = -np.sin(phi)
    dx = -np.cos(phi) * np.cos(theta)
    dz = -np.cos(phi) * np.sin(theta)
    sun = np.asarray(sun_dir, np.float32)
    sun = sun / np.linalg.norm(sun)
    cosang = dx * sun[0] + dy * sun[1] + dz * sun[2]
    t = np.clip(dy * 0.5 + 0.5, 0, 1)
    rgb = np.stack([0.55 - 0.35 * t, 0.70 - 0.25 * t, 0.95 - 0.05 * t], -1)
    rgb = np.where((dy < 0)[..., None], np.array([0.35, 0.32, 0.30], np.float32) * (1 + dy[..., None]), rgb)
    rgb = rgb + np.clip((cosang - 0.995) / 0.005, 0, 1)[..., None] * np.array([1.0, 0.9, 0.6], np.float32)
    out = np.empty((h, w, 4), np.uint8)
    out[..., :3] = np.clip(rgb * 255.0, 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


DEFAULT_SPHERE_MATERIAL = dict(ambient_coeff=(1, 1, 1), diffuse=(0.74, 0.36, 0.05), specular=(0.5, 0.5, 0.5),
                               emission=(0, 0, 0), reflection=0.0, roughness=0.0, ns=100.0)
