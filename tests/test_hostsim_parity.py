"""CPU-side parity of the KERNEL SOURCE: raytracercpp_b200/csrc/rt_device.h + octree_build.cpp compiled for the
host (tests/hostsim) and driven through the same C ABI and the same Python mirror as the CUDA library, against the
oracle.  Host libm == the oracle's libm, so everything here is bit-exact; the GPU run of the same assertions
(test_gpu_parity.py) relaxes only what CUDA's powf/atan2f/asinf can change."""
import ctypes

import numpy as np
import pytest

import raytracercpp_b200 as rt
from raytracercpp_b200 import api, scenes
from tests import common


@pytest.mark.parametrize("params", [(12, 40), (10, 8), (3, 2), (0, 5)])
def test_flattened_tree_and_closest_hit(hostsim_lib, oracle, robot, golden_rays, params):
    ctx = api.Context(0, hostsim_lib)
    ctx.set_triangles(robot["xyz9"], robot["uv6"], robot["mat"])
    info = ctx.build_bvh(*params)
    ob = oracle.bvh(robot["xyz9"], *params)
    st = ob.stats()
    for k in ("nodes", "leaves", "empty_leaves", "interior", "max_depth_reached", "max_leaf_size"):
        assert info[k] == st[k], k
    assert info["child_records"] >= st["nodes"] - st["empty_leaves"]
    o, d = golden_rays["o"], golden_rays["d"]
    got, want = ctx.intersect(o, d), ob.intersect(o, d)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    tag = f"_{params[0]}_{params[1]}"
    if "tri" + tag in golden_rays:                       # and straight against the reference's own vectors
        assert np.array_equal(got[0], golden_rays["tri" + tag]) and np.array_equal(got[1], golden_rays["t" + tag])


def test_edge_cases(hostsim_lib, oracle):
    ctx = api.Context(0, hostsim_lib)
    o, d = common.random_rays(2000, 1, (-1, -1, -5), (1, 1, -3))
    # empty scene
    ctx.set_triangles(np.zeros((0, 9), np.float32))
    ctx.build_bvh(12, 40)
    assert (ctx.intersect(o, d)[0] == -1).all()
    # one triangle, back-face culled from behind
    one = np.float32([[-1, -1, -4, 1, -1, -4, 0, 1, -4]])
    ctx.set_triangles(one)
    ctx.build_bvh(12, 40)
    tri, t, u, v = ctx.intersect([[0, 0, 0], [0, 0, -8]], [[0, 0, -1], [0, 0, 1]])
    assert list(tri) == [0, -1] and t[0] == 4.0
    # many coincident triangles: the cell can never split below the depth limit; ties keep the first (bvh.h:241)
    same = np.repeat(one, 100, 0)
    ctx.set_triangles(same)
    info = ctx.build_bvh(5, 8)
    ob = oracle.bvh(same, 5, 8)
    assert info["max_leaf_size"] == ob.stats()["max_leaf_size"] == 100 and info["max_depth_reached"] == 5
    a, b = ctx.intersect(o, d), ob.intersect(o, d)
    assert np.array_equal(a[0], b[0]) and set(np.unique(a[0])) <= {-1, 0}
    # ragged soup incl. degenerate (zero-area) triangles and rays with zero direction components
    soup = common.triangle_soup(3000, 8)
    soup[::50, 3:9] = np.tile(soup[::50, 0:3], 2)
    ctx.set_triangles(soup)
    ctx.build_bvh(8, 4)
    ob = oracle.bvh(soup, 8, 4)
    d2 = d.copy()
    d2[::3, 0] = 0
    d2[::5, 1] = 0
    d2[::7] = [0, 0, -1]
    for g, w in zip(ctx.intersect(o, d2), ob.intersect(o, d2)):
        assert np.array_equal(g, w)
    with pytest.raises(api.RtError):
        ctx.build_bvh(64, 8)


def test_call_order_and_unsupported_switches(hostsim_lib, robot):
    r = rt.Renderer(0, hostsim_lib)
    s = r.render_settings()
    s.image_width, s.image_height = 32, 32
    with pytest.raises(api.RtError) as e:
        r.ray_trace()                                     # no triangles / BVH yet
    assert e.value.code == api.RT_ERR_STATE
    r.set_triangles(robot["xyz9"], robot["uv6"], robot["mat"])
    with pytest.raises(api.RtError):
        r.ray_trace()                                     # RT_SHADING without materials (reference: assert, materials.h:117)
    r.set_materials(robot["materials"])
    r.ray_trace()
    s.enable_ssao = 1                                      # whole frames only: a tile shard's samples would read its neighbours' z-buffer
    with pytest.raises(api.RtError) as e:
        r.ctx.render_device(s, r.get_image().ctypes.data, 16, 2, 0)
    assert e.value.code == api.RT_ERR_UNSUPPORTED
    s.enable_ssao = 0
    s.enable_ao_mapping = 1
    with pytest.raises(api.RtError):
        r.ray_trace()                                     # mapping enabled, no map
    s.enable_ao_mapping = 0
    s.enable_skybox = 1
    with pytest.raises(api.RtError):
        r.ray_trace()                                     # skybox enabled, faces missing
    s.enable_skybox = 0
    s.max_recursion_depth = 99
    with pytest.raises(api.RtError):
        r.ray_trace()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg2_pom", "cfg3", "cfg3_mirror5", "cfg3_skybox", "cfg3_shapes"])
def test_frames_bit_exact(hostsim_lib, oracle, robot, golden_images, name):
    kw, mats, tex = common.config_table(robot["materials"])[name]
    img, stats = common.product_image(hostsim_lib, robot, kw, mats, tex)
    assert np.array_equal(img, common.oracle_image(oracle, robot, kw, mats, tex))
    assert np.array_equal(img, golden_images[name + "_strict"])
    f = kw.get("ssaa_factor", 1) if kw.get("enable_ssaa") else 1
    assert stats.primary_rays == kw["image_width"] * kw["image_height"] * f * f
    assert stats.shadow_rays == stats.primary_hits > 0
    r = common.oracle_renderer(oracle, robot, kw, mats, tex)
    cnt = r.count_rows()
    for k in ("primary_rays", "shadow_rays", "reflection_rays", "reflection_shadow_rays", "primary_hits"):
        assert getattr(stats, k) == cnt[k], k


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_debug_shading_modes(hostsim_lib, oracle, robot, mode):
    kw = dict(image_width=96, image_height=54, shading_method=mode, enable_ao_mapping=1)
    tex = {0: common.scenes.noise_texture((64, 64), 2)}
    img, _ = common.product_image(hostsim_lib, robot, kw, robot["materials"], tex)
    assert np.array_equal(img, common.oracle_image(oracle, robot, kw, robot["materials"], tex))


def test_moved_camera_light_and_f32_textures(hostsim_lib, oracle, robot):
    kw, mats, tex = common.config_table(robot["materials"])["cfg2"]
    tex = {k: (v.astype(np.float32) * np.float32(1 / 255.0)) for k, v in tex.items()}     # Image storage of the reference
    c, s = np.cos(0.4), np.sin(0.4)
    cam = np.float32([[c, 0, s, 1.0], [0, 1, 0, -0.5], [-s, 0, c, 0.5], [0, 0, 0, 1]])
    a, _ = common.product_image(hostsim_lib, robot, kw, mats, tex, cam=cam, light=(-2, 4, 1), fov=55.0)
    b = common.oracle_image(oracle, robot, kw, mats, tex, cam=cam, light=(-2, 4, 1), fov=55.0)
    assert np.array_equal(a, b)


def test_occlusion_matches_is_shadowed(hostsim_lib, oracle, robot):
    """rt_occluded == the shadow flag the oracle's is_shadowed produces for the same points (via two renders:
    shadows off vs on differ exactly where the point is shadowed and the direct term is non-zero)."""
    ctx = api.Context(0, hostsim_lib)
    ctx.set_triangles(robot["xyz9"], robot["uv6"], robot["mat"])
    ctx.build_bvh(12, 40)
    ctx.set_light(common.LIGHT)
    o, d = common.random_rays(30000, 5, (-1, -1, -5), (1, 1, -3))
    o[:] = 0
    tri, t, u, v = ctx.intersect(o, d)
    hit = tri >= 0
    p = (o + d * t[:, None])[hit]
    xyz = robot["xyz9"][tri[hit]]
    n = np.cross(xyz[:, 3:6] - xyz[:, 0:3], xyz[:, 6:9] - xyz[:, 0:3])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    occ = ctx.occluded(p, n)
    # brute force with the oracle: closest hit of the shadow ray, then the reference's distance predicate
    so = (p + n.astype(np.float32) * np.float32(1e-4)).astype(np.float32)
    sd = (np.float32(common.LIGHT) - p)
    sd = (sd / np.linalg.norm(sd, axis=1, keepdims=True)).astype(np.float32)
    bt, bt_t, _, _ = oracle.bvh(robot["xyz9"], 12, 40).intersect(so, sd)
    q = so + sd * bt_t[:, None]
    want = (bt >= 0) & (((p - q) ** 2).sum(1) < ((p - np.float32(common.LIGHT)) ** 2).sum(1))
    assert (occ == want).mean() >= 0.9995           # numpy's normalisation differs from the kernel's in the last ulp
    assert want.sum() > 100


def test_primary_rays_and_resolve(hostsim_lib, oracle, golden_images):
    ctx = api.Context(0, hostsim_lib)
    s = api.default_settings(hostsim_lib, image_width=64, image_height=36, enable_ssaa=1, ssaa_factor=2)
    ctx.set_camera(golden_images["proj_inv_80_16x9"], np.eye(4), (0, 0, 0))
    o, d = ctx.generate_primary_rays(s)
    assert o.shape == (128 * 72, 3) and np.allclose(np.linalg.norm(d, axis=1), 1, atol=1e-6)
    assert d[0, 0] < 0 and d[0, 1] < 0 and d[-1, 0] > 0 and d[-1, 1] > 0       # row 0 = bottom row (renderer.cpp:1086)
    for f in (2, 3, 4):
        assert np.array_equal(ctx.resolve_ssaa(golden_images["resolve_in"], f), golden_images[f"resolve_out_{f}"])
    with pytest.raises(api.RtError):
        ctx.resolve_ssaa(golden_images["resolve_in"], 5)                          # imageUtils.h:100-112


@pytest.mark.parametrize("mod", [2, 3, 8])
def test_tile_shards_tile_the_frame(hostsim_lib, robot, mod):
    kw, mats, tex = common.config_table(robot["materials"])["cfg2"]
    r = common.product_renderer(hostsim_lib, robot, kw, mats, tex)
    r.ray_trace()
    full = r.get_image().copy()
    s = r.render_settings()
    frame = np.zeros_like(full)
    total = 0
    for rem in range(mod):
        shard = np.full_like(full, 0xdeadbeef)
        st = r.ctx.render_device(s, shard.ctypes.data, 16, mod, rem)
        total += st.primary_rays
        n = r.ctx.tile_count(s, 16, mod, rem)
        staging = np.zeros(n * 16 * 16, np.uint32)
        r.ctx.pack_tiles(s, shard.ctypes.data, staging.ctypes.data, 16, mod, rem)
        r.ctx.unpack_tiles(s, frame.ctypes.data, staging.ctypes.data, 16, mod, rem)
    assert np.array_equal(frame, full)
    assert total == full.size * 4


def test_object_transform_rebuilds(hostsim_lib, oracle, robot):
    """Renderer::set_object_transform (renderer.cpp:214-224) == re-creating the triangles at the new placement."""
    kw = dict(image_width=96, image_height=54, compute_shadows=1)
    r = common.product_renderer(hostsim_lib, robot, kw, robot["materials"], {})
    m = np.eye(4, dtype=np.float32)
    m[0, 3], m[2, 3] = 0.5, -1.0
    r.set_object_transform(m)
    r.ray_trace()
    moved = dict(robot)
    pts = robot["xyz9"].reshape(-1, 3)
    moved["xyz9"] = (pts + np.float32([0.5, 0, -1.0])).astype(np.float32).reshape(-1, 9)
    assert np.array_equal(r.get_image(), common.oracle_image(oracle, moved, kw, robot["materials"], {}))


def test_cpp_adapter_compiles_against_the_abi(hostsim_lib, tmp_path):
    """include/rtb200_renderer.hpp (reference method names over the C ABI), linked against the host emulation here;
    the GPU run links the same example against librtb200.so (test_gpu_parity.py)."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    exe = tmp_path / "adapter_example"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(root / "include"), str(root / "tests" / "adapter_example.cpp"), "-o", str(exe),
                    "-L", str(root / "tests" / "hostsim"), "-lrtb200_hostsim", f"-Wl,-rpath,{root / 'tests' / 'hostsim'}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "adapter ok" in out


@pytest.mark.parametrize("split", [0, 1, 3, 16])
def test_leaf_refinement_never_changes_a_result(hostsim_lib, oracle, robot, golden_rays, split):
    """RT_OPT_LEAF_SPLIT only reshapes the device-side hierarchy below the reference's leaves: ids, t, u, v and frames
    stay bit-identical, including with the reference's own leaves (split = 0)."""
    ctx = api.Context(0, hostsim_lib)
    ctx.set_option(api.RT_OPT_LEAF_SPLIT, split)
    ctx.set_triangles(robot["xyz9"], robot["uv6"], robot["mat"])
    info = ctx.build_bvh(12, 40)
    assert info["nodes"] == 585 and info["max_leaf_size"] == 40          # statistics always describe the reference tree
    got = ctx.intersect(golden_rays["o"], golden_rays["d"])
    assert np.array_equal(got[0], golden_rays["tri_12_40"]) and np.array_equal(got[1], golden_rays["t_12_40"])
    same = np.repeat(np.float32([[-1, -1, -4, 1, -1, -4, 0, 1, -4]]), 100, 0)
    ctx.set_triangles(same)
    ctx.build_bvh(5, 8)
    o, d = common.random_rays(2000, 1, (-1, -1, -5), (1, 1, -3))
    assert set(np.unique(ctx.intersect(o, d)[0])) <= {-1, 0}            # a tie goes to the lowest index, bvh.h:241
    ctx.close()
    kw, mats, tex = common.config_table(robot["materials"])["cfg3"]
    r = common.product_renderer(hostsim_lib, robot, dict(kw, image_width=96, image_height=54), mats, tex)
    r.ctx.set_option(api.RT_OPT_LEAF_SPLIT, split)
    r.reconstruct_bvh_new()
    r.ray_trace()
    common.assert_image_close(r.get_image(), common.oracle_image(oracle, robot, dict(kw, image_width=96, image_height=54), mats, tex), what="split")
    r.close()


def test_hair_scene_frame_bit_exact(hostsim_lib, oracle):
    """BASELINE.json configs[4] at reduced size: thin double-sided Bezier strands (deep tree, incoherent rays), hard shadows."""
    from raytracercpp_b200 import scenes
    xyz9, uv6, mat = scenes.hair_ball(n_strands=600, segments=8, width=0.02)
    scene = dict(xyz9=xyz9, uv6=uv6, mat=mat)
    mats = rt.precompute_materials([scenes.DEFAULT_SPHERE_MATERIAL])
    kw = dict(image_width=96, image_height=54, compute_shadows=1, bvh_max_depth=12, bvh_leaf_object_count=40)
    img, stats = common.product_image(hostsim_lib, scene, kw, mats, {})
    assert np.array_equal(img, common.oracle_image(oracle, scene, kw, mats, {}))
    assert stats.primary_hits > 300 and stats.shadow_rays == stats.primary_hits


def test_skybox_faces_from_every_side(hostsim_lib, oracle, robot):
    """Skybox::sample (skybox.cpp:12-51) through all six faces: cameras turned left, right, up, down and backwards."""
    kw, mats, tex = common.config_table(robot["materials"])["cfg3_skybox"]
    kw = dict(kw, image_width=96, image_height=54)
    def rot(yaw, pitch):
        cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
        ry = np.float32([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]])
        rx = np.float32([[1, 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1]])
        return (ry @ rx).astype(np.float32)
    seen = set()
    for yaw, pitch in ((1.6, 0.0), (-1.6, 0.0), (0.3, 1.3), (0.3, -1.3), (3.1, 0.2)):
        cam = rot(yaw, pitch)
        img, _ = common.product_image(hostsim_lib, robot, kw, mats, tex, cam=cam)
        assert np.array_equal(img, common.oracle_image(oracle, robot, kw, mats, tex, cam=cam)), (yaw, pitch)
        seen.add(int(img[27, 48]))
    assert len(seen) >= 4                                  # the views really differ


@pytest.mark.parametrize("split", [8, 0])
def test_parallel_build_of_a_large_mesh(hostsim_lib, oracle, split):
    """octree_build.cpp hands subtrees to threads (jobs) that build and flatten them into buffers of their own; the top
    of the tree and the relocation of the jobs' links are assembled afterwards.  A 300 000-triangle mesh makes ~100 jobs:
    the tree statistics and the closest hit of 30 000 rays must equal the oracle's (= the reference's sequential
    insertion), with and without the device-side leaf refinement."""
    from raytracercpp_b200 import scenes
    xyz9, uv6, mat = scenes.displaced_sphere(*scenes.sphere_grid_for(300_000))
    ctx = api.Context(0, hostsim_lib)
    ctx.set_option(api.RT_OPT_LEAF_SPLIT, split)
    ctx.set_triangles(xyz9, uv6, mat)
    info = ctx.build_bvh(12, 40)
    ob = oracle.bvh(xyz9, 12, 40)
    st = ob.stats()
    for k in ("nodes", "leaves", "empty_leaves", "interior", "max_depth_reached", "max_leaf_size"):
        assert info[k] == st[k], k
    o, d = common.random_rays(30000, 9, (-1.2, -1.2, -4.2), (1.2, 1.2, -1.8))
    got, want = ctx.intersect(o, d), ob.intersect(o, d)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert (want[0] >= 0).sum() > 5000
    ctx.close()


def test_analytic_shapes_gate_order_and_refusals(hostsim_lib, oracle, robot):
    """trace_ray takes the CLOSEST of the BVH hit and the analytic shapes and only then applies min_t (renderer.cpp:
    1029-1040): a plane 0.05 in front of the camera hides the robot AND is itself rejected, so the frame is background; a
    far plane behind the robot changes only the background pixels.  Texture mapping and the barycentric / AO debug modes
    are refused while shapes exist (the reference reads a stale HitInfo::triangle there)."""
    mats = list(robot["materials"])
    kw = dict(image_width=96, image_height=54, compute_shadows=1)
    bg = 0xff000000 | (135 << 16) | (206 << 8) | 235
    near = {"shapes": [("plane", (0, 0, -0.05), (0, 0, 1), 0)]}
    img, st = common.product_image(hostsim_lib, robot, kw, mats, near)
    assert (img == bg).all() and st.primary_hits == 0
    assert np.array_equal(img, common.oracle_image(oracle, robot, kw, mats, near))
    far = {"shapes": [("plane", (0, 0, -9), (0, 0, 1), 1), ("sphere", (0, 0, 0), 0.08, 0)]}     # the sphere around the camera: t = 0.08, rejected too
    img2, st2 = common.product_image(hostsim_lib, robot, kw, mats, far)
    want2 = common.oracle_image(oracle, robot, kw, mats, far)
    assert np.array_equal(img2, want2)
    plain, _ = common.product_image(hostsim_lib, robot, kw, mats, {})
    assert (img2 == bg).all() and (plain != bg).any()        # the rejected sphere is the closest hit of every ray
    r = common.product_renderer(hostsim_lib, robot, dict(kw, enable_ao_mapping=1), mats, {0: scenes.noise_texture((32, 32), 2), "shapes": far["shapes"][:1]})
    with pytest.raises(api.RtError) as e:
        r.ray_trace()
    assert e.value.code == api.RT_ERR_UNSUPPORTED
    r.render_settings().enable_ao_mapping = 0
    r.render_settings().shading_method = api.RT_BARYCENTRIC_COORDINATES_SHADING
    with pytest.raises(api.RtError) as e:
        r.ray_trace()
    assert e.value.code == api.RT_ERR_UNSUPPORTED
    r.render_settings().shading_method = api.RT_ABS_NORMALS_SHADING
    r.ray_trace()
    want3 = common.oracle_image(oracle, robot, dict(kw, shading_method=api.RT_ABS_NORMALS_SHADING), mats, {"shapes": far["shapes"][:1]})
    assert np.array_equal(r.get_image(), want3)
    r.ctx.clear_analytic_shapes()
    r.render_settings().shading_method = api.RT_SHADING
    r.ray_trace()
    assert np.array_equal(r.get_image(), plain)
    r.close()


def test_batched_is_shadowed_sees_the_shapes(hostsim_lib, robot):
    """rt_occluded = Renderer::is_shadowed: a big sphere between the points and the light shadows points the robot alone does not."""
    ctx = api.Context(0, hostsim_lib)
    ctx.set_triangles(robot["xyz9"], robot["uv6"], robot["mat"])
    ctx.build_bvh(12, 40)
    ctx.set_light(common.LIGHT)
    rng = np.random.default_rng(3)
    p = (rng.uniform(-1, 1, size=(500, 3)) * np.float32([2, 0.2, 1]) + np.float32([0, -3.5, -4])).astype(np.float32)   # below the robot
    n = np.tile(np.float32([0, 1, 0]), (500, 1))
    before = ctx.occluded(p, n)
    ctx.add_sphere((1.5, 0.0, -1.0), 2.5, 0)             # swallows the way to the light at (3, 3, 2)
    after = ctx.occluded(p, n)
    assert after.all() and not before.all() and (after >= before).all()
    ctx.clear_analytic_shapes()
    assert np.array_equal(ctx.occluded(p, n), before)
    ctx.close()


@pytest.mark.parametrize("name", ["r_cfg1", "r_cfg2", "r_inside", "r_inside_noclip", "r_mirror", "r_big", "r_rough", "r_debug1", "r_debug2", "r_debug3", "r_debug4"])
def test_raster_trace_vs_oracle(hostsim_lib, oracle, robot, golden_raster, name):
    """The device source of the hybrid path (csrc/raster_device.h: clipping, piece set-up, fragment test, z-keys, shading of
    the winning piece) compiled for the host, through the C ABI (RtSettings::hybrid_rasterization_tracing) -- bit-exact
    against the oracle and against the frames of the compiled reference."""
    scene, kw, mats, tex, cam = common.raster_table(robot)[name]
    img, st = common.product_image(hostsim_lib, scene, kw, mats, tex, cam=cam)
    want = common.oracle_image(oracle, scene, kw, mats, tex, cam=cam)
    assert np.array_equal(img, want)
    if name in common.RASTER_PINNED:
        assert np.array_equal(img, golden_raster[name + "_reference"])
    if kw.get("shading_method", 0) == 0:
        assert st.primary_rays > 0 and st.shadow_rays == st.primary_hits


@pytest.mark.parametrize("name", ["ssao_ssaa2", "ssao_normal_mapped", "ssao_leftover_columns", "ssao_debug_shading", "r_ssao"])
def test_ssao_vs_oracle(hostsim_lib, oracle, robot, golden_ssao, golden_raster, name):
    """The device source of the SSAO pass (csrc/ssao_device.h: one lane of the reference's AVX2 loop, its scalar loop for the
    left-over columns, the 7x7 blur) compiled for the host, on the G-buffers the emulated shade stage / rasterizer writes:
    bit-exact against the oracle's per-pixel stream and against the frames stored with the reference's goldens."""
    table = common.ssao_table(robot["materials"])
    cam = None
    scene = robot
    if name == "ssao_leftover_columns":
        kw, mats, tex = table["ssao_ssaa2"]
        kw = dict(kw, image_width=99, image_height=57, enable_ssaa=0)
    elif name == "ssao_debug_shading":
        kw, mats, tex = table["ssao_ssaa2"]
        kw = dict(kw, shading_method=api.RT_ABS_NORMALS_SHADING)
    elif name == "r_ssao":
        scene, kw, mats, tex, cam = common.raster_table(robot)[name]
    else:
        kw, mats, tex = table[name]
    img, _ = common.product_image(hostsim_lib, scene, kw, mats, tex, cam=cam)
    if name == "r_ssao":
        want = common.oracle_renderer(oracle, scene, kw, mats, tex, cam=cam).raster()[0]
        assert np.array_equal(img, golden_raster[name + "_per_pixel"])
    else:
        want, _ = common.oracle_renderer(oracle, scene, kw, mats, tex).render_ssao()
        if name in table:
            assert np.array_equal(img, golden_ssao[name + "_per_pixel"])
    assert np.array_equal(img, want)
    plain, _ = common.product_image(hostsim_lib, scene, dict(kw, enable_ssao=0), mats, tex, cam=cam)
    assert (plain != img).mean() > 0.02


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_raster_trace_random_soups(hostsim_lib, oracle, seed):
    """raster_trace on triangle soups that surround the camera: triangles behind it, through it, across every clip plane
    (pieces up to the 12 the reference's arrays hold), degenerate ones, and coincident duplicates whose fragments tie on
    depth -- the first in (triangle, piece) order must win, as in the reference's sequential loop.  The host build of the
    device source against the oracle, and (where it is built) against the compiled reference on one thread, bit for bit."""
    from oracle import bindings
    rng = np.random.default_rng(seed)
    n = 260
    c = rng.uniform(-3, 3, size=(n, 1, 3)) + np.float32([0, 0, -1.0])
    soup = (c + rng.normal(scale=rng.uniform(0.1, 2.5, size=(n, 1, 1)), size=(n, 3, 3))).reshape(n, 9).astype(np.float32)
    soup[::17, 3:9] = np.tile(soup[::17, 0:3], 2)                    # degenerate
    dup = soup[5:45].copy()                                          # coincident copies with other materials, later in the order
    xyz9 = np.concatenate([soup, dup]).astype(np.float32)
    uv6 = rng.uniform(0, 1, size=(len(xyz9), 6)).astype(np.float32)
    mat = np.concatenate([np.zeros(n, np.int32), np.ones(len(dup), np.int32)])
    mats = rt.precompute_materials([dict(scenes.DEFAULT_SPHERE_MATERIAL, diffuse=(0.9, 0.2, 0.1)), dict(scenes.DEFAULT_SPHERE_MATERIAL, diffuse=(0.1, 0.3, 0.9))])
    scene = dict(xyz9=xyz9, uv6=uv6, mat=mat)
    for clipping in (1, 0):
        kw = dict(image_width=136, image_height=88, compute_shadows=1, hybrid_rasterization_tracing=1, enable_clipping=clipping,
                  enable_ssaa=int(seed == 2), ssaa_factor=2)
        img, st = common.product_image(hostsim_lib, scene, kw, mats, {})
        want = common.oracle_image(oracle, scene, kw, mats, {})
        assert np.array_equal(img, want), (seed, clipping)
        # without clipping the triangles behind the camera project through w < 0 to negative depths, win their pixels and are
        # then missed by the pixels' rays (black, as in the reference): only the clipped frame has hits to speak of
        assert st.primary_rays > 5000 and (st.primary_hits > 1000 or not clipping)
        if bindings.available("ref_strict"):
            ref = common.oracle_image(bindings.CpuTracer("ref_strict"), scene, kw, mats, {})
            assert np.array_equal(img, ref), (seed, clipping, "compiled reference")
