"""The parity tests proper: the CUDA path, through the C ABI of librtb200.so, against the oracle, the golden vectors
of the compiled reference, and -- at the full BASELINE.json sizes -- size-independent properties.

Tolerances (north_star): identical hit-triangle ids on >= 99.9 % of rays, hit distance within 1e-4 relative, final
RGB within 1/255 mean and 4/255 max per channel.  The traversal and the triangle test use no libm call and are
compiled without FMA contraction, so ids / t / u / v are in fact asserted BIT-EXACT; only shading (powf, atan2f,
asinf) is given the tolerance.
"""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

import raytracercpp_b200 as rt
from raytracercpp_b200 import api, scenes
from tests import common

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_cuda_library_is_the_one_loaded(cuda_lib):
    assert Path(cuda_lib._name).name == "librtb200.so"
    ctx = api.Context(0, cuda_lib)
    ctx.close()


@pytest.mark.parametrize("params", [(12, 40), (10, 8), (3, 2), (0, 5)])
def test_closest_hit_vs_oracle_and_reference_vectors(cuda_lib, oracle, robot, golden_rays, params):
    ctx = api.Context(0, cuda_lib)
    ctx.set_triangles(robot["xyz9"], robot["uv6"], robot["mat"])
    info = ctx.build_bvh(*params)
    ob = oracle.bvh(robot["xyz9"], *params)
    st = ob.stats()
    for k in ("nodes", "leaves", "empty_leaves", "interior", "max_depth_reached", "max_leaf_size"):
        assert info[k] == st[k], k
    o, d = golden_rays["o"], golden_rays["d"]
    got, want = ctx.intersect(o, d), ob.intersect(o, d)
    assert (got[0] == want[0]).mean() >= 0.999
    hit = (want[0] >= 0) & (got[0] == want[0])
    assert np.all(np.abs(got[1][hit] - want[1][hit]) <= 1e-4 * np.abs(want[1][hit]))
    for g, w in zip(got, want):                           # in fact bit-exact
        assert np.array_equal(g, w)
    tag = f"_{params[0]}_{params[1]}"
    if "tri" + tag in golden_rays:
        assert np.array_equal(got[0], golden_rays["tri" + tag]) and np.array_equal(got[1], golden_rays["t" + tag])
        assert (got[0] == golden_rays["tri_fma" + tag]).mean() >= 0.999     # reference built with its own flags
    ctx.close()


def test_edge_cases(cuda_lib, oracle):
    ctx = api.Context(0, cuda_lib)
    o, d = common.random_rays(2000, 1, (-1, -1, -5), (1, 1, -3))
    ctx.set_triangles(np.zeros((0, 9), np.float32))
    ctx.build_bvh(12, 40)
    assert (ctx.intersect(o, d)[0] == -1).all()
    assert ctx.intersect(np.zeros((0, 3)), np.zeros((0, 3)))[0].size == 0
    one = np.float32([[-1, -1, -4, 1, -1, -4, 0, 1, -4]])
    ctx.set_triangles(one)
    ctx.build_bvh(12, 40)
    tri, t, u, v = ctx.intersect([[0, 0, 0], [0, 0, -8]], [[0, 0, -1], [0, 0, 1]])
    assert list(tri) == [0, -1] and t[0] == 4.0             # second ray: back-face culled (tests.cpp:109)
    same = np.repeat(one, 100, 0)
    ctx.set_triangles(same)
    info = ctx.build_bvh(5, 8)
    assert info["max_leaf_size"] == 100 and info["max_depth_reached"] == 5
    a = ctx.intersect(o, d)
    assert set(np.unique(a[0])) <= {-1, 0}                   # ties keep the first triangle (bvh.h:241)
    soup = common.triangle_soup(3000, 8)
    soup[::50, 3:9] = np.tile(soup[::50, 0:3], 2)            # zero-area triangles
    ctx.set_triangles(soup)
    ctx.build_bvh(8, 4)
    ob = oracle.bvh(soup, 8, 4)
    d2 = d.copy()
    d2[::3, 0] = 0
    d2[::5, 1] = 0
    d2[::7] = [0, 0, -1]
    for g, w in zip(ctx.intersect(o, d2), ob.intersect(o, d2)):
        assert np.array_equal(g, w)
    ctx.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg2_pom", "cfg3", "cfg3_mirror5", "cfg3_skybox", "cfg3_shapes"])
def test_frames_vs_oracle_and_reference(cuda_lib, oracle, robot, golden_images, name):
    kw, mats, tex = common.config_table(robot["materials"])[name]
    img, stats = common.product_image(cuda_lib, robot, kw, mats, tex)
    want = common.oracle_image(oracle, robot, kw, mats, tex)
    common.assert_image_close(img, want, what=name + " vs oracle")
    common.assert_image_close(img, golden_images[name + "_strict"], what=name + " vs reference (no contraction)")
    common.assert_image_close(img, golden_images[name + "_fma"], what=name + " vs reference (its own flags)")
    assert (img == want).mean() >= 0.99
    r = common.oracle_renderer(oracle, robot, kw, mats, tex)
    cnt = r.count_rows()
    assert stats.primary_rays == cnt["primary_rays"] and stats.primary_hits == cnt["primary_hits"]
    assert stats.shadow_rays == cnt["shadow_rays"]
    # ray trees can differ where a powf/atan2f ulp flips nothing: the fan's ray counts are exact
    assert stats.reflection_rays == cnt["reflection_rays"] and stats.reflection_shadow_rays == cnt["reflection_shadow_rays"]
    assert stats.kernel_launches >= 2


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_debug_shading_modes(cuda_lib, oracle, robot, mode):
    kw = dict(image_width=96, image_height=54, shading_method=mode, enable_ao_mapping=1)
    tex = {0: scenes.noise_texture((64, 64), 2)}
    img, _ = common.product_image(cuda_lib, robot, kw, robot["materials"], tex)
    assert np.array_equal(img, common.oracle_image(oracle, robot, kw, robot["materials"], tex))


def test_moved_camera_light_and_f32_textures(cuda_lib, oracle, robot):
    kw, mats, tex = common.config_table(robot["materials"])["cfg2"]
    tex = {k: (v.astype(np.float32) * np.float32(1 / 255.0)) for k, v in tex.items()}
    c, s = np.cos(0.4), np.sin(0.4)
    cam = np.float32([[c, 0, s, 1.0], [0, 1, 0, -0.5], [-s, 0, c, 0.5], [0, 0, 0, 1]])
    a, _ = common.product_image(cuda_lib, robot, kw, mats, tex, cam=cam, light=(-2, 4, 1), fov=55.0)
    b = common.oracle_image(oracle, robot, kw, mats, tex, cam=cam, light=(-2, 4, 1), fov=55.0)
    common.assert_image_close(a, b, what="moved camera")


def test_shadow_rays_vs_reference_predicate(cuda_lib, oracle, robot):
    ctx = api.Context(0, cuda_lib)
    ctx.set_triangles(robot["xyz9"], robot["uv6"], robot["mat"])
    ctx.build_bvh(12, 40)
    ctx.set_light(common.LIGHT)
    o, d = common.random_rays(60000, 5, (-1, -1, -5), (1, 1, -3))
    o[:] = 0
    tri, t, u, v = ctx.intersect(o, d)
    hit = tri >= 0
    p = (o + d * t[:, None])[hit]
    xyz = robot["xyz9"][tri[hit]]
    n = np.cross(xyz[:, 3:6] - xyz[:, 0:3], xyz[:, 6:9] - xyz[:, 0:3])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    occ = ctx.occluded(p, n)
    so = (p + n.astype(np.float32) * np.float32(1e-4)).astype(np.float32)
    sd = np.float32(common.LIGHT) - p
    sd = (sd / np.linalg.norm(sd, axis=1, keepdims=True)).astype(np.float32)
    bt, bt_t, _, _ = oracle.bvh(robot["xyz9"], 12, 40).intersect(so, sd)
    q = so + sd * bt_t[:, None]
    want = (bt >= 0) & (((p - q) ** 2).sum(1) < ((p - np.float32(common.LIGHT)) ** 2).sum(1))
    assert (occ == want).mean() >= 0.999
    assert want.sum() > 100
    ctx.close()


def test_primary_rays_resolve_and_call_order(cuda_lib, golden_images, robot):
    ctx = api.Context(0, cuda_lib)
    s = api.default_settings(cuda_lib, image_width=64, image_height=36, enable_ssaa=1, ssaa_factor=2)
    ctx.set_camera(golden_images["proj_inv_80_16x9"], np.eye(4), (0, 0, 0))
    o, d = ctx.generate_primary_rays(s)
    assert o.shape == (128 * 72, 3) and d[0, 0] < 0 and d[0, 1] < 0 and d[-1, 0] > 0 and d[-1, 1] > 0
    for f in (2, 3, 4):
        assert np.array_equal(ctx.resolve_ssaa(golden_images["resolve_in"], f), golden_images[f"resolve_out_{f}"])
    with pytest.raises(api.RtError):
        ctx.resolve_ssaa(golden_images["resolve_in"], 5)
    with pytest.raises(api.RtError) as e:
        ctx.render(s)
    assert e.value.code == api.RT_ERR_STATE
    ctx.close()
    r = rt.Renderer(0, cuda_lib)
    st = r.render_settings()
    st.image_width, st.image_height = 32, 32
    r.set_triangles(robot["xyz9"], robot["uv6"], robot["mat"])
    with pytest.raises(api.RtError):
        r.ray_trace()                                          # RT_SHADING without materials
    r.set_materials(robot["materials"])
    r.ray_trace()
    r.close()


@pytest.mark.parametrize("mod", [2, 8])
def test_tile_shards_tile_the_frame(cuda_lib, robot, mod):
    import torch
    kw, mats, tex = common.config_table(robot["materials"])["cfg2"]
    r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
    r.ray_trace()
    full = r.get_image().copy()
    s = r.render_settings()
    frame = torch.zeros(full.shape, dtype=torch.int32, device="cuda")
    for rem in range(mod):
        shard = torch.full(full.shape, 0x5a5a5a5a, dtype=torch.int32, device="cuda")
        r.ctx.render_device(s, shard.data_ptr(), 16, mod, rem)
        n = r.ctx.tile_count(s, 16, mod, rem)
        staging = torch.zeros(n * 256, dtype=torch.int32, device="cuda")
        r.ctx.pack_tiles(s, shard.data_ptr(), staging.data_ptr(), 16, mod, rem)
        r.ctx.unpack_tiles(s, frame.data_ptr(), staging.data_ptr(), 16, mod, rem)
    torch.cuda.synchronize()
    assert np.array_equal(frame.cpu().numpy().view(np.uint32), full)
    # the same through the single-launch receiving side of the all-gather (rt_unpack_gathered)
    slot = max(r.ctx.tile_count(s, 16, mod, rem) for rem in range(mod)) * 256
    gathered = torch.zeros(mod * slot, dtype=torch.int32, device="cuda")
    shard = torch.from_numpy(full.view(np.int32)).cuda()
    for rem in range(mod):
        r.ctx.pack_tiles(s, shard.data_ptr(), gathered.data_ptr() + rem * slot * 4, 16, mod, rem)
    again = torch.full(full.shape, 0x5a5a5a5a, dtype=torch.int32, device="cuda")
    r.ctx.unpack_gathered(s, again.data_ptr(), gathered.data_ptr(), 16, mod, 1)
    torch.cuda.synchronize()
    got = again.cpu().numpy().view(np.uint32)
    mine = np.zeros(full.shape, bool)
    ty, tx = np.divmod(np.arange(full.size).reshape(full.shape), full.shape[1])
    own = torch.zeros(full.shape, dtype=torch.int32, device="cuda")
    stg = torch.zeros(slot, dtype=torch.int32, device="cuda")
    ones = torch.ones(full.shape, dtype=torch.int32, device="cuda")
    r.ctx.pack_tiles(s, ones.data_ptr(), stg.data_ptr(), 16, mod, 1)
    r.ctx.unpack_tiles(s, own.data_ptr(), stg.data_ptr(), 16, mod, 1)
    torch.cuda.synchronize()
    mine = own.cpu().numpy() == 1                                 # pixels of shard 1: left untouched by self_rem = 1
    assert np.array_equal(got[~mine], full[~mine]) and (got[mine] == 0x5a5a5a5a).all() and mine.any()
    r.close()


def test_object_transform_and_counters(cuda_lib, oracle, robot):
    kw = dict(image_width=96, image_height=54, compute_shadows=1)
    r = common.product_renderer(cuda_lib, robot, kw, robot["materials"], {})
    r.ctx.set_option(api.RT_OPT_COUNT_WORK, 1)
    m = np.eye(4, dtype=np.float32)
    m[0, 3], m[2, 3] = 0.5, -1.0
    r.set_object_transform(m)
    r.ray_trace()
    st = r.last_stats()
    moved = dict(robot)
    moved["xyz9"] = (robot["xyz9"].reshape(-1, 3) + np.float32([0.5, 0, -1.0])).astype(np.float32).reshape(-1, 9)
    common.assert_image_close(r.get_image(), common.oracle_image(oracle, moved, kw, robot["materials"], {}), what="moved object")
    assert st.primary_volume_tests >= st.primary_rays and st.shadow_triangle_tests > 0 and st.work_bytes > 0
    # the same frame with tiny wavefront chunks is the same frame
    img_a = r.get_image().copy()
    r.ctx.set_option(api.RT_OPT_COUNT_WORK, 0)
    r.ctx.set_option(api.RT_OPT_CHUNK_PIXELS, 4096)
    r.ray_trace()
    assert np.array_equal(img_a, r.get_image()) and r.last_stats().primary_volume_tests == 0
    r.close()


def test_work_counters_equal_the_host_emulation(cuda_lib, hostsim_lib, robot):
    """The V/T tallies of the instrumented kernels == the same source run serially on the host (cfg1 + a fan config)."""
    for name in ("cfg1", "cfg3"):
        kw, mats, tex = common.config_table(robot["materials"])[name]
        r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
        r.ctx.set_option(api.RT_OPT_COUNT_WORK, 1)
        r.ctx.set_option(api.RT_OPT_PACKETS, 0)            # per-ray scheduling does exactly the host emulation's tests
        r.ray_trace()
        a = r.last_stats().as_dict()
        r.close()
        _, b = common.product_image(hostsim_lib, robot, kw, mats, tex)
        b = b.as_dict()
        for k in ("primary_volume_tests", "primary_triangle_tests", "shadow_volume_tests", "shadow_triangle_tests",
                  "reflection_volume_tests", "reflection_triangle_tests", "reflection_rays", "reflection_shadow_rays"):
            assert a[k] == b[k], (name, k, a[k], b[k])
        assert a["primary_volume_tests"] > 0


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_packet_and_single_ray_kernels_agree(cuda_lib, oracle, robot, name):
    """RT_OPT_PACKETS only changes how primary / shadow rays are scheduled (32-ray packets vs one state machine per
    lane): frames and ray counts are identical, and both equal the oracle's."""
    kw, mats, tex = common.config_table(robot["materials"])[name]
    out = []
    for packets in (1, 0):
        r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
        r.ctx.set_option(api.RT_OPT_PACKETS, packets)
        r.ray_trace()
        out.append((r.get_image().copy(), r.last_stats().as_dict()))
        r.close()
    assert np.array_equal(out[0][0], out[1][0])
    for k in ("primary_rays", "shadow_rays", "primary_hits", "reflection_rays", "reflection_shadow_rays"):
        assert out[0][1][k] == out[1][1][k]
    common.assert_image_close(out[0][0], common.oracle_image(oracle, robot, kw, mats, tex), what=name)


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_split_shadow_packets_never_change_a_frame(cuda_lib, oracle, robot, name, fused):
    """RT_OPT_PACKET_ROUNDS / RT_OPT_ITEM_ROUNDS split shadow packets that run out of rounds into work items (one per
    unvisited cell) that other warps trace for the same rays; answers are merged and the pixels stored by k_shade_finish.
    With budgets of 1..20 rounds nearly every packet is split, items are split again, and on the larger frames the item
    regions overflow (finish-in-place path): the frames must equal the unlimited-packet frame bit for bit.
    fused = 0 (the default): six item passes and a finish kernel per stage; fused = 1 (RT_OPT_FUSED_ITEMS): the packet kernel's own
    warps consume the items through ticket queues and the last item of a record stores its pixels."""
    kw, mats, tex = common.config_table(robot["materials"])[name]
    frames = {}
    for rounds, item_rounds in ((0, 64), (1, 1), (3, 2), (20, 5), (2, 64)):
        r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
        r.ctx.set_option(api.RT_OPT_FUSED_ITEMS, fused)
        r.ctx.set_option(api.RT_OPT_PACKET_ROUNDS, rounds)
        r.ctx.set_option(api.RT_OPT_PRIMARY_ROUNDS, rounds)                 # primary packets: closest hit merged with atomicMin
        r.ctx.set_option(api.RT_OPT_ITEM_ROUNDS, item_rounds)
        r.ray_trace()
        frames[(rounds, item_rounds)] = (r.get_image().copy(), r.last_stats().as_dict())
        r.close()
    base = frames[(0, 64)]
    for key, (img, st) in frames.items():
        if key == (0, 64):
            continue
        assert np.array_equal(img, base[0]), key
        for k in ("primary_rays", "shadow_rays", "primary_hits", "reflection_rays", "reflection_shadow_rays"):
            assert st[k] == base[1][k]
        if fused:
            assert st["kernel_launches"] == base[1]["kernel_launches"]         # one launch per stage whatever is split
        else:
            assert st["kernel_launches"] >= base[1]["kernel_launches"] + 14     # (6 item passes + finish) for primary and for shadow packets, per chunk
    common.assert_image_close(frames[(1, 1)][0], common.oracle_image(oracle, robot, kw, mats, tex), what=name + " through split packets")


def test_split_primary_packets_keep_ids_and_ties(cuda_lib, oracle):
    """Split primary packets on a scene made of 60 coincident copies of a few triangles (every hit is a tie on t between
    copies in different leaves' order): the merged closest hit must be the lowest original index, as without splitting."""
    tri = common.triangle_soup(40, 3, size=0.6, center=(0, 0, -4), spread=0.8)
    soup = np.tile(tri, (60, 1)).astype(np.float32)
    # copy k of every triangle has material k, in a random order of the copies: which copy wins a tie is visible
    perm = np.random.default_rng(5).permutation(60)
    mats = rt.precompute_materials([dict(scenes.DEFAULT_SPHERE_MATERIAL, diffuse=(k / 60.0, 1 - k / 60.0, 0.5)) for k in range(60)])
    scene = dict(xyz9=soup, uv6=None, mat=np.repeat(perm, 40).astype(np.int32))
    kw = dict(image_width=128, image_height=72, compute_shadows=0)
    out = []
    for rounds, fused in ((0, 1), (1, 1), (2, 1), (1, 0), (2, 0)):
        r = common.product_renderer(cuda_lib, scene, kw, mats, {})
        r.ctx.set_option(api.RT_OPT_FUSED_ITEMS, fused)
        r.ctx.set_option(api.RT_OPT_PRIMARY_ROUNDS, rounds)
        r.ctx.set_option(api.RT_OPT_ITEM_ROUNDS, 1)
        r.ray_trace()
        out.append((r.get_image().copy(), r.last_stats().primary_hits))
        r.close()
    assert out[0][1] > 200
    for img, hits in out[1:]:
        assert np.array_equal(img, out[0][0]) and hits == out[0][1]
    common.assert_image_close(out[1][0], common.oracle_image(oracle, scene, kw, mats, {}), what="split primary packets, coincident triangles")


def test_screen_cull_never_changes_a_frame(cuda_lib, oracle, robot):
    """RT_OPT_SCREEN_CULL writes the primary packets outside the screen-space bound of the root box as misses without
    tracing them.  Cameras in front of, beside, inside and behind the scene, with and without a skysphere: the frame
    equals the one traced without the cull bit for bit, and the oracle's within tolerance."""
    table = common.config_table(robot["materials"])
    c, s = np.cos(0.9), np.sin(0.9)
    cams = [None,
            np.float32([[c, 0, s, 1.5], [0, 1, 0, -0.5], [-s, 0, c, -1.0], [0, 0, 0, 1]]),       # from the side, scene partly off-screen
            np.float32([[1, 0, 0, 0.0], [0, 1, 0, -1.5], [0, 0, 1, -3.9], [0, 0, 0, 1]]),        # inside the scene's box
            np.float32([[-1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, -1, 2.0], [0, 0, 0, 1]]),        # looking away: nothing on screen
            np.float32([[1, 0, 0, 3.0], [0, 1, 0, 2.0], [0, 0, 1, 4.0], [0, 0, 0, 1]])]          # far away: a small bound
    for name in ("cfg1", "cfg3", "cfg3_skybox"):
        kw, mats, tex = table[name]
        kw = dict(kw, image_width=160, image_height=90)
        for i, cam in enumerate(cams):
            frames = []
            for cull in (1, 0):
                r = common.product_renderer(cuda_lib, robot, kw, mats, tex, cam=cam)
                r.ctx.set_option(api.RT_OPT_SCREEN_CULL, cull)
                r.ray_trace()
                frames.append((r.get_image().copy(), r.last_stats().as_dict()))
                r.close()
            assert np.array_equal(frames[0][0], frames[1][0]), (name, i)
            for k in ("primary_rays", "primary_hits", "shadow_rays", "reflection_rays"):
                assert frames[0][1][k] == frames[1][1][k], (name, i, k)
            common.assert_image_close(frames[0][0], common.oracle_image(oracle, robot, kw, mats, tex, cam=cam), what=f"{name} camera {i}")


def test_hair_scene_vs_oracle(cuda_lib, oracle):
    """BASELINE.json configs[4] at reduced size (the oracle must finish in seconds): 256 000 thin double-sided strand
    triangles -- a deep tree and incoherent packets, many of them split -- with hard shadows, 4 spp; the frame against the
    oracle, closest hits of a ray batch bit-exact, and the frame unchanged by the scheduling knobs."""
    xyz9, uv6, mat = scenes.hair_ball(n_strands=4000, segments=16, width=4.0e-3)
    scene = dict(xyz9=xyz9, uv6=uv6, mat=mat)
    mats = rt.precompute_materials([scenes.DEFAULT_SPHERE_MATERIAL])
    kw = dict(image_width=320, image_height=180, enable_ssaa=1, ssaa_factor=2, compute_shadows=1)
    r = common.product_renderer(cuda_lib, scene, kw, mats, {})
    assert r.bvh_info["triangles"] == 256000
    r.ray_trace()
    img, st = r.get_image().copy(), r.last_stats().as_dict()
    want = common.oracle_image(oracle, scene, kw, mats, {})
    common.assert_image_close(img, want, what="hair scene")
    assert (img == want).mean() >= 0.999 and st["primary_hits"] > 20000 and st["shadow_rays"] == st["primary_hits"]
    o, d = common.random_rays(20000, 23, (-1, -1, -4), (1, 1, -2))
    got, ref = r.ctx.intersect(o, d), oracle.bvh(xyz9, 12, 40).intersect(o, d)
    for g, w in zip(got, ref):
        assert np.array_equal(g, w)
    for opts in ({api.RT_OPT_PACKETS: 0}, {api.RT_OPT_PACKET_ROUNDS: 8, api.RT_OPT_PRIMARY_ROUNDS: 8, api.RT_OPT_ITEM_ROUNDS: 4},
                 {api.RT_OPT_PACKET_ROUNDS: 8, api.RT_OPT_PRIMARY_ROUNDS: 8, api.RT_OPT_ITEM_ROUNDS: 4, api.RT_OPT_FUSED_ITEMS: 1},
                 {api.RT_OPT_SCREEN_CULL: 0}, {api.RT_OPT_SHADOW_SORT: 0}):
        for k, v in opts.items():
            r.ctx.set_option(k, v)
        r.ray_trace()
        assert np.array_equal(r.get_image(), img), opts
        r.ctx.set_option(api.RT_OPT_PACKETS, 1)
        r.ctx.set_option(api.RT_OPT_PACKET_ROUNDS, -256); r.ctx.set_option(api.RT_OPT_PRIMARY_ROUNDS, -256); r.ctx.set_option(api.RT_OPT_ITEM_ROUNDS, -16)
        r.ctx.set_option(api.RT_OPT_SCREEN_CULL, 1); r.ctx.set_option(api.RT_OPT_SHADOW_SORT, 2)
        r.ctx.set_option(api.RT_OPT_FUSED_ITEMS, 0)
    r.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_top_table_never_changes_a_frame(cuda_lib, robot, name):
    """RT_OPT_TOP_TABLE: the packet kernels take the first tree levels from a per-CTA shared-memory copy that one
    cp.async.bulk loads (instead of fetching those cells): same frame, same counts -- also on a tree so small that the
    whole of it is in the table (depth 1) and on one whose root is a leaf."""
    kw, mats, tex = common.config_table(robot["materials"])[name]
    out = []
    for top, depth, leaf in ((0, 12, 40), (1, 12, 40), (1, 1, 40), (1, 0, 5)):
        r = common.product_renderer(cuda_lib, robot, dict(kw, bvh_max_depth=depth, bvh_leaf_object_count=leaf), mats, tex)
        r.ctx.set_option(api.RT_OPT_TOP_TABLE, top)
        r.ray_trace()
        out.append((r.get_image().copy(), r.last_stats().as_dict()))
        r.close()
    for img, st in out[1:]:
        assert np.array_equal(img, out[0][0])
        for k in ("primary_rays", "shadow_rays", "primary_hits", "reflection_rays", "reflection_shadow_rays"):
            assert st[k] == out[0][1][k]


@pytest.mark.parametrize("light", [(3.0, 3.0, 2.0), (0.0, 0.4, -4.0), (40.0, 60.0, 10.0)])
def test_light_space_queue_order_never_changes_a_frame(cuda_lib, robot, light):
    """RT_OPT_SHADOW_SORT reorders the hit queue by the direction of the hits from the light before the shadow packets are
    formed: same frame, same counts -- with the light outside the scene's bounding sphere (gnomonic map), inside it
    (octahedral map) and far away (a narrow cone)."""
    kw, mats, tex = common.config_table(robot["materials"])["cfg2"]
    out = []
    for sort in (1, 0):
        r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
        r.ctx.set_light(light)
        r.ctx.set_option(api.RT_OPT_SHADOW_SORT, sort)
        for _ in range(2):
            r.ray_trace()
        out.append((r.get_image().copy(), r.last_stats().as_dict()))
        r.close()
    assert np.array_equal(out[0][0], out[1][0])
    for k in ("primary_rays", "shadow_rays", "primary_hits"):
        assert out[0][1][k] == out[1][1][k]
    assert out[0][1]["kernel_launches"] > out[1][1]["kernel_launches"]


@pytest.mark.parametrize("name", ["ssao_ssaa2", "ssao_normal_mapped", "ssao_leftover_columns", "ssao_debug_shading"])
def test_ssao_vs_oracle(cuda_lib, oracle, robot, golden_ssao, name):
    """Renderer::post_process_ssao_SIMD (renderer.cpp:1229-1434) on the GPU: G-buffers from the shade stage, per-pixel
    occlusion counts, 7x7 blur applied before the SSAA resolve -- against the oracle's restatement with the same per-pixel
    random stream (whose arithmetic is pinned bit-exactly to the compiled reference, tests/test_oracle_vs_reference.py)
    and against the frame stored with the reference's golden.  Also a frame width that is not a multiple of 8 (the
    reference's scalar loop for the left-over columns) and a debug shading mode (geometric normals in the G-buffer)."""
    table = common.ssao_table(robot["materials"])
    if name == "ssao_leftover_columns":
        kw, mats, tex = table["ssao_ssaa2"]
        kw = dict(kw, image_width=99, image_height=57, enable_ssaa=0)
    elif name == "ssao_debug_shading":
        kw, mats, tex = table["ssao_ssaa2"]
        kw = dict(kw, shading_method=api.RT_ABS_NORMALS_SHADING)
    else:
        kw, mats, tex = table[name]
    r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
    r.ray_trace()
    r.post_process()
    img = r.get_image().copy()
    want, _ = common.oracle_renderer(oracle, robot, kw, mats, tex).render_ssao()
    common.assert_image_close(img, want, what=name)
    assert (img == want).mean() >= 0.999
    if name in table:
        assert (img == golden_ssao[name + "_per_pixel"]).mean() >= 0.999
        # against the reference's own frame the two differ as two samplings of the same estimator do
        assert common.image_error(img, golden_ssao[name + "_reference"])[0] < 0.5
    plain = common.product_renderer(cuda_lib, robot, dict(kw, enable_ssao=0), mats, tex)
    plain.ray_trace()
    assert (plain.get_image() != img).mean() > 0.05
    plain.close()
    # the single-ray kernels fill the same G-buffers; the stream does not depend on the schedule
    r.ctx.set_option(api.RT_OPT_PACKETS, 0)
    r.ray_trace()
    assert np.array_equal(r.get_image(), img)
    r.ctx.set_option(api.RT_OPT_PACKETS, 1)
    # a tile shard cannot run the pass: its samples read the z-buffer of neighbouring tiles
    import torch
    frame = torch.zeros((kw["image_height"], kw["image_width"]), dtype=torch.int32, device="cuda")
    with pytest.raises(api.RtError) as e:
        r.ctx.render_device(r.render_settings(), frame.data_ptr(), 32, 2, 0)
    assert e.value.code == api.RT_ERR_UNSUPPORTED
    r.close()


STAT_KEYS = ("triangles", "nodes", "interior", "leaves", "empty_leaves", "max_depth_reached", "max_leaf_size")


@pytest.mark.parametrize("params", [(12, 40), (10, 8), (3, 2), (0, 5), (20, 1)])
@pytest.mark.parametrize("leaf_split", [8, 0])
def test_device_build_is_the_host_build(cuda_lib, oracle, robot, golden_rays, params, leaf_split):
    """RT_OPT_DEVICE_BUILD: the octree built on the GPU (octree_device.cuh) against the host builder (octree_build.cpp)
    and the oracle's BVH::BVH restatement: the statistics of the reference-shaped tree, closest hits (ids, t, u, v
    bit-exact), a frame, and the tallies of the single-ray traversal (same cells, same leaves in the same order, same
    refinement groups below oversized leaves)."""
    depth, leaf = params
    kw, mats, tex = common.config_table(robot["materials"])["cfg1"]
    kw = dict(kw, bvh_max_depth=depth, bvh_leaf_object_count=leaf)
    o, d = golden_rays["o"], golden_rays["d"]
    out = []
    for device in (1, 0):
        r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
        r.ctx.set_option(api.RT_OPT_DEVICE_BUILD, device)
        r.ctx.set_option(api.RT_OPT_LEAF_SPLIT, leaf_split)
        r.reconstruct_bvh_new()
        info = dict(r.bvh_info)
        hits = r.ctx.intersect(o, d)
        r.ctx.set_option(api.RT_OPT_COUNT_WORK, 1)
        r.ctx.set_option(api.RT_OPT_PACKETS, 0)
        r.ray_trace()
        out.append((info, hits, r.get_image().copy(), r.last_stats().as_dict()))
        r.close()
    (di, dh, dimg, dst), (hi, hh, himg, hst) = out
    for k in STAT_KEYS:
        assert di[k] == hi[k], (k, di[k], hi[k])
    want = oracle.bvh(robot["xyz9"], depth, leaf).intersect(o, d)
    for g, h, w in zip(dh, hh, want):
        assert np.array_equal(g, h) and np.array_equal(g, w)
    assert np.array_equal(dimg, himg)
    assert di["child_records"] >= hi["child_records"]             # same records; the device layout pads every block to an even size
    for k in ("primary_volume_tests", "primary_triangle_tests", "shadow_volume_tests", "shadow_triangle_tests"):
        assert dst[k] == hst[k], (k, dst[k], hst[k])


def test_device_build_edge_cases(cuda_lib, oracle):
    """One triangle, 60 coincident triangles (a chain of single-child cells, then a leaf that cannot be refined), a
    degenerate triangle among others, and all centroids on one plane: device build == host build."""
    rng = np.random.default_rng(5)
    one = np.float32([[0, 0, -3, 1, 0, -3, 0, 1, -3]])
    many = rng.uniform(-1, 1, (500, 9)).astype(np.float32) + np.float32([0, 0, -3] * 3)
    flat = many.copy(); flat[:, 2::3] = -3.0
    scenes_ = {"one": one, "coincident": np.repeat(one, 60, 0), "degenerate": np.concatenate([many, np.float32([[0.2, 0.2, -3] * 3])]), "flat": flat}
    o, d = common.random_rays(4000, 3, (-1, -1, 0), (1, 1, 0.5))
    d = (np.float32([0, 0, -3]) + rng.uniform(-1, 1, (4000, 3)).astype(np.float32) - o).astype(np.float32)
    for name, xyz9 in scenes_.items():
        res = []
        for device in (1, 0):
            ctx = api.Context(0, cuda_lib)
            ctx.set_option(api.RT_OPT_DEVICE_BUILD, device)
            ctx.set_triangles(xyz9, None, None)
            info = ctx.build_bvh(12, 4)
            res.append((info, ctx.intersect(o, d)))
            ctx.close()
        for k in STAT_KEYS:
            assert res[0][0][k] == res[1][0][k], (name, k)
        # (overlapping coplanar triangles tie on t across leaves; the reference keeps the first in traversal order there, this
        # library the lowest index -- "flat" is compared between the two builders only)
        want = res[1][1] if name == "flat" else oracle.bvh(xyz9, 12, 4).intersect(o, d)
        for g, h, w in zip(res[0][1], res[1][1], want):
            assert np.array_equal(g, h) and np.array_equal(g, w), name


def test_device_transform_and_rebuild(cuda_lib, oracle, robot):
    """Renderer::set_object_transform (renderer.cpp:214-224) with the triangles resident on the device: two successive
    transforms, each followed by a device rebuild, give the frame of a host transform + host build; switching to the host
    builder afterwards sees the transformed vertices."""
    kw = dict(image_width=160, image_height=90, compute_shadows=1)
    m1, m2 = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    m1[0, 3], m1[2, 3] = 0.5, -1.0
    c, s_ = np.float32(np.cos(0.3)), np.float32(np.sin(0.3))
    m2[0, 0], m2[0, 2], m2[2, 0], m2[2, 2], m2[1, 3] = c, s_, -s_, c, 0.25
    imgs, infos = [], []
    for device in (1, 0):
        r = common.product_renderer(cuda_lib, robot, kw, robot["materials"], {})
        r.ctx.set_option(api.RT_OPT_DEVICE_BUILD, device)
        r.reconstruct_bvh_new()
        r.set_object_transform(m1)
        r.set_object_transform(m2)
        r.ray_trace()
        imgs.append(r.get_image().copy()); infos.append(dict(r.bvh_info))
        if device:
            r.ctx.set_option(api.RT_OPT_DEVICE_BUILD, 0)       # the host copy of the vertices catches up with the device copy
            r.reconstruct_bvh_new()
            r.ray_trace()
            assert np.array_equal(r.get_image(), imgs[0])
        r.close()
    assert np.array_equal(imgs[0], imgs[1]) and (imgs[0] != imgs[0][0, 0]).any()
    for k in STAT_KEYS:
        assert infos[0][k] == infos[1][k], k


def test_two_lanes_never_change_a_frame(cuda_lib, robot):
    """RT_OPT_LANES runs two wavefront chunks at a time on two streams with their own queues: same frame, same counts."""
    kw, mats, tex = common.config_table(robot["materials"])["cfg3"]
    out = []
    for lanes in (1, 0):
        r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
        r.ctx.set_option(api.RT_OPT_LANES, lanes)
        for _ in range(2):
            r.ray_trace()
        out.append((r.get_image().copy(), r.last_stats().as_dict()))
        r.close()
    assert np.array_equal(out[0][0], out[1][0])
    for k in ("primary_rays", "shadow_rays", "primary_hits", "reflection_rays", "reflection_shadow_rays"):
        assert out[0][1][k] == out[1][1][k]
    assert out[0][1]["kernel_launches"] > out[1][1]["kernel_launches"]


def test_analytic_shapes_gate_order_and_refusals(cuda_lib, oracle, robot):
    """trace_ray takes the CLOSEST of the BVH hit and the analytic shapes and only then applies min_t (renderer.cpp:
    1029-1040): a plane 0.05 in front of the camera hides the robot AND is itself rejected, so the frame is background; a
    far plane behind the robot changes only the background pixels.  Texture mapping and the barycentric / AO debug modes
    are refused while shapes exist (the reference reads a stale HitInfo::triangle there)."""
    mats = list(robot["materials"])
    kw = dict(image_width=96, image_height=54, compute_shadows=1)
    bg = 0xff000000 | (135 << 16) | (206 << 8) | 235
    near = {"shapes": [("plane", (0, 0, -0.05), (0, 0, 1), 0)]}
    img, st = common.product_image(cuda_lib, robot, kw, mats, near)
    assert (img == bg).all() and st.primary_hits == 0
    assert np.array_equal(img, common.oracle_image(oracle, robot, kw, mats, near))
    far = {"shapes": [("plane", (0, 0, -9), (0, 0, 1), 1), ("sphere", (0, 0, 0), 0.08, 0)]}     # the sphere around the camera: t = 0.08, rejected too
    img2, st2 = common.product_image(cuda_lib, robot, kw, mats, far)
    want2 = common.oracle_image(oracle, robot, kw, mats, far)
    common.assert_image_close(img2, want2, what="far plane + rejected sphere")
    plain, _ = common.product_image(cuda_lib, robot, kw, mats, {})
    assert (img2 == bg).all() and (plain != bg).any()        # the rejected sphere is the closest hit of every ray
    r = common.product_renderer(cuda_lib, robot, dict(kw, enable_ao_mapping=1), mats, {0: scenes.noise_texture((32, 32), 2), "shapes": far["shapes"][:1]})
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
    common.assert_image_close(r.get_image(), want3, what="normals of a plane")
    r.ctx.clear_analytic_shapes()
    r.render_settings().shading_method = api.RT_SHADING
    r.ray_trace()
    assert np.array_equal(r.get_image(), plain)
    r.close()


def test_cpp_adapter_example(cuda_lib, tmp_path):
    """include/rtb200_renderer.hpp (the reference's method names over the C ABI) compiles and renders."""
    exe = tmp_path / "adapter_example"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "adapter_example.cpp"),
                    "-o", str(exe), "-L", str(ROOT / "raytracercpp_b200"), "-lrtb200", f"-Wl,-rpath,{ROOT / 'raytracercpp_b200'}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "adapter ok" in out


# ---- BASELINE.json configs[0..2] and [4] at their FULL sizes ----------------------------------------------------------
def check_fullsize_frame(lib, oracle, robot, golden_fullsize, name):
    kw, mats, tex = common.fullsize_table(robot["materials"])[name]
    img, stats = common.product_image(lib, robot, kw, mats, tex)
    assert img.shape == (kw["image_height"], kw["image_width"])
    want = common.oracle_image(oracle, robot, kw, mats, tex)                     # live oracle, the whole frame
    common.assert_image_close(img, want, what=name + " vs oracle")
    assert (img == want).mean() >= 0.999
    rows = golden_fullsize[name + "_rows"]                                       # committed rows of the compiled reference's frame
    common.assert_image_close(img[::common.FULL_ROW_STEP], rows, what=name + " vs reference rows")
    assert (img[::common.FULL_ROW_STEP] == rows).mean() >= 0.999
    f = kw.get("ssaa_factor", 1) if kw.get("enable_ssaa") else 1
    assert stats.primary_rays == kw["image_width"] * kw["image_height"] * f * f
    assert stats.primary_hits == int(golden_fullsize[name + "_hits"]) and stats.shadow_rays == stats.primary_hits
    cnt = common.oracle_renderer(oracle, robot, kw, mats, tex).count_rows()
    assert stats.reflection_rays == cnt["reflection_rays"] and stats.reflection_shadow_rays == cnt["reflection_shadow_rays"]
    return stats


@pytest.mark.parametrize("name", ["cfg1_full", "cfg2_full", "cfg3_full"])
def test_fullsize_frames_vs_oracle_and_reference_rows(cuda_lib, oracle, robot, golden_fullsize, name):
    """cfg1 at 1280x720, cfg2 at 1280x720 x ssaa 2 with 2048^2 u8 maps, cfg3 at 1920x1080 with the seeded 16-ray fan
    (renderer.cpp:1068-1116, :283-338): the whole frame against the live oracle and against the committed rows of the
    compiled reference's frame; hit / shadow / fan ray counts exact."""
    stats = check_fullsize_frame(cuda_lib, oracle, robot, golden_fullsize, name)
    if name == "cfg3_full":
        assert stats.reflection_rays > 1_000_000


def test_hair_fullsize_band(cuda_lib, oracle, golden_fullsize):
    """BASELINE.json configs[4] at its size: 1 000 000 thin strand triangles, 3840x2160, hard shadows.  The whole frame is
    rendered; a band of rows is compared with the live oracle and with the committed rows of the compiled reference
    (incoherent packets, deep tree: the worst case of the packet traversal)."""
    xyz9, uv6, mat = scenes.hair_ball(**common.HAIR_FULL)
    scene = dict(xyz9=xyz9, uv6=uv6, mat=mat)
    mats = rt.precompute_materials([scenes.DEFAULT_SPHERE_MATERIAL])
    r = common.product_renderer(cuda_lib, scene, common.HAIR_KW, mats, {})
    assert r.bvh_info["triangles"] == 1_000_000
    r.ray_trace()
    img, st = r.get_image().copy(), r.last_stats()
    assert st.primary_rays == 3840 * 2160 and st.shadow_rays == st.primary_hits and st.primary_hits > 500_000
    b0, b1, bs = common.HAIR_BAND
    band = img[b0:b1:bs]
    rows = golden_fullsize["cfg5_band_rows"]
    common.assert_image_close(band, rows, what="hair band vs reference rows")
    assert (band == rows).mean() >= 0.999
    orc = common.oracle_renderer(oracle, scene, common.HAIR_KW, mats, {})
    sup, _ = orc.trace_rows(row_begin=b0, row_end=b1, row_step=bs)
    assert np.array_equal(sup[b0:b1:bs], rows)                                   # the oracle is bit-exact on the band
    # closest hits of an incoherent ray batch, bit-exact against the oracle's octree
    o, d = common.random_rays(50_000, 29, (-1, -1, -4), (1, 1, -2))
    for g, w in zip(r.ctx.intersect(o, d), oracle.bvh(xyz9, 12, 40).intersect(o, d)):
        assert np.array_equal(g, w)
    # scheduling knobs leave the 4K frame bit-identical
    for opt, val, back in ((api.RT_OPT_PACKETS, 0, 1), (api.RT_OPT_SCREEN_CULL, 0, 1), (api.RT_OPT_FUSED_ITEMS, 1, 0), (api.RT_OPT_LANES, 4, 1), (api.RT_OPT_TOP_TABLE, 1, 0),
                           (api.RT_OPT_SHADOW_SORT, 0, 2)):
        r.ctx.set_option(opt, val)
        r.ray_trace()
        assert np.array_equal(r.get_image(), img), opt
        r.ctx.set_option(opt, back)
    r.close()


# ---- full BASELINE.json size: 10 M triangles, 3840x2160, 16 spp, shadows --------------------------------------------
@pytest.fixture(scope="module")
def big_sphere():
    return scenes.displaced_sphere(*scenes.sphere_grid_for(10_000_000))


def test_full_size_properties(cuda_lib, oracle, big_sphere):
    import torch
    xyz9, uv6, mat = big_sphere
    mats = rt.precompute_materials([scenes.DEFAULT_SPHERE_MATERIAL])
    scene = dict(xyz9=xyz9, uv6=uv6, mat=mat)
    kw = dict(image_width=3840, image_height=2160, enable_ssaa=1, ssaa_factor=4, compute_shadows=1)
    r = common.product_renderer(cuda_lib, scene, kw, mats, {})
    info = r.bvh_info
    assert info["triangles"] == len(xyz9) and info["max_depth_reached"] <= 12
    # (1) closest hit does not depend on the tree parameters: (12,40) vs (16,8) agree on a ray sample; and both agree with
    #     brute-force-equivalent reasoning: every hit triangle really is hit at that t (re-intersect the single triangle)
    o, d = common.random_rays(400_000, 17, (-2, -2, -5), (2, 2, -1))
    a = r.ctx.intersect(o, d)
    other = api.Context(0, cuda_lib)
    other.set_triangles(xyz9)
    other.build_bvh(16, 8)
    b = other.intersect(o, d)
    other.close()
    assert (a[0] == b[0]).mean() >= 0.999
    same = a[0] == b[0]
    assert np.array_equal(a[1][same], b[1][same])
    hit = np.flatnonzero(a[0] >= 0)[:2000]
    for i in hit[:200]:
        ok, t, u, v = oracle.triangle_intersect(xyz9[a[0][i]], o[i], d[i])
        assert ok and t == a[1][i] and u == a[2][i] and v == a[3][i]
    assert len(hit) > 1000
    # (2) the frame: ray counts, shards tile it bit-exactly, every pixel written, background where nothing is hit
    r.ray_trace()
    full, st = r.get_image(), r.last_stats()
    assert st.primary_rays == 3840 * 2160 * 16 and 0 < st.primary_hits < st.primary_rays and st.shadow_rays == st.primary_hits
    s = r.render_settings()
    frame = torch.zeros(full.shape, dtype=torch.int32, device="cuda")
    total_hits = 0
    for rem in range(2):
        stats = r.ctx.render_device(s, frame.data_ptr(), 64, 2, rem)
        total_hits += stats.primary_hits
    torch.cuda.synchronize()
    assert np.array_equal(frame.cpu().numpy().view(np.uint32), full) and total_hits == st.primary_hits
    bg = 0xff000000 | (135 << 16) | (206 << 8) | 235
    assert full[0, 0] == bg and full[1080, 1920] != bg
    # (3) idempotence: same frame twice; and the scheduling knobs at full size (item passes as separate launches, 4 chunks in flight)
    r.ray_trace()
    assert np.array_equal(full, r.get_image())
    for opt, val, back in ((api.RT_OPT_FUSED_ITEMS, 1, 0), (api.RT_OPT_LANES, 4, 1), (api.RT_OPT_TOP_TABLE, 1, 0), (api.RT_OPT_SHADOW_SORT, 0, 2)):
        r.ctx.set_option(opt, val)
        r.ray_trace()
        assert np.array_equal(full, r.get_image()), opt
        r.ctx.set_option(opt, back)
    # (3b) the tree was built on the device (the default); the host builder gives the same statistics, hits and frame
    dev_info = dict(info)
    r.reconstruct_bvh_new()                                    # once more with the builder's buffers allocated: what a transform + rebuild costs
    assert r.bvh_info["build_ms"] < 50.0 and r.bvh_info["upload_ms"] < 1.0, (r.bvh_info["build_ms"], r.bvh_info["upload_ms"])
    for k in STAT_KEYS + ("child_records",):
        assert r.bvh_info[k] == dev_info[k]
    r.ctx.set_option(api.RT_OPT_DEVICE_BUILD, 0)
    r.reconstruct_bvh_new()
    for k in STAT_KEYS:
        assert r.bvh_info[k] == dev_info[k], (k, r.bvh_info[k], dev_info[k])
    for g, w in zip(r.ctx.intersect(o, d), a):
        assert np.array_equal(g, w)
    r.ray_trace()
    assert np.array_equal(full, r.get_image())
    r.ctx.set_option(api.RT_OPT_DEVICE_BUILD, 1)
    # (4) a band of the frame against the oracle (the oracle needs ~10 s for these rows at this size)
    orc = common.oracle_renderer(oracle, scene, kw, mats, {})
    rows = list(range(4000, 4640, 64))
    sup, _ = orc.trace_rows(row_begin=rows[0], row_end=rows[-1] + 1, row_step=64)
    no = api.default_settings(cuda_lib, **dict(kw, enable_ssaa=0, image_width=3840 * 4, image_height=2160 * 4))
    big = np.empty((2160 * 4, 3840 * 4), np.uint32)
    r.ctx.set_camera(r.ctx.perspective_inverse(80.0, np.float32(3840 * 4) / np.float32(2160 * 4)), np.eye(4), (0, 0, 0))
    r.ctx.render(no, big)
    common.assert_image_close(big[rows], sup[rows], what="10M-triangle band vs oracle")
    assert (big[rows] == sup[rows]).mean() >= 0.999
    r.close()


@pytest.mark.parametrize("split", [0, 1, 3, 16])
def test_leaf_refinement_never_changes_a_result(cuda_lib, oracle, robot, golden_rays, split):
    """RT_OPT_LEAF_SPLIT only reshapes the device-side hierarchy below the reference's leaves: ids, t, u, v and frames
    stay bit-identical, including with the reference's own leaves (split = 0)."""
    ctx = api.Context(0, cuda_lib)
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
    r = common.product_renderer(cuda_lib, robot, dict(kw, image_width=96, image_height=54), mats, tex)
    r.ctx.set_option(api.RT_OPT_LEAF_SPLIT, split)
    r.reconstruct_bvh_new()
    r.ray_trace()
    common.assert_image_close(r.get_image(), common.oracle_image(oracle, robot, dict(kw, image_width=96, image_height=54), mats, tex), what="split")
    r.close()


# ---- hybrid path: Renderer::raster_trace (renderer.cpp:869-1006), SURVEY.md section 8(f)4 -------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["r_cfg1", "r_cfg2", "r_inside", "r_inside_noclip", "r_mirror", "r_big", "r_ssao", "r_rough",
                                  "r_debug1", "r_debug2", "r_debug3", "r_debug4"])
def test_raster_trace_vs_oracle(cuda_lib, oracle, robot, golden_raster, name):
    """hybrid_rasterization_tracing on the GPU (csrc/raster.cuh: depth pass with 64-bit z-keys, emit pass, shade pass) through
    the C ABI, against the oracle's restatement and against the frames of the compiled reference run on one thread.
    Coverage, depth order and the shaded hit are integer / bit-exact work: every pixel must have the oracle's WINNER; only
    powf of the specular term can move a channel by one unit."""
    scene, kw, mats, tex, cam = common.raster_table(robot)[name]
    img, st = common.product_image(cuda_lib, scene, kw, mats, tex, cam=cam)
    want = common.oracle_image(oracle, scene, kw, mats, tex, cam=cam)
    common.assert_image_close(img, want, what=name)
    assert (img == want).mean() >= 0.999
    if name in common.RASTER_PINNED:
        ref = golden_raster[name + ("_per_pixel" if kw.get("enable_ssao") else "_reference")]
        common.assert_image_close(img, ref, what=name + " vs the compiled reference")
        assert (img == ref).mean() >= 0.999
    if kw.get("shading_method", 0) == 0:
        assert st.primary_rays > 0 and st.shadow_rays == st.primary_hits and st.kernel_launches >= 3
    if kw.get("shading_method", 0) in (1, 2, 3):
        assert np.array_equal(img, want)                   # no libm call on these paths: bit-exact


@pytest.mark.gpu
def test_raster_trace_schedule_independent(cuda_lib, oracle, robot):
    """The frame does not depend on how the pieces are scheduled: a unit list that is too short at first (the depth pass is
    repeated with a longer one), the host-built tree (another leaf order), a second frame from the same context; and
    ray_trace() after raster_trace() is the ray-traced frame again."""
    scene, kw, mats, tex, cam = common.raster_table(robot)["r_big"]
    base, _ = common.product_image(cuda_lib, scene, kw, mats, tex, cam=cam)
    r = common.product_renderer(cuda_lib, scene, kw, mats, tex, cam=cam)
    r.ctx.set_option(api.RT_OPT_RASTER_UNITS, 3)
    r.raster_trace()
    assert np.array_equal(r.get_image(), base)
    r.raster_trace()
    assert np.array_equal(r.get_image(), base)
    r.ctx.set_option(api.RT_OPT_DEVICE_BUILD, 0)
    r.reconstruct_bvh_new()
    r.raster_trace()
    assert np.array_equal(r.get_image(), base)
    r.ray_trace()
    traced = r.get_image().copy()
    r.close()
    plain, _ = common.product_image(cuda_lib, scene, dict(kw, hybrid_rasterization_tracing=0), mats, tex, cam=cam)
    assert np.array_equal(traced, plain)
    # the rasterised and the ray-traced frame show the same scene: they differ at silhouettes and where the two
    # paths' hit points differ by rounding, not wholesale
    assert (traced == base).mean() > 0.8
    # a tile shard cannot own a z-buffer of its own
    import torch
    frame = torch.zeros((kw["image_height"], kw["image_width"]), dtype=torch.int32, device="cuda")
    r2 = common.product_renderer(cuda_lib, scene, kw, mats, tex, cam=cam)
    with pytest.raises(api.RtError) as e:
        r2.ctx.render_device(r2.render_settings(), frame.data_ptr(), 32, 2, 0)
    assert e.value.code == api.RT_ERR_UNSUPPORTED
    r2.close()


@pytest.mark.gpu
def test_raster_trace_full_size(cuda_lib, oracle, robot):
    """BASELINE configs[0] and [1] at their sizes through the hybrid path: cfg1 1280x720, cfg2 1280x720 x ssaa 2 with the
    2048^2 u8 maps, against the oracle live."""
    table = common.fullsize_table(robot["materials"])
    for name in ("cfg1_full", "cfg2_full"):
        kw, mats, tex = table[name]
        kw = dict(kw, hybrid_rasterization_tracing=1)
        img, st = common.product_image(cuda_lib, robot, kw, mats, tex)
        want = common.oracle_image(oracle, robot, kw, mats, tex)
        common.assert_image_close(img, want, what=name + " raster")
        assert (img == want).mean() >= 0.999
        assert st.primary_hits > 100_000


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["cfg3", "samples5", "samples20", "samples1", "no_roughness_map", "shapes"])
def test_fan_lanes_equal_sequential_fans(cuda_lib, oracle, robot, variant):
    """k_reflect_fan (one lane per fan ray: the reference's sequential stream offsets and stale hit record found as a fixed
    point) against k_reflect (one thread walks the fan): same frame bit for bit, same ray counts; and against the oracle."""
    table = common.config_table(robot["materials"])
    kw, mats, tex = table["cfg3"]
    if variant == "samples5":
        kw = dict(kw, rough_reflections_sample_count=5)
    elif variant == "samples20":
        kw = dict(kw, rough_reflections_sample_count=20, image_width=160, image_height=90)
    elif variant == "samples1":
        kw = dict(kw, rough_reflections_sample_count=1)
    elif variant == "no_roughness_map":
        kw = dict(kw, enable_roughness_mapping=0)
    elif variant == "shapes":
        kw, mats, tex = table["cfg3_shapes"]
        kw = dict(kw, max_recursion_depth=1, rough_reflections_sample_count=8, rng_seed=11)
    r = common.product_renderer(cuda_lib, robot, kw, mats, tex)
    r.ctx.set_option(api.RT_OPT_GRAPH, 0)
    r.ray_trace()
    lanes, st_lanes = r.get_image().copy(), r.last_stats()
    r.ctx.set_option(api.RT_OPT_FAN_LANES, 0)
    r.ray_trace()
    seq, st_seq = r.get_image().copy(), r.last_stats()
    r.close()
    assert np.array_equal(lanes, seq)
    assert st_lanes.reflection_rays == st_seq.reflection_rays > 0 and st_lanes.reflection_shadow_rays == st_seq.reflection_shadow_rays
    want = common.oracle_image(oracle, robot, kw, mats, tex)
    common.assert_image_close(lanes, want, what="fan lanes " + variant)
    assert (lanes == want).mean() >= 0.999


@pytest.mark.gpu
def test_pyramid_cull_and_frame_graphs_never_change_a_frame(cuda_lib, oracle, robot):
    """RT_OPT_PACKET_CULL (the bounding-pyramid cull of a cell's children and a leaf's triangles, primary and shadow packets)
    and RT_OPT_GRAPH (a repeated frame replayed as a CUDA graph) are pure scheduling: frames, hits and ray counts equal the
    frame without them bit for bit -- from outside the scene, from INSIDE it (the pyramids' apex inside cells, every direction
    of the frame's packets against the clip of the boxes), with the light inside the scene's box (shadow pyramids with their
    apex among the occluders), with split packets and items, on a hair ball (thin, incoherent), and with SSAA."""
    from raytracercpp_b200 import scenes
    mats = robot["materials"]
    tab = common.config_table(mats)
    hx, huv, hmat = scenes.hair_ball(n_strands=1500, segments=8)
    hair = dict(xyz9=hx, uv6=huv, mat=hmat)
    hair_mats = rt.precompute_materials([scenes.DEFAULT_SPHERE_MATERIAL])
    cases = [
        ("outside", robot, tab["cfg1"][0], mats, {}, None, common.LIGHT),
        ("inside", robot, tab["cfg2"][0], mats, tab["cfg2"][2], common.CAM_INSIDE, common.LIGHT),
        ("light_inside", robot, dict(tab["cfg1"][0], enable_ssaa=1, ssaa_factor=3), mats, {}, None, (0.1, -1.2, -3.9)),
        ("hair", hair, dict(image_width=320, image_height=180, compute_shadows=1), hair_mats, {}, None, common.LIGHT),
    ]
    for name, scene, kw, m, tex, cam, light in cases:
        r = common.product_renderer(cuda_lib, scene, kw, m, tex, cam=cam, light=light)
        r.ctx.set_option(api.RT_OPT_GRAPH, 0)
        r.ctx.set_option(api.RT_OPT_PACKET_CULL, 0)
        r.ray_trace()
        base, st0 = r.get_image().copy(), r.last_stats().as_dict()
        assert st0["primary_hits"] > 500, name
        for cull, budgets in ((1, None), (2, None), (3, None), (3, (3, 2))):
            r.ctx.set_option(api.RT_OPT_PACKET_CULL, cull)
            if budgets:
                r.ctx.set_option(api.RT_OPT_PACKET_ROUNDS, budgets[0]); r.ctx.set_option(api.RT_OPT_PRIMARY_ROUNDS, budgets[0])
                r.ctx.set_option(api.RT_OPT_ITEM_ROUNDS, budgets[1])
            r.ray_trace()
            st = r.last_stats().as_dict()
            assert np.array_equal(r.get_image(), base), (name, cull, budgets)
            for k in ("primary_rays", "shadow_rays", "primary_hits", "traced_primary_rays"):
                assert st[k] == st0[k], (name, k)
        r.ctx.set_option(api.RT_OPT_PACKET_ROUNDS, -256); r.ctx.set_option(api.RT_OPT_PRIMARY_ROUNDS, -256); r.ctx.set_option(api.RT_OPT_ITEM_ROUNDS, -16)
        # the same frame three more times with graphs on: the second is captured, the third replayed; then a moved light
        # (another key: enqueued the ordinary way) and back (the captured graph again)
        r.ctx.set_option(api.RT_OPT_GRAPH, 1)
        launches = []
        for _ in range(3):
            r.ray_trace()
            assert np.array_equal(r.get_image(), base), name
            launches.append(r.last_stats().kernel_launches)
        assert launches[0] == launches[1] == launches[2]
        r.set_light_position((light[0] + 0.5, light[1], light[2]))
        r.ray_trace()
        moved = r.get_image().copy()
        assert not np.array_equal(moved, base), name
        r.set_light_position(light)
        r.ray_trace()
        assert np.array_equal(r.get_image(), base), name
        r.close()
        common.assert_image_close(base, common.oracle_image(oracle, scene, kw, m, tex, cam=cam, light=light), what="cull test " + name)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_raster_trace_random_soups(cuda_lib, oracle, seed):
    """The hybrid path on triangle soups that surround the camera (triangles behind it, through it, across every clip plane,
    degenerate ones, coincident duplicates whose fragments tie on depth): the GPU's z-keys must pick the reference's winner
    in every pixel, with clipping on and off.  (tests/test_hostsim_parity.py runs the same soups through the host build of
    the device source, against the oracle and the compiled reference, bit for bit.)"""
    rng = np.random.default_rng(seed)
    n = 260
    c = rng.uniform(-3, 3, size=(n, 1, 3)) + np.float32([0, 0, -1.0])
    soup = (c + rng.normal(scale=rng.uniform(0.1, 2.5, size=(n, 1, 1)), size=(n, 3, 3))).reshape(n, 9).astype(np.float32)
    soup[::17, 3:9] = np.tile(soup[::17, 0:3], 2)
    dup = soup[5:45].copy()
    xyz9 = np.concatenate([soup, dup]).astype(np.float32)
    uv6 = rng.uniform(0, 1, size=(len(xyz9), 6)).astype(np.float32)
    mat = np.concatenate([np.zeros(n, np.int32), np.ones(len(dup), np.int32)])
    mats = rt.precompute_materials([dict(scenes.DEFAULT_SPHERE_MATERIAL, diffuse=(0.9, 0.2, 0.1)), dict(scenes.DEFAULT_SPHERE_MATERIAL, diffuse=(0.1, 0.3, 0.9))])
    scene = dict(xyz9=xyz9, uv6=uv6, mat=mat)
    for clipping in (1, 0):
        kw = dict(image_width=136, image_height=88, compute_shadows=1, hybrid_rasterization_tracing=1, enable_clipping=clipping,
                  enable_ssaa=int(seed == 2), ssaa_factor=2)
        img, st = common.product_image(cuda_lib, scene, kw, mats, {})
        want = common.oracle_image(oracle, scene, kw, mats, {})
        common.assert_image_close(img, want, what=f"raster soup {seed} clipping {clipping}")
        assert (img == want).mean() >= 0.999
        assert st.primary_rays > 5000
