"""Helpers shared by the parity tests: scene set-up on either side and image comparison."""
from __future__ import annotations

import numpy as np

import raytracercpp_b200 as rt
from raytracercpp_b200 import api, scenes
from raytracercpp_b200.renderer import precompute_materials
from oracle import bindings as ob

LIGHT = (3.0, 3.0, 2.0)
FOV = 80.0


def channels(argb):
    argb = np.asarray(argb)
    return np.stack([(argb >> 16) & 255, (argb >> 8) & 255, argb & 255], -1).astype(np.int32)


def image_error(a, b):
    """(mean, max) absolute per-channel difference in 8-bit units."""
    d = np.abs(channels(a) - channels(b))
    return float(d.mean()), int(d.max())


def assert_image_close(a, b, mean_tol=1.0, max_tol=4, what=""):
    """north_star tolerance: final RGB within 1/255 mean and 4/255 max per channel."""
    assert a.shape == b.shape, (what, a.shape, b.shape)
    mean, mx = image_error(a, b)
    assert mean <= mean_tol and mx <= max_tol, f"{what}: mean {mean:.4f} max {mx} (tolerance {mean_tol}/{max_tol})"


def config_table(mats):
    """The reduced-size cfg1/cfg2/cfg3 set-ups of tests/golden/make_golden.py (must stay in sync with it)."""
    tex2 = {0: scenes.noise_texture((128, 128), 2), 1: scenes.noise_texture((128, 128), 1, "rgb"),
            2: scenes.normal_map_texture((128, 128), 4), 3: scenes.noise_texture((128, 128), 3)}
    mats3 = [dict(m) for m in mats]
    mats3[0].update(reflection=0.9, roughness=0.0, specular=(0.2, 0.2, 0.2), diffuse=(0.5, 0.5, 0.5))
    mats3[1].update(reflection=0.5, roughness=0.4)
    mats3 = precompute_materials(mats3)
    tex3 = {3: tex2[3], 4: scenes.sky_texture((128, 256))}
    mats_shapes = mats3 + precompute_materials([
        dict(ambient_coeff=(1, 1, 1), diffuse=(0.2, 0.6, 0.3), specular=(0.4, 0.4, 0.4), emission=(0, 0, 0), reflection=0.0, roughness=0.0, ns=30.0),
        dict(ambient_coeff=(1, 1, 1), diffuse=(0.7, 0.7, 0.8), specular=(0.6, 0.6, 0.6), emission=(0, 0, 0), reflection=0.6, roughness=0.3, ns=80.0)])
    return {
        "cfg1": (dict(image_width=320, image_height=180, compute_shadows=1), mats, {}),
        "cfg2": (dict(image_width=160, image_height=90, compute_shadows=1, enable_ssaa=1, ssaa_factor=2, enable_ao_mapping=1,
                      enable_diffuse_mapping=1, enable_normal_mapping=1), mats, tex2),
        "cfg3": (dict(image_width=240, image_height=135, compute_shadows=1, rough_reflections_sample_count=16, max_recursion_depth=1,
                      enable_roughness_mapping=1, enable_skysphere=1, rng_seed=7), mats3, tex3),
        "cfg3_mirror5": (dict(image_width=160, image_height=90, compute_shadows=1, max_recursion_depth=5, enable_skysphere=1), mats3, tex3),
        # parallax occlusion mapping on top of cfg2's maps (renderer.cpp:518-554)
        "cfg2_pom": (dict(image_width=160, image_height=90, compute_shadows=1, enable_ao_mapping=1, enable_diffuse_mapping=1, enable_normal_mapping=1,
                          enable_displacement_mapping=1, displacement_mapping_strength=0.05, parallax_mapping_steps=16), mats,
                     {**tex2, 11: scenes.noise_texture((96, 96), 9)}),
        # analytic shapes of the GUI (QT/mainwindow.cpp:828,883,914,930): a ground plane, a mirror sphere, a matte sphere, with the
        # robot's own mirror materials; materials len(mats3) and len(mats3)+1 are appended for them
        "cfg3_shapes": (dict(image_width=200, image_height=112, compute_shadows=1, max_recursion_depth=2, enable_skysphere=1), mats_shapes,
                        {4: tex3[4], "shapes": [("plane", (0, -2, 0), (0, 1, 0), len(mats3)), ("sphere", (1.4, -1.2, -3.5), 0.5, len(mats3) + 1),
                                                ("sphere", (-1.6, -1.5, -3.0), 0.4, len(mats3))]}),
        # the GUI's default miss shader (QT/mainwindow.cpp:45-46): cube-map skybox, seen directly and in mirror reflections
        "cfg3_skybox": (dict(image_width=200, image_height=112, compute_shadows=1, max_recursion_depth=2, enable_skybox=1), mats3,
                       {5 + i: scenes.noise_texture((48 + 8 * i, 40 + 4 * i), 20 + i, "rgb") for i in range(6)}),
    }


def ssao_table(mats):
    """SSAO set-ups (renderer.cpp:1229-1434): name -> (settings, materials, textures).  The supersampled width of the
    cases that are pinned against the compiled reference is a multiple of 8: its AVX2 loop runs its last partial
    8-pixel group past the end of the row (and of the buffers at the last row)."""
    tex2 = config_table(mats)["cfg2"][2]
    return {
        "ssao_ssaa2": (dict(image_width=160, image_height=96, enable_ssaa=1, ssaa_factor=2, compute_shadows=1, enable_ssao=1, ssao_sample_count=16,
                            ssao_radius=0.5, ssao_amount=1.0, rng_seed=3), mats, {}),
        "ssao_normal_mapped": (dict(image_width=192, image_height=108, compute_shadows=1, enable_ssao=1, ssao_sample_count=24, ssao_radius=0.3,
                                    ssao_amount=0.8, enable_ao_mapping=1, enable_diffuse_mapping=1, enable_normal_mapping=1, rng_seed=5), mats, tex2),
    }


CAM_INSIDE = np.float32([[1, 0, 0, 0], [0, 1, 0, -1.0], [0, 0, 1, -2.6], [0, 0, 0, 1]])   # camera inside the robot's bounding box


def raster_table(robot):
    """hybrid_rasterization_tracing set-ups (Renderer::raster_trace, renderer.cpp:869-1006):
    name -> (scene, settings, materials, textures, camera).  The reference shades its rough-reflection fans from
    per-thread generators that cannot be reseeded per pixel, so the cases pinned against the compiled reference have no
    ROUGH reflections (mirror reflections draw no random numbers); `r_rough` is compared between oracle and product only.
    `r_big`: the robot in front of two wall triangles that cover the frame (the banded work units of raster.cuh) with the
    camera inside the scene's box, so that every clip plane cuts something."""
    mats = robot["materials"]
    tab = config_table(mats)
    tex2 = tab["cfg2"][2]
    mirror = [dict(m, roughness=0.0) for m in tab["cfg3_mirror5"][1]]
    hy = dict(hybrid_rasterization_tracing=1)
    wall = np.float32([[-30, -12, -9, 30, -12, -9, 30, 20, -7], [-30, -12, -9, 30, 20, -7, -30, 20, -7],
                       [-40, -2.5, 6, 40, -2.5, 6, 40, -2.5, -30], [-40, -2.5, 6, 40, -2.5, -30, -40, -2.5, -30]])
    wall_uv = np.float32([[0, 4, 4, 0, 0, 3], [0, 4, 0, 0, 3, 3], [0, 6, 6, 0, 0, 5], [0, 6, 0, 0, 5, 5]])
    big = dict(xyz9=np.concatenate([wall[:2], robot["xyz9"], wall[2:]]), uv6=np.concatenate([wall_uv[:2], robot["uv6"], wall_uv[2:]]),
               mat=np.concatenate([np.int32([1, 1]), robot["mat"], np.int32([0, 0])]))
    out = {
        "r_cfg1": (robot, dict(tab["cfg1"][0], **hy), mats, {}, None),
        "r_cfg2": (robot, dict(tab["cfg2"][0], **hy), mats, tex2, None),
        "r_inside": (robot, dict(tab["cfg2"][0], **hy), mats, tex2, CAM_INSIDE),
        "r_inside_noclip": (robot, dict(tab["cfg2"][0], enable_clipping=0, **hy), mats, tex2, CAM_INSIDE),
        "r_mirror": (robot, dict(tab["cfg3_mirror5"][0], **hy), mirror, tab["cfg3_mirror5"][2], None),
        "r_big": (big, dict(image_width=256, image_height=144, compute_shadows=1, enable_ssaa=1, ssaa_factor=3, enable_diffuse_mapping=1,
                            enable_normal_mapping=1, **hy), mats, tex2, CAM_INSIDE),
        "r_ssao": (robot, dict(ssao_table(mats)["ssao_ssaa2"][0], **hy), mats, {}, None),
        "r_rough": (robot, dict(tab["cfg3"][0], **hy), tab["cfg3"][1], tab["cfg3"][2], None),
    }
    for mode in (1, 2, 3, 4):
        out[f"r_debug{mode}"] = (robot, dict(image_width=200, image_height=120, shading_method=mode, enable_ao_mapping=1, **hy), mats,
                                 {0: tex2[0]}, CAM_INSIDE if mode > 2 else None)
    return out


RASTER_PINNED = ["r_cfg1", "r_cfg2", "r_inside", "r_inside_noclip", "r_mirror", "r_big", "r_ssao", "r_debug1", "r_debug2", "r_debug3", "r_debug4"]
RASTER_SRAND = 20261019


def ssao_reference_seeds(rand_values):
    """The nine generator seeds of a single-threaded Renderer::post_process_ssao_SIMD run after srand(): the two
    default-constructed generators (renderer.cpp:1252-1253; xorshift.h:10,39) and the private copy of the parallel region
    consume 8 + 1 + 8 values of std::rand() first; _mm256_set_epi32's arguments are evaluated right to left, so lane 0 gets
    the next value, ..., lane 7 the eighth, and the scalar generator the ninth (pinned by tests/test_oracle_vs_reference.py)."""
    return np.asarray(rand_values[17:26], np.uint32)


def oracle_renderer(tracer, scene, kw, mats, tex, fov=FOV, light=LIGHT, cam=None):
    s = ob.default_settings(**kw)
    r = tracer.renderer()
    r.configure(s, fov)
    r.set_triangles(scene["xyz9"], scene.get("uv6"), scene.get("mat"))
    r.set_materials(mats)
    r.set_light(light)
    if cam is not None:
        r.set_camera_transform(cam)
    for slot, img in tex.items():
        if slot == "shapes":
            add_shapes(r, img)
        else:
            r.set_texture(slot, img)
    return r


def add_shapes(renderer, shapes):
    """shapes: [("sphere", center, radius, mat_index) | ("plane", point, normal, mat_index)], in order (analyticShape.h)."""
    for sh in shapes:
        if sh[0] == "sphere":
            renderer.add_sphere(sh[1], sh[2], sh[3])
        else:
            renderer.add_plane(sh[1], sh[2], sh[3])


def oracle_image(tracer, scene, kw, mats, tex, **k):
    r = oracle_renderer(tracer, scene, kw, mats, tex, **k)
    if kw.get("hybrid_rasterization_tracing"):
        return r.raster()[0]                   # raster_trace() + post_process() in sequential triangle order
    if tracer.kind == "oracle":
        return r.render()[0]
    sup, _ = r.trace_rows()            # compiled reference: seeded pixel loop, then its own downscale
    s = r.settings
    f = s.ssaa_factor if s.enable_ssaa else 1
    return tracer.downscale(sup, f) if f > 1 else sup


def product_renderer(lib, scene, kw, mats, tex, fov=FOV, light=LIGHT, cam=None, device=0):
    r = rt.Renderer(device, lib)
    s = r.render_settings()
    for k, v in kw.items():
        setattr(s, k, float(v) if k in ("displacement_mapping_strength", "ssao_radius", "ssao_amount") else int(v))
    r.change_render_size(s.image_width, s.image_height)
    r.change_camera_fov(fov)
    if cam is not None:
        r.set_camera_transform(cam)
    r.set_triangles(scene["xyz9"], scene.get("uv6"), scene.get("mat"))
    r.set_materials(mats)
    r.set_light_position(light)
    for slot, img in tex.items():
        if slot == "shapes":
            add_shapes(r.ctx, img)
        else:
            r.ctx.set_texture(slot, img)
    return r


def product_image(lib, scene, kw, mats, tex, **k):
    r = product_renderer(lib, scene, kw, mats, tex, **k)
    if kw.get("hybrid_rasterization_tracing"):
        r.raster_trace()                       # what RenderThread::run calls for this setting (QT/mainWindowThreads.cpp:46-49)
    else:
        r.ray_trace()
    r.post_process()
    img, stats = r.get_image(), r.last_stats()
    r.close()
    return img, stats


def random_rays(n, seed, box_lo, box_hi):
    """Half camera-like rays from the origin, half incoherent un-normalised rays from inside the box."""
    rng = np.random.default_rng(seed)
    h = n // 2
    o1 = np.zeros((h, 3), np.float32)
    d1 = rng.normal(size=(h, 3)).astype(np.float32)
    d1[:, 2] = -np.abs(d1[:, 2]) * 3
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True).astype(np.float32)
    o2 = rng.uniform(box_lo, box_hi, size=(n - h, 3)).astype(np.float32)
    d2 = rng.normal(size=(n - h, 3)).astype(np.float32)
    return np.concatenate([o1, o2]), np.concatenate([d1, d2])


def triangle_soup(n, seed, size=0.2, center=(0, 0, -4), spread=1.5):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-spread, spread, size=(n, 1, 3)) + np.asarray(center)
    return (c + rng.normal(scale=size, size=(n, 3, 3))).reshape(n, 9).astype(np.float32)


# ---- BASELINE.json configs at their FULL sizes (SURVEY.md section 8(d) inputs 1, 2, 3 and 5) ---------------------------
_FULL_CACHE = {}


def fullsize_table(mats):
    """cfg1 1280x720; cfg2 1280x720 x ssaa 2 with 2048^2 u8 maps; cfg3 1920x1080 with the 16-ray fan, recursion depth 1,
    seeded, roughness map and a 4096x2048 equirect sky.  Must stay in sync with tests/golden/make_golden_fullsize.py
    (which imports this function)."""
    if "tex" not in _FULL_CACHE:
        n = 2048
        _FULL_CACHE["tex"] = {0: scenes.noise_texture((n, n), 2), 1: scenes.noise_texture((n, n), 1, "rgb"),
                              2: scenes.normal_map_texture((n, n), 4), 3: scenes.noise_texture((n, n), 3),
                              4: scenes.sky_texture((2048, 4096))}
    t = _FULL_CACHE["tex"]
    mats3 = [dict(m) for m in mats]
    mats3[0].update(reflection=0.9, roughness=0.0, specular=(0.2, 0.2, 0.2), diffuse=(0.5, 0.5, 0.5))   # load_obj override, QT/mainwindow.cpp:259-262
    mats3[1].update(reflection=0.5, roughness=0.4)
    mats3 = precompute_materials(mats3)
    return {
        "cfg1_full": (dict(image_width=1280, image_height=720, compute_shadows=1), mats, {}),
        "cfg2_full": (dict(image_width=1280, image_height=720, compute_shadows=1, enable_ssaa=1, ssaa_factor=2, enable_ao_mapping=1,
                           enable_diffuse_mapping=1, enable_normal_mapping=1), mats, {k: t[k] for k in (0, 1, 2, 3)}),
        "cfg3_full": (dict(image_width=1920, image_height=1080, compute_shadows=1, rough_reflections_sample_count=16, max_recursion_depth=1,
                           enable_roughness_mapping=1, enable_skysphere=1, rng_seed=7), mats3, {3: t[3], 4: t[4]}),
    }


FULL_ROW_STEP = 24                       # the committed goldens keep every 24th row of the reference's frame
HAIR_FULL = dict(n_strands=15625, segments=16)        # 15 625 strands x 16 segments x 2 triangles x 2 sides = 1 000 000 triangles
HAIR_KW = dict(image_width=3840, image_height=2160, compute_shadows=1)
HAIR_BAND = (960, 1216, 16)              # rows of the 4K hair frame that the oracle / reference trace (row_begin, row_end, row_step)
