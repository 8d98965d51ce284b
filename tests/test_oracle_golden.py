"""The oracle (oracle/oracle.cpp) against the golden vectors cut from the compiled reference (tests/golden/)."""
import numpy as np
import pytest

from tests import common


@pytest.mark.parametrize("params", [(12, 40), (10, 8)])
def test_bvh_intersect_matches_reference_vectors(oracle, robot, golden_rays, params):
    depth, leaf = params
    tag = f"_{depth}_{leaf}"
    bvh = oracle.bvh(robot["xyz9"], depth, leaf)
    st = bvh.stats()
    assert [st[k] for k in ("nodes", "leaves", "empty_leaves", "interior", "max_depth_reached", "max_leaf_size")] == list(golden_rays["stats" + tag])
    tri, t, u, v = bvh.intersect(golden_rays["o"], golden_rays["d"])
    # bit-exact against the reference built without FMA contraction
    assert np.array_equal(tri, golden_rays["tri" + tag])
    assert np.array_equal(t, golden_rays["t" + tag]) and np.array_equal(u, golden_rays["u" + tag]) and np.array_equal(v, golden_rays["v" + tag])
    # and >= 99.9 % identical triangle ids against the reference built with its own flags (-mfma, contraction on)
    assert (tri == golden_rays["tri_fma" + tag]).mean() >= 0.999
    assert (tri >= 0).sum() > 1000


def test_reference_kats(oracle, golden_images):
    """tp2/projets/tests.cpp:97-112: four rays that must miss (back face, outside, two grazing) + two hits."""
    g = golden_images
    for i in range(len(g["kat_tris"])):
        hit, t, u, v = oracle.triangle_intersect(g["kat_tris"][i], g["kat_o"][i], g["kat_d"][i])
        assert hit == bool(g["kat_hit"][i])
        if hit:
            assert np.array_equal(np.float32([t, u, v]), g["kat_tuv"][i])
    assert list(g["kat_hit"][:4]) == [False] * 4


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg2_pom", "cfg3", "cfg3_mirror5", "cfg3_skybox", "cfg3_shapes"])
def test_images_match_reference(oracle, robot, golden_images, name):
    kw, mats, tex = common.config_table(robot["materials"])[name]
    img = common.oracle_image(oracle, robot, kw, mats, tex)
    assert np.array_equal(img, golden_images[name + "_strict"])          # bit-exact vs the no-contraction build
    common.assert_image_close(img, golden_images[name + "_fma"], what=name + " vs reference flags")


def test_camera_matrices_and_inverse(oracle, golden_images):
    p, pi = oracle.camera_matrices(80.0, 16.0 / 9.0)
    assert np.array_equal(p, golden_images["proj_80_16x9"]) and np.array_equal(pi, golden_images["proj_inv_80_16x9"])
    assert np.array_equal(oracle.camera_matrices(45.0, 1.0)[1], golden_images["proj_inv_45_1"])
    assert np.array_equal(oracle.transform_inverse(golden_images["inv_in"]), golden_images["inv_out"])


@pytest.mark.parametrize("factor", [2, 3, 4])
def test_ssaa_resolve(oracle, golden_images, factor):
    assert np.array_equal(oracle.downscale(golden_images["resolve_in"], factor), golden_images[f"resolve_out_{factor}"])


# ---- BASELINE.json configs at their FULL sizes: the oracle is bit-exact against the compiled reference there too ------
@pytest.mark.parametrize("name", ["cfg1_full", "cfg2_full", "cfg3_full"])
def test_fullsize_frames_match_reference_crc(oracle, robot, golden_fullsize, name):
    import zlib
    kw, mats, tex = common.fullsize_table(robot["materials"])[name]
    img = common.oracle_image(oracle, robot, kw, mats, tex)
    assert img.shape == (kw["image_height"], kw["image_width"])
    assert zlib.crc32(np.ascontiguousarray(img, np.uint32).tobytes()) == int(golden_fullsize[name + "_crc"])
    assert np.array_equal(img[::common.FULL_ROW_STEP], golden_fullsize[name + "_rows"])
    r = common.oracle_renderer(oracle, robot, kw, mats, tex)
    assert r.count_rows()["shadow_rays"] == int(golden_fullsize[name + "_hits"])


def test_hair_band_matches_reference(oracle, golden_fullsize):
    """cfg5 at its BASELINE size (1 M strand triangles, 3840x2160, shadows): a band of rows, bit-exact."""
    import zlib
    import raytracercpp_b200 as rt
    from raytracercpp_b200 import scenes
    xyz9, uv6, mat = scenes.hair_ball(**common.HAIR_FULL)
    assert len(xyz9) == 1_000_000
    r = common.oracle_renderer(oracle, dict(xyz9=xyz9, uv6=uv6, mat=mat), common.HAIR_KW, rt.precompute_materials([scenes.DEFAULT_SPHERE_MATERIAL]), {})
    b0, b1, bs = common.HAIR_BAND
    sup, _ = r.trace_rows(row_begin=b0, row_end=b1, row_step=bs)
    band = sup[b0:b1:bs]
    assert np.array_equal(band, golden_fullsize["cfg5_band_rows"])
    assert zlib.crc32(np.ascontiguousarray(band, np.uint32).tobytes()) == int(golden_fullsize["cfg5_band_crc"])
    assert r.count_rows(row_begin=b0, row_end=b1, row_step=bs)["shadow_rays"] == int(golden_fullsize["cfg5_band_hits"])


@pytest.mark.parametrize("name", ["ssao_ssaa2", "ssao_normal_mapped"])
def test_ssao_matches_reference(oracle, robot, golden_ssao, name):
    """SSAO frames cut from the compiled reference (tests/golden/make_golden_ssao.py): the oracle in reference order, from the
    generator seeds stored with the frame, is bit-exact; its per-pixel stream (what the CUDA path is compared with) is
    stored too, so a change of that discipline shows up here."""
    kw, mats, tex = common.ssao_table(robot["materials"])[name]
    orc = common.oracle_renderer(oracle, robot, kw, mats, tex)
    got, _ = orc.render_ssao(ref_seeds9=golden_ssao[name + "_seeds9"])
    assert np.array_equal(got, golden_ssao[name + "_reference"])
    per_pixel, _ = orc.render_ssao()
    assert np.array_equal(per_pixel, golden_ssao[name + "_per_pixel"])


@pytest.mark.parametrize("name", common.RASTER_PINNED)
def test_raster_trace_matches_reference(oracle, robot, golden_raster, name):
    """Renderer::raster_trace + post_process (hybrid_rasterization_tracing, renderer.cpp:869-1006): frames cut from the
    compiled reference run on one thread (tests/golden/make_golden_raster.py) -- clipping on and off, the camera inside the
    scene (every clip plane cuts), frame-filling triangles, textures + SSAA, mirror reflections, the four debug modes,
    SSAO on the rasterizer's z-buffer.  Bit-exact."""
    scene, kw, mats, tex, cam = common.raster_table(robot)[name]
    r = common.oracle_renderer(oracle, scene, kw, mats, tex, cam=cam)
    img, counters = r.raster(ref_seeds9=golden_raster[name + "_seeds9"] if kw.get("enable_ssao") else None)
    assert np.array_equal(img, golden_raster[name + "_reference"])
    if kw.get("enable_ssao"):
        assert np.array_equal(r.raster()[0], golden_raster[name + "_per_pixel"])
    if kw.get("shading_method", 0) == 0:
        assert counters["fragments"] > 0 and counters["shadow_rays"] == counters["fragment_hits"]
