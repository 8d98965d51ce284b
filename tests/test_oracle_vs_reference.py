"""The oracle against the compiled reference itself (oracle/_ref), live: skipped where /root/reference was never
available to build it.  Pins, among others, the argument evaluation order of renderer.cpp:313."""
import numpy as np
import pytest

from tests import common


@pytest.mark.parametrize("params", [(12, 40), (10, 8), (4, 3), (0, 8)])
def test_random_rays(oracle, ref_strict, robot, params):
    o, d = common.random_rays(40000, 99, (-1.5, -2.5, -5.5), (1.5, 0.5, -2.5))
    a = ref_strict.bvh(robot["xyz9"], *params)
    b = oracle.bvh(robot["xyz9"], *params)
    assert a.stats() == b.stats()
    for x, y in zip(a.intersect(o, d), b.intersect(o, d)):
        assert np.array_equal(x, y)


def test_triangle_soup(oracle, ref_strict):
    soup = common.triangle_soup(5000, 3)
    o, d = common.random_rays(20000, 4, (-1.5, -1.5, -5.5), (1.5, 1.5, -2.5))
    for params in ((12, 40), (6, 2)):
        a, b = ref_strict.bvh(soup, *params), oracle.bvh(soup, *params)
        assert a.stats() == b.stats()
        for x, y in zip(a.intersect(o, d), b.intersect(o, d)):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg3_mirror5"])
def test_images(oracle, ref_strict, robot, name):
    kw, mats, tex = common.config_table(robot["materials"])[name]
    kw = dict(kw, image_width=kw["image_width"] // 2, image_height=kw["image_height"] // 2, rng_seed=11)
    assert np.array_equal(common.oracle_image(oracle, robot, kw, mats, tex), common.oracle_image(ref_strict, robot, kw, mats, tex))


def test_rough_reflection_depth2_and_normal_mapping(oracle, ref_strict, robot):
    kw, mats, tex = common.config_table(robot["materials"])["cfg3"]
    tex = dict(tex)
    tex[2] = common.scenes.normal_map_texture((64, 64), 8)
    kw = dict(kw, image_width=96, image_height=54, max_recursion_depth=2, rough_reflections_sample_count=5, enable_normal_mapping=1)
    assert np.array_equal(common.oracle_image(oracle, robot, kw, mats, tex), common.oracle_image(ref_strict, robot, kw, mats, tex))
