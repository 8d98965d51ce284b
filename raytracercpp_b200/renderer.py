"""Host-side mirror of the reference's `Renderer` interface for the ray-tracing path
(tp2/projets/renderer/renderer.h:38-169): same method names, argument meaning and call order, implemented as
one C-ABI call each (include/rtb200.h).  The C++ twin is include/rtb200_renderer.hpp; this Python one exists so
that tests and bench.py read like the reference's own harness (utils/mainUtils.cpp:6-21).

Members that belong to the rasterizer / SSAO are not on the path and are
absent on purpose (the reference GUI keeps calling its own code for those).
"""
from __future__ import annotations

import time

import numpy as np

from . import api


def precompute_materials(mats):
    """MainWindow::precompute_materials (QT/mainwindow.cpp:240-249): specular_threshold = pow(eps / luminance, 1 / ns).
    `Material` leaves the field uninitialised otherwise (materials.h:35-37), so every caller has to do this."""
    out = []
    for m in mats:
        m = dict(m)
        s = [np.float32(x) for x in m.get("specular", (0, 0, 0))]
        lum = np.float32(np.float32(0.2126) * s[0] + np.float32(0.7152) * s[1]) + 0.0722 * float(s[2])   # double, as written
        lum = np.float32(lum)
        with np.errstate(divide="ignore", invalid="ignore"):
            tau = np.power(np.float32(np.float32(1.0e-3) / lum), np.float32(1.0) / np.float32(m.get("ns", 0.0)), dtype=np.float32)
        m["specular_threshold"] = float(tau)
        out.append(m)
    return out


class Renderer:
    """Drop-in for the ray-tracing part of `class Renderer`."""

    def __init__(self, device: int = 0, lib=None):
        self.ctx = api.Context(device, lib)
        self._settings = api.default_settings(self.ctx.lib)
        self._materials: list[dict] = []
        self._fov = 45.0                                   # Camera() default, scene/camera.h:11
        self._near, self._far = 0.1, 1000.0
        self._aspect = 1.0
        self._camera_to_world = np.eye(4, dtype=np.float32)
        self._position = np.zeros(3, np.float32)
        self._previous_object_transform = np.eye(4, dtype=np.float32)
        self._image = None
        self._stats = None
        self._bvh_dirty = True
        self._have_triangles = False
        self.change_render_size(self._settings.image_width, self._settings.image_height)

    # -- settings ---------------------------------------------------------------------------------------------
    def render_settings(self) -> api.RtSettings:
        """Mutable reference, as Renderer::render_settings() (renderer.cpp:111-114)."""
        return self._settings

    def get_render_width_height(self):
        s = self._settings
        f = s.ssaa_factor if s.enable_ssaa else 1          # renderer.cpp:116-120
        return s.image_width * f, s.image_height * f

    def change_render_size(self, width: int, height: int):
        self._settings.image_width, self._settings.image_height = int(width), int(height)
        rw, rh = self.get_render_width_height()            # renderer.cpp:250-261
        self.change_camera_aspect_ratio(np.float32(rw) / np.float32(rh))

    # -- geometry ---------------------------------------------------------------------------------------------
    def set_triangles(self, xyz9, uv6=None, mat=None):
        """Renderer::set_triangles (renderer.cpp:137-144): copies the triangles and builds the BVH."""
        self.ctx.set_triangles(xyz9, uv6, mat)
        self._have_triangles = True
        self.reconstruct_bvh_new()

    def load_obj(self, filepath, transform=None):
        """MainWindow::load_obj (QT/mainwindow.cpp:251-282): the native loader (rt_obj_load = read_meshio_data +
        MeshIOUtils::create_triangles), the GUI's overrides of material 0, set_triangles, the mesh's materials appended to
        the renderer's, precompute_materials, reset_previous_transform."""
        xyz9, uv6, mat, mats, _ = api.load_obj(filepath, transform, len(self._materials), self.ctx.lib)
        if mats:
            mats[0].update(roughness=0.0, reflection=0.9, specular=(0.2, 0.2, 0.2), diffuse=(0.5, 0.5, 0.5))
        self.set_triangles(xyz9, uv6, mat)
        self.set_materials(precompute_materials(self._materials + mats))
        self.reset_previous_transform()

    def reconstruct_bvh_new(self):
        s = self._settings                                 # renderer.cpp:243-246
        self.bvh_info = self.ctx.build_bvh(s.bvh_max_depth, s.bvh_leaf_object_count)
        self._bvh_dirty = False

    def reset_previous_transform(self):
        self._previous_object_transform = np.eye(4, dtype=np.float32)

    def set_object_transform(self, object_transform):
        """Renderer::set_object_transform (renderer.cpp:214-224): undo the previous transform, apply the new one,
        rebuild the tree."""
        m = np.asarray(object_transform, dtype=np.float32).reshape(4, 4)
        prev_inv = self.ctx.invert_transform(self._previous_object_transform)
        composed = _compose(m, prev_inv)
        s = self._settings
        self.ctx.transform_triangles(composed, s.bvh_max_depth, s.bvh_leaf_object_count)
        self.bvh_info = self.ctx.bvh_info()
        self._previous_object_transform = m.copy()

    # -- materials / textures -----------------------------------------------------------------------------------
    def get_materials(self):
        return self._materials

    def set_materials(self, materials):
        self._materials = [dict(m) for m in materials]
        self.ctx.set_materials(self._materials)

    def set_ao_map(self, image): self.ctx.set_texture(api.RT_TEX_AO, image)
    def set_diffuse_map(self, image): self.ctx.set_texture(api.RT_TEX_DIFFUSE, image)
    def set_normal_map(self, image): self.ctx.set_texture(api.RT_TEX_NORMAL, image)
    def set_roughness_map(self, image): self.ctx.set_texture(api.RT_TEX_ROUGHNESS, image)
    def set_displacement_map(self, image): self.ctx.set_texture(api.RT_TEX_DISPLACEMENT, image)
    def set_skysphere(self, image): self.ctx.set_texture(api.RT_TEX_SKYSPHERE, image)

    def set_skybox(self, faces):
        """Renderer::set_skybox(Skybox(faces)) -- renderer.cpp:199; faces: right, left, top, bottom, back, front (skybox.h:13-17)."""
        assert len(faces) == 6
        for i, image in enumerate(faces):
            self.ctx.set_texture(api.RT_TEX_SKYBOX_RIGHT + i, image)

    def clear_ao_map(self): self.ctx.clear_texture(api.RT_TEX_AO)
    def clear_diffuse_map(self): self.ctx.clear_texture(api.RT_TEX_DIFFUSE)
    def clear_normal_map(self): self.ctx.clear_texture(api.RT_TEX_NORMAL)
    def clear_roughness_map(self): self.ctx.clear_texture(api.RT_TEX_ROUGHNESS)
    def clear_displacement_map(self): self.ctx.clear_texture(api.RT_TEX_DISPLACEMENT)

    # -- camera / light -------------------------------------------------------------------------------------------
    def change_camera_fov(self, fov: float):
        self._fov = float(fov)                             # Camera::set_fov, scene/camera.cpp:13-19
        self._push_camera()

    def change_camera_aspect_ratio(self, aspect: float):
        self._aspect = float(aspect)                       # Camera::set_aspect_ratio, scene/camera.cpp:5-11
        self._push_camera()

    def set_camera_transform(self, camera_transform):
        self._camera_to_world = np.asarray(camera_transform, dtype=np.float32).reshape(4, 4).copy()   # renderer.cpp:226-233
        self._position = self.ctx.transform_point(self._camera_to_world, (0, 0, 0))
        self._push_camera()

    def set_light_position(self, position):
        self.ctx.set_light(position)

    def add_analytic_shape(self, shape):
        """Renderer::add_analytic_shape -- renderer.cpp:146.  shape = ("sphere", center, radius, mat_index) like
        Sphere(center, radius, mat_index), or ("plane", point, normal, mat_index) like Plane(point, normal, mat_index)."""
        if shape[0] == "sphere":
            self.ctx.add_sphere(shape[1], shape[2], shape[3])
        elif shape[0] == "plane":
            self.ctx.add_plane(shape[1], shape[2], shape[3])
        else:
            raise ValueError(shape[0])

    def _push_camera(self):
        proj_inv = self.ctx.perspective_inverse(self._fov, self._aspect, self._near, self._far)
        self.ctx.set_camera(proj_inv, self._camera_to_world, self._position)
        self.ctx.set_projection(self._fov, self._aspect, self._near, self._far)       # read by the SSAO pass and by raster_trace

    # -- rendering ------------------------------------------------------------------------------------------------
    def ray_trace(self):
        """Renderer::ray_trace() + the SSAA half of post_process() (renderer.cpp:1068-1135).  The resolve runs on the
        device in the same call -- and, with enable_ssao, the SSAO pass before it (renderer.cpp:1229-1434) -- so post_process() has
        nothing left to do for this path."""
        self._render(0)

    def raster_trace(self):
        """Renderer::raster_trace() (renderer.cpp:869-1006), what RenderThread::run calls instead of ray_trace() when
        render_settings().hybrid_rasterization_tracing is set (QT/mainWindowThreads.cpp:46-49): primary visibility by
        clipping + rasterisation + z-buffer on the device, shading with shadow rays and reflection fans as in ray_trace().
        The image starts from clear_image() (QT/mainwindow.cpp:186-190): uncovered pixels are the background colour."""
        self._render(1)

    def _render(self, hybrid: int):
        rw, rh = self.get_render_width_height()
        self._aspect = float(np.float32(rw) / np.float32(rh))
        self._push_camera()
        keep = self._settings.hybrid_rasterization_tracing
        self._settings.hybrid_rasterization_tracing = hybrid
        try:
            self._image, self._stats = self.ctx.render(self._settings)
        finally:
            self._settings.hybrid_rasterization_tracing = keep

    def post_process(self):
        """Renderer::post_process (renderer.cpp:1118-1124): SSAO and the SSAA resolve ran on the device inside ray_trace()."""

    def get_image(self):
        """ARGB32 [H, W], row 0 = bottom row like the reference's QImage (renderer.cpp:1086)."""
        return self._image

    def last_stats(self) -> api.RtRenderStats:
        return self._stats

    def close(self):
        self.ctx.close()


def _compose(a, b):
    """compose_transform(a, b) = a * b in float32, row-major (mat.cpp)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    out = np.zeros((4, 4), np.float32)
    for i in range(4):
        for j in range(4):
            acc = np.float32(0)
            for k in range(4):
                acc = np.float32(acc + np.float32(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def render(renderer: Renderer) -> float:
    """The reference's timed harness entry `render(Renderer&)` (utils/mainUtils.cpp:6-21): trace + post-process, ms."""
    t0 = time.perf_counter()
    if renderer.render_settings().hybrid_rasterization_tracing:
        renderer.raster_trace()
    else:
        renderer.ray_trace()
    renderer.post_process()
    return (time.perf_counter() - t0) * 1e3
